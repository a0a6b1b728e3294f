"""Frame-sharding host logic (v3d_b200.viewshard) on CPU, world_size 2 and 3 over gloo.

Every rank holds the same seeded full-video tensors, computes the frame-coupled operators for its own block of frames
using only the exchanges of ViewShard, and compares with the unsharded computation:
  * 3-D GroupNorm + SiLU + Conv3d (3,1,1) (openaimodel.py:267-271 dims=3; video_model.py:42-55) through the fp64
    statistics all-reduce (rescaled the way v3d_groupnorm_apply consumes them) and the halo exchange, with the 3-tap
    operand addressed exactly like v3d_gemm_bf16's a_rows / a_row0 halo mode;
  * temporal self-attention (video_attention.py:114-125) through the K|V all-gather and the per-frame row table that
    v3d_attention_temporal_kv walks;
  * the final frame gather, the per-frame slicing of the conditioning, the guider's scale slice and the time context.
The arithmetic here is plain fp32 torch standing in for the kernels; the CUDA path itself is checked against the
unsharded CUDA path in tests/test_viewshard_gpu.py.
"""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

ROOT = str(Path(__file__).resolve().parent.parent)
NB, C, HW, HEADS = 2, 64, 4, 1


def _full_inputs(T: int):
    g = torch.Generator().manual_seed(1234 + T)
    x = torch.randn(NB, T, HW, C, generator=g)                 # frame-major token layout [b][t][s][c]
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    w = torch.randn(C, C, 3, generator=g) / (3 * C) ** 0.5      # [co][ci][tap]
    bias = torch.randn(C, generator=g)
    qkv = torch.randn(NB, T, HW, 3 * C, generator=g)
    return x, gamma, beta, w, bias, qkv


def _reference_gn_tconv(x, gamma, beta, w, bias, eps=1e-5):
    nb, T, hw, c = x.shape
    v = x.permute(0, 3, 1, 2).reshape(nb, c, T, hw, 1)         # [b, c, t, h, w]
    a = F.silu(F.group_norm(v, 32, gamma, beta, eps))
    out = F.conv3d(a, w.reshape(c, c, 3, 1, 1), bias, padding=(1, 0, 0))
    return out.reshape(nb, c, T, hw).permute(0, 2, 3, 1)       # back to [b][t][s][c]


def _reference_attention(qkv, scale):
    nb, T, hw, c3 = qkv.shape
    c = c3 // 3
    q, k, v = (t.permute(0, 2, 1, 3) for t in qkv.split(c, dim=-1))   # [b][s][t][c] ("(b s) t c")
    out = F.scaled_dot_product_attention(q.reshape(nb * hw, T, HEADS, c // HEADS).transpose(1, 2),
                                         k.reshape(nb * hw, T, HEADS, c // HEADS).transpose(1, 2),
                                         v.reshape(nb * hw, T, HEADS, c // HEADS).transpose(1, 2), scale=scale)
    return out.transpose(1, 2).reshape(nb, hw, T, c).permute(0, 2, 1, 3)


def _worker(rank: int, world: int, port: int, T: int, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist

    from v3d_b200.sampling import LinearPredictionGuider, VanillaCFG
    from v3d_b200.viewshard import ViewShard

    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    vs = ViewShard.create(T)
    tl, t0 = vs.tl, vs.t0
    x, gamma, beta, w, bias, qkv = _full_inputs(T)
    err = {}

    # ---- 3-D GroupNorm + SiLU + temporal conv --------------------------------------------------------------
    xl = x[:, vs.frames]                                                       # [nb, tl, hw, c]
    grp = xl.reshape(NB, tl * HW, 32, C // 32).double()
    stats = torch.stack([grp.sum(dim=(1, 3)), (grp * grp).sum(dim=(1, 3))], dim=-1).contiguous()  # [nb, 32, 2]
    vs.allreduce_stats_(stats)
    n_local = tl * HW * (C // 32)                                              # what the apply kernel divides by
    mean = stats[..., 0] / n_local
    var = stats[..., 1] / n_local - mean * mean
    rstd = (var + 1e-5).rsqrt()
    norm = ((xl.reshape(NB, tl * HW, 32, C // 32).double() - mean[:, None, :, None]) * rstd[:, None, :, None])
    a = F.silu(norm.float().reshape(NB, tl, HW, C) * gamma + beta)
    pad = torch.full((NB, tl + 2, HW, C), float("nan"))                        # halos must be overwritten
    pad[:, 1:tl + 1] = a
    vs.exchange_halos(pad)
    # v3d_gemm_bf16 halo mode: a_rows = (tl+2)*hw per batch item, a_row0 = hw, tap_shift = hw:
    # output row r of batch b reads A rows r + tap*hw (tap 0..2) of the padded item
    rows = pad.reshape(NB, (tl + 2) * HW, C)
    out = bias.expand(NB, tl * HW, C).clone()
    for tap in range(3):
        out += rows[:, tap * HW: tap * HW + tl * HW] @ w[:, :, tap].t()
    ref = _reference_gn_tconv(x, gamma, beta, w, bias)[:, vs.frames].reshape(NB, tl * HW, C)
    err["gn_tconv"] = float((out - ref).abs().max())

    # ---- temporal attention through the K|V all-gather ------------------------------------------------------
    scale = (C // HEADS) ** -0.5
    ql = qkv[:, vs.frames].reshape(NB * tl * HW, 3 * C)                         # local rows (b, t, s)
    send = torch.full((NB * vs.tmax * HW, 2 * C), float("nan"))
    send[: NB * tl * HW] = ql[:, C:]
    buf = vs.gather_rows(send)
    row, bstride = vs.kv_table(NB, HW)
    assert len(row) == T
    o = torch.empty(NB, tl, HW, C)
    for b in range(NB):
        for s in range(HW):
            kv_rows = torch.tensor([row[f] + b * bstride[f] + s for f in range(T)])
            k, v = buf[kv_rows, :C], buf[kv_rows, C:]
            qq = ql.reshape(NB, tl, HW, 3 * C)[b, :, s, :C]
            p = torch.softmax(qq @ k.t() * scale, dim=-1)
            o[b, :, s] = p @ v
    ref = _reference_attention(qkv, scale)[:, vs.frames]
    err["attention"] = float((o - ref).abs().max())

    # ---- frame gather, conditioning slices, guider ----------------------------------------------------------
    frames = torch.arange(t0, t0 + tl, dtype=torch.uint8).reshape(tl, 1, 1, 1).expand(tl, 2, 2, 3).contiguous()
    gathered = vs.gather_frames(frames)
    ok_gather = gathered[:, 0, 0, 0].tolist() == list(range(T))
    cond = {"vector": torch.arange(T * 3.0).reshape(T, 3), "crossattn": torch.arange(T * 2.0).reshape(T, 1, 2),
            "concat": torch.zeros(T, 4, 2, 2), "flag": 7}
    uc = {k: (v + 100 if torch.is_tensor(v) else v) for k, v in cond.items()}
    cl = vs.shard_cond(cond)
    ok_cond = (torch.equal(cl["vector"], cond["vector"][t0:t0 + tl]) and cl["flag"] == 7
               and cl["concat"].shape[0] == tl)
    tc = vs.time_context(cond, uc)
    ok_tc = torch.equal(tc, torch.stack([uc["crossattn"][0], cond["crossattn"][0]]))
    g = LinearPredictionGuider(max_scale=3.5, min_scale=1.0, num_frames=T)
    gl = vs.shard_guider(g)
    ok_guider = (gl.num_frames == tl and torch.equal(gl.scale, g.scale[:, t0:t0 + tl]) and g.num_frames == T
                 and vs.shard_guider(VanillaCFG(2.0)).num_frames == 1)
    q.put((rank, (t0, tl), err, ok_gather, ok_cond, ok_tc, ok_guider, dict(vs.exchanges)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,T", [(2, 5), (3, 7)])
def test_view_shard_exchanges_match_unsharded(world, T):
    sys.path.insert(0, str(Path(ROOT) / "tests"))
    from mp_util import run_workers

    res = sorted(run_workers(_worker, world, (T,), timeout=180))
    covered = []
    for rank, (t0, tl), err, ok_gather, ok_cond, ok_tc, ok_guider, counts in res:
        covered += list(range(t0, t0 + tl))
        assert err["gn_tconv"] < 2e-5, (rank, err)     # fp32 arithmetic on both sides; fp64 statistics
        assert err["attention"] < 2e-6, (rank, err)
        assert ok_gather and ok_cond and ok_tc and ok_guider, (rank, ok_gather, ok_cond, ok_tc, ok_guider)
        assert counts == {"gn_allreduce": 1, "halo": 1, "kv_allgather": 1, "frame_gather": 1}
    assert covered == list(range(T))                    # contiguous blocks in rank order, no frame lost


def test_partition_frames_properties():
    from v3d_b200.viewshard import ViewShard, partition_frames

    assert partition_frames(18, 2) == [(0, 9), (9, 9)]
    assert [n for _, n in partition_frames(18, 4)] == [5, 5, 4, 4]
    assert [n for _, n in partition_frames(18, 8)] == [3, 3, 2, 2, 2, 2, 2, 2]
    assert [n for _, n in partition_frames(24, 8)] == [3] * 8
    for T in (14, 18, 24, 25):
        for w in (1, 2, 4, 8):
            blocks = partition_frames(T, w)
            assert blocks[0][0] == 0 and sum(n for _, n in blocks) == T
            assert all(blocks[i][0] + blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            assert max(n for _, n in blocks) - min(n for _, n in blocks) <= 1
    with pytest.raises(ValueError):
        partition_frames(3, 4)
    # row table of the K|V gather buffer: rank-major blocks padded to the largest block
    vs = ViewShard(num_frames=5, rank=1, world=2)
    row, bstride = vs.kv_table(nb=2, rows_per_frame=4)
    assert (vs.t0, vs.tl, vs.tmax) == (3, 2, 3)
    assert row == [0, 4, 8, 24, 28] and bstride == [12, 12, 12, 8, 8]


def test_single_rank_view_shard_is_identity():
    """world 1: no process group needed for the exchanges; halos are the zero padding of the reference."""
    from v3d_b200.viewshard import ViewShard

    vs = ViewShard(num_frames=4, rank=0, world=1)
    pad = torch.ones(2, 6, 3, 8)
    vs.exchange_halos(pad)
    assert float(pad[:, 0].abs().sum()) == 0 and float(pad[:, 5].abs().sum()) == 0 and float(pad[:, 1:5].min()) == 1
    stats = torch.ones(2, 32, 2, dtype=torch.float64)
    assert torch.equal(vs.allreduce_stats_(stats), torch.ones(2, 32, 2, dtype=torch.float64))
    send = torch.randn(8, 4)
    assert vs.gather_rows(send) is send
    assert vs.kv_table(2, 1) == ([0, 1, 2, 3], [4, 4, 4, 4])
