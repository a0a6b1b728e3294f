"""View-sharding ONE image across GPUs: the T view-frames of a video are split into contiguous blocks, one block
per rank (BASELINE.json configs[2-3]).

Everything that is per-frame or per-token stays rank-local (all spatial layers, the temporal feed-forwards and
projections: ~93 % of the UNet's FLOPs).  Three things couple the frames and each gets one exchange:

  * temporal self-attention (video_attention.py:114-125 around attention.py:337-341): every query frame attends to
    the keys/values of all T frames -> all-gather of the packed K|V rows (`gather_rows`), consumed in place by
    v3d_attention_temporal_kv through a per-frame row table (`kv_table`), so uneven blocks need no compaction;
  * the (3,1,1) temporal convolutions (video_model.py:42-55, temporal_ae.py:94-99): one halo frame from each
    neighbour (`exchange_halos`); the video's first / last frame keep the reference's ZERO padding;
  * the 3-D GroupNorms of the time_stack ResBlocks (openaimodel.py:267-271 with dims=3): (sum, sum of squares)
    all-reduce of the fp64 statistics (`allreduce_stats_`).

The CFG pair [uc; c] of a frame lives on the same rank, so the guidance combine and the Euler update are local.
This module is host logic over torch.distributed only (NCCL on GPUs; gloo in the CPU tests, and - staged through
host memory - for 2 processes sharing one GPU in tests/test_viewshard_gpu.py); it launches no kernels of its own.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


def partition_frames(num_frames: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous (first frame, count) per rank; the first `num_frames % world` ranks take one extra frame.
    T=18: 2 ranks 9/9, 4 ranks 5/5/4/4, 8 ranks 3/3/2/2/2/2/2/2."""
    if world <= 0 or num_frames < world:
        raise ValueError(f"cannot split {num_frames} frames over {world} ranks (every rank needs a frame)")
    base, rem = divmod(num_frames, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((start, n))
        start += n
    return out


@dataclass
class ViewShard:
    """This rank's block of a T-frame video and the exchanges that stitch the blocks together."""

    num_frames: int                      # T of the whole video
    rank: int
    world: int
    group: Optional[object] = None       # torch.distributed process group (None = default group)
    blocks: List[Tuple[int, int]] = field(default_factory=list)
    exchanges: Dict[str, int] = field(default_factory=dict)  # per-kind call counters (bench / tests)

    @classmethod
    def create(cls, num_frames: int, group=None) -> "ViewShard":
        if not dist.is_initialized():
            raise RuntimeError("ViewShard.create needs an initialised torch.distributed process group")
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        return cls(num_frames, rank, world, group, partition_frames(num_frames, world))

    def __post_init__(self):
        if not self.blocks:
            self.blocks = partition_frames(self.num_frames, self.world)
        self._staged: Optional[bool] = None

    # ---- geometry -------------------------------------------------------------------------------
    @property
    def t0(self) -> int:
        return self.blocks[self.rank][0]

    @property
    def tl(self) -> int:
        return self.blocks[self.rank][1]

    @property
    def tmax(self) -> int:
        return max(n for _, n in self.blocks)

    @property
    def frames(self) -> slice:
        return slice(self.t0, self.t0 + self.tl)

    def kv_table(self, nb: int, rows_per_frame: int) -> Tuple[List[int], List[int]]:
        """Row offsets of every key frame in the buffer `gather_rows` returns for a [nb, tl, rows_per_frame] block
        per rank: frame f of video b, pixel s sits at row[f] + b * bstride[f] + s."""
        row, bstride = [], []
        block_rows = nb * self.tmax * rows_per_frame      # every rank's slot is padded to the largest block
        for r, (_, n) in enumerate(self.blocks):
            for j in range(n):
                row.append(r * block_rows + j * rows_per_frame)
                bstride.append(n * rows_per_frame)
        return row, bstride

    # ---- transport ------------------------------------------------------------------------------
    def _is_gloo(self) -> bool:
        if self._staged is None:
            self._staged = dist.get_backend(self.group) == "gloo"
        return self._staged

    def _via_host(self, t: torch.Tensor) -> bool:
        """gloo cannot move CUDA memory for every primitive: stage through the host (tests only; NCCL is direct)."""
        return self._is_gloo() and t.is_cuda

    def _count(self, kind: str) -> None:
        self.exchanges[kind] = self.exchanges.get(kind, 0) + 1

    def _global_rank(self, r: int) -> int:
        return r if self.group is None else dist.get_global_rank(self.group, r)

    def allreduce_stats_(self, stats: torch.Tensor) -> torch.Tensor:
        """stats: fp64 [nsamples, groups, 2] = local (sum, sumsq) of a 3-D GroupNorm.  After the call it holds the
        GLOBAL sums rescaled by tl / T, so that the apply kernel - which divides by the LOCAL element count - yields
        the global mean and variance."""
        self._count("gn_allreduce")
        if self.world > 1:
            if self._via_host(stats):
                h = stats.cpu()
                dist.all_reduce(h, group=self.group)
                stats.copy_(h)
            else:
                dist.all_reduce(stats, group=self.group)
        stats.mul_(self.tl / self.num_frames)
        return stats

    def exchange_halos(self, pad: torch.Tensor) -> torch.Tensor:
        """pad: [nb, tl + 2, ...] contiguous; frames 1..tl are this rank's, frame 0 / tl+1 receive the previous /
        next rank's last / first frame.  The ends of the video get zeros (Conv3d padding (1,0,0))."""
        self._count("halo")
        nb, tl = pad.shape[0], self.tl
        assert pad.shape[1] == tl + 2 and pad.is_contiguous()
        prev_r, next_r = self.rank - 1, self.rank + 1
        if prev_r < 0:
            pad[:, 0].zero_()
        if next_r >= self.world:
            pad[:, tl + 1].zero_()
        if self.world == 1:
            return pad
        staged = self._via_host(pad)
        tagged = self._is_gloo()
        ops, landing = [], []
        for b in range(nb):
            for peer, send_idx, recv_idx, tag_s, tag_r in ((prev_r, 1, 0, 0, 1), (next_r, tl, tl + 1, 1, 0)):
                if peer < 0 or peer >= self.world:
                    continue
                src, dst = pad[b, send_idx], pad[b, recv_idx]
                if staged:
                    src = src.cpu()
                    host = torch.empty(dst.shape, dtype=dst.dtype)
                    landing.append((dst, host))
                    dst = host
                g = self._global_rank(peer)
                # a frame sent "towards lower ranks" (tag_s 0) is received by the peer as its right halo (tag_r 0).
                # Tags disambiguate the messages on gloo; NCCL matches sends and receives of a pair in issue order
                # (the order here is the same on both sides) and takes no tags.
                ts_, tr_ = (2 * b + tag_s, 2 * b + tag_r) if tagged else (0, 0)
                ops.append(dist.P2POp(dist.isend, src, g, group=self.group, tag=ts_))
                ops.append(dist.P2POp(dist.irecv, dst, g, group=self.group, tag=tr_))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for dst, host in landing:
            dst.copy_(host)
        return pad

    def gather_rows(self, send: torch.Tensor) -> torch.Tensor:
        """send: [nb * tmax * rows_per_frame, W], of which this rank filled its first nb * tl * rows_per_frame rows.
        Returns [world * rows, W] with rank r's block at r * rows (see `kv_table`)."""
        self._count("kv_allgather")
        if self.world == 1:
            return send
        out = send.new_empty((self.world * send.shape[0],) + tuple(send.shape[1:]))
        if self._via_host(send):
            h = torch.empty(out.shape, dtype=out.dtype)
            dist.all_gather_into_tensor(h, send.cpu(), group=self.group)
            out.copy_(h)
        else:
            dist.all_gather_into_tensor(out, send.contiguous(), group=self.group)
        return out

    def gather_frames(self, local: torch.Tensor) -> torch.Tensor:
        """[tl, ...] per rank -> [T, ...] on every rank in frame order (the final decoded-frame gather)."""
        self._count("frame_gather")
        if self.world == 1:
            return local
        pad = local.new_zeros((self.tmax,) + tuple(local.shape[1:]))
        pad[: self.tl] = local
        out = local.new_empty((self.world * self.tmax,) + tuple(local.shape[1:]))
        if self._via_host(local):
            h = torch.empty(out.shape, dtype=out.dtype)
            dist.all_gather_into_tensor(h, pad.cpu(), group=self.group)
            out.copy_(h)
        else:
            dist.all_gather_into_tensor(out, pad, group=self.group)
        out = out.reshape(self.world, self.tmax, *local.shape[1:])
        return torch.cat([out[r, :n] for r, (_, n) in enumerate(self.blocks)], dim=0)

    # ---- sampler-side slicing -------------------------------------------------------------------
    def shard_cond(self, cond: Dict) -> Dict:
        """Per-frame conditioning tensors ([T, ...]: vector, crossattn, concat) -> this rank's frames; other entries
        pass through (guiders.py:88-98 treats them as shared)."""
        out = {}
        for k, v in cond.items():
            if torch.is_tensor(v) and v.ndim > 0 and v.shape[0] == self.num_frames:
                out[k] = v[self.frames].contiguous()
            else:
                out[k] = v
        return out

    def time_context(self, c: Dict, uc: Dict) -> torch.Tensor:
        """The temporal cross-attention context is `context[::T]` = frame 0 of each CFG half
        (video_attention.py:250): [2, 1, ctx] in [uc; c] order, the same on every rank."""
        return torch.cat([uc["crossattn"][:1], c["crossattn"][:1]], dim=0).contiguous()

    def shard_guider(self, guider):
        """Copy of a per-frame guider (guiders.py:60-146) restricted to this rank's frames."""
        if getattr(guider, "num_frames", 1) != self.num_frames:
            return guider  # VanillaCFG / IdentityGuider: one scale for every sample
        g = copy.copy(guider)
        g.num_frames = self.tl
        g.scale = guider.scale[:, self.frames].clone()
        if hasattr(g, "_scale_dev"):
            g._scale_dev = None
        return g


# ---------------------------------------------------------------------------------------------------------------------
# CFG-pair split: the second free axis of SURVEY.md 8(e)
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class CfgSplit:
    """The unconditional and the conditional half of the classifier-free-guidance batch (guiders.py:88-101, batch order
    [uc; c]) on two ranks: rank 0 of the pair evaluates the network on (x, uc), rank 1 on (x, c); one all-gather of the
    two denoised halves ([T, 4, h, w] fp32 each: 1.2 MB at V3D_512) per network evaluation puts [uc; c] on both ranks,
    which then apply the guidance and the sampler update redundantly.  25 small exchanges per image instead of the
    ~2600 of frame-sharding, and exactly the unsplit arithmetic (per-sample operators see the same samples)."""

    rank: int                         # 0 = uc half, 1 = c half
    group: Optional[object] = None
    exchanges: int = 0

    @classmethod
    def create(cls, group=None) -> "CfgSplit":
        if not dist.is_initialized() or dist.get_world_size(group) != 2:
            raise RuntimeError("CfgSplit needs a process group of exactly two ranks")
        return cls(dist.get_rank(group), group)

    def gather_halves(self, mine: torch.Tensor) -> torch.Tensor:
        """[n, ...] on each rank -> [2n, ...] = [rank 0's; rank 1's] on both."""
        self.exchanges += 1
        mine = mine.contiguous()
        out = mine.new_empty((2 * mine.shape[0],) + tuple(mine.shape[1:]))
        if mine.is_cuda and dist.get_backend(self.group) == "gloo":      # tests: two processes sharing one GPU
            h = torch.empty(out.shape, dtype=out.dtype)
            dist.all_gather_into_tensor(h, mine.cpu(), group=self.group)
            out.copy_(h)
        else:
            dist.all_gather_into_tensor(out, mine, group=self.group)
        return out


class CfgSplitGuider:
    """Wraps a classifier-free guider so that each rank of a CfgSplit pair feeds the network only its half of
    `prepare_inputs`' [uc; c] batch and the guidance sees the all-gathered pair."""

    def __init__(self, inner, split: CfgSplit):
        if not hasattr(inner, "scale"):
            raise RuntimeError("CfgSplit needs a classifier-free guider ([uc; c] batch); got " + type(inner).__name__)
        self.inner, self.split = inner, split

    def prepare_inputs(self, x, s, c, uc):
        x2, s2, c2 = self.inner.prepare_inputs(x, s, c, uc)
        n, r = x.shape[0], self.split.rank
        half = slice(r * n, (r + 1) * n)
        cond = {k: (v[half] if torch.is_tensor(v) and v.ndim > 0 and v.shape[0] == 2 * n else v) for k, v in c2.items()}
        return x2[half], s2[half], cond

    def __call__(self, x_half, sigma):
        return self.inner(self.split.gather_halves(x_half), sigma)


@dataclass
class ShardPlan:
    """How ONE image is spread over the ranks of a job: `sample` = frame blocks of the sampler/UNet (None = all frames),
    `cfg` = CFG-pair split (None = both halves on every rank), `decode` = frame blocks of the first-stage decode (always
    set: the decode has no CFG axis, so with a CFG split every pair divides its frames once more).

        views      world = P      sample = decode = P frame blocks
        cfg        world = 2      sample = None, cfg pair = the two ranks, decode = 2 frame blocks
        cfg+views  world = 2 P    rank g = 2 v + r: frame block v of P, CFG half r; decode = each block split in two
    """

    mode: str
    num_frames: int
    sample: Optional[ViewShard]
    cfg: Optional[CfgSplit]
    decode: ViewShard

    @classmethod
    def create(cls, num_frames: int, mode: str = "views") -> "ShardPlan":
        if not dist.is_initialized():
            raise RuntimeError("ShardPlan.create needs an initialised torch.distributed process group")
        rank, world = dist.get_rank(), dist.get_world_size()
        if mode == "views":
            vs = ViewShard.create(num_frames)
            return cls(mode, num_frames, vs, None, vs)
        if mode == "cfg":
            if world != 2:
                raise RuntimeError("mode 'cfg' runs on exactly 2 ranks (use 'cfg+views' for more)")
            return cls(mode, num_frames, None, CfgSplit.create(), ViewShard.create(num_frames))
        if mode == "cfg+views":
            if world < 4 or world % 2:
                raise RuntimeError("mode 'cfg+views' needs an even number (>= 4) of ranks")
            pv = world // 2
            v, r = divmod(rank, 2)
            # every rank creates every group, in the same order (torch.distributed requirement)
            pair_groups = [dist.new_group([2 * i, 2 * i + 1]) for i in range(pv)]
            view_groups = [dist.new_group([2 * i + j for i in range(pv)]) for j in range(2)]
            sample = ViewShard(num_frames, v, pv, view_groups[r], partition_frames(num_frames, pv))
            blocks = []
            for t0, n in sample.blocks:
                if n < 2:
                    raise ValueError(f"{num_frames} frames over {pv} blocks leave a block too small to split for the decode")
                blocks += [(t0, (n + 1) // 2), (t0 + (n + 1) // 2, n // 2)]
            return cls(mode, num_frames, sample, CfgSplit(r, pair_groups[v]), ViewShard(num_frames, rank, world, None, blocks))
        raise ValueError(f"unknown shard mode {mode!r} (views | cfg | cfg+views)")

    def gather_frames(self, local: torch.Tensor) -> torch.Tensor:
        return self.decode.gather_frames(local)

    def describe(self) -> Dict:
        return {"mode": self.mode, "sample_blocks": self.sample.blocks if self.sample else None,
                "cfg_rank": self.cfg.rank if self.cfg else None, "decode_blocks": self.decode.blocks,
                "exchanges": {**(self.sample.exchanges if self.sample else {}),
                              "cfg_gather": self.cfg.exchanges if self.cfg else 0,
                              **{"decode_" + k: v for k, v in self.decode.exchanges.items()
                                 if self.decode is not self.sample}}}
