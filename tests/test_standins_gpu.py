"""Contract test between the kernels and their CPU stand-ins (tests/emu_ops.py): every `v3d_b200.ops` entry point and
its torch restatement run on the SAME CUDA tensors and must agree to output rounding.  This is what lets the CPU host-
schedule tests (tests/test_host_schedule_cpu.py: parity vs oracle and vs the real reference's golden outputs, sharded ==
unsharded bit for bit) stand for the GPU path: the stand-ins are pinned to the kernels here, the schedules to the
stand-ins there.

Tolerance: bf16 outputs may differ by one rounding step where fp32 (kernel) and fp64 (stand-in) accumulation land on
different sides of a tie -> |a - b| <= 2^-7 |b| + 2^-7 max|b| * 1e-2; fp32 outputs 2e-3 relative to the tensor scale
(fast-math transcendentals: tanh.approx / ex2.approx).
"""
import os
import sys
from pathlib import Path

import pytest
import torch

pytestmark = [pytest.mark.gpu]

sys.path.insert(0, str(Path(__file__).resolve().parent))
DEV = "cuda"


def _pair():
    import emu_ops
    from v3d_b200 import ops

    return ops, emu_ops


def bf(*shape, seed=0, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(*shape, device=DEV, generator=g) * scale).to(torch.bfloat16)


def f32(*shape, seed=0, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return torch.randn(*shape, device=DEV, generator=g) * scale


def close(a, b, what):
    a, b = a.float(), b.float()
    scale = b.abs().max().clamp_min(1e-6)
    if a.dtype == torch.float32 and what.endswith("f32"):
        bad = (a - b).abs() > 2e-3 * scale
    else:
        bad = (a - b).abs() > (2.0 ** -7) * b.abs() + (2.0 ** -7) * 1e-2 * scale + 2e-3 * scale * (what.startswith("att"))
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.numel()} elements differ, max |d| {(a - b).abs().max().item():.4g}"


def both(fn_name, make_out, *args, **kw):
    ops, emu = _pair()
    o1, o2 = make_out(), make_out()
    getattr(ops, fn_name)(*[o1 if a is OUT else a for a in args], **{k: (o1 if v is OUT else v) for k, v in kw.items()})
    getattr(emu, fn_name)(*[o2 if a is OUT else a for a in args], **{k: (o2 if v is OUT else v) for k, v in kw.items()})
    torch.cuda.synchronize()
    return o1, o2


OUT = object()


@pytest.mark.parametrize("M,K,N", [(384, 320, 320), (1000, 64, 960), (256, 1280, 160)])
def test_gemm_linear_with_full_epilogue(M, K, N):
    a, w = bf(M, K, seed=1), bf(N, K, seed=2, scale=K ** -0.5)
    bias, fb = f32(N, seed=3), f32(M // 8 + 1, N, seed=4)
    r1, r2 = bf(M, N, seed=5), bf(M, N, seed=6)
    o1, o2 = both("gemm", lambda: torch.empty(M, N, device=DEV, dtype=torch.bfloat16), a, w, OUT, K=K, N=N,
                  rows_per_batch=M, bias=bias, fbias=fb, rows_per_frame=8, r1=r1, s1=0.75, r2=r2, s2=0.25, s0=0.5)
    close(o1, o2, "gemm linear + bias + fbias + R1 + R2")


@pytest.mark.parametrize("n_out", [160, 1280])
def test_gemm_geglu(n_out):
    ops, _ = _pair()
    M, K = 512, 320
    w = bf(2 * n_out, K, seed=2, scale=K ** -0.5)
    bias = f32(2 * n_out, seed=3)
    bn = ops.pick_block_n(2 * n_out, ops.ACT_GEGLU)
    perm = ops.geglu_perm(n_out, bn).to(DEV)
    o1, o2 = both("gemm", lambda: torch.empty(M, n_out, device=DEV, dtype=torch.bfloat16), bf(M, K, seed=1),
                  w[perm].contiguous(), OUT, K=K, N=2 * n_out, rows_per_batch=M, bias=bias[perm].contiguous(),
                  act=ops.ACT_GEGLU)
    close(o1, o2, "gemm geglu")


def test_gemm_conv3x3_temporal_and_halo():
    n, h, w_, c, co = 3, 16, 32, 128, 192
    o1, o2 = both("gemm", lambda: torch.empty(n * h * w_, co, device=DEV, dtype=torch.bfloat16), bf(n * h * w_, c, seed=1),
                  bf(co, 9 * c, seed=2, scale=(9 * c) ** -0.5), OUT, K=c, N=co, rows_per_batch=n * h * w_,
                  bias=f32(co, seed=3), conv=(n, h, w_))
    close(o1, o2, "gemm conv3x3")
    nb, T, hw, c = 2, 5, 64, 128
    a, wt, bias = bf(nb, T, hw, c, seed=4), bf(c, 3 * c, seed=5, scale=(3 * c) ** -0.5), f32(c, seed=6)
    o1, o2 = both("gemm", lambda: torch.empty(nb * T * hw, c, device=DEV, dtype=torch.bfloat16), a.view(-1, c), wt, OUT,
                  K=c, N=c, rows_per_batch=T * hw, batch=nb, a_batch_stride=T * hw * c, bias=bias, ntaps=3, tap_shift=hw)
    close(o1, o2, "gemm temporal")
    pad = bf(nb, T + 2, hw, c, seed=7)
    o1, o2 = both("gemm", lambda: torch.empty(nb * T * hw, c, device=DEV, dtype=torch.bfloat16), pad.view(-1, c), wt, OUT,
                  K=c, N=c, rows_per_batch=T * hw, batch=nb, a_batch_stride=(T + 2) * hw * c, bias=bias, ntaps=3,
                  tap_shift=hw, a_rows=(T + 2) * hw, a_row0=hw)
    close(o1, o2, "gemm temporal, halo'd operand")
    o1, o2 = both("gemm", lambda: torch.empty(n * h * w_, 16, device=DEV, dtype=torch.float32), bf(n * h * w_, c, seed=8),
                  bf(16, 9 * c, seed=9, scale=(9 * c) ** -0.5), OUT, K=c, N=16, rows_per_batch=n * h * w_,
                  bias=f32(16, seed=10), conv=(n, h, w_))
    close(o1, o2, "gemm conv3x3 f32")


@pytest.mark.parametrize("silu", [False, True])
def test_groupnorm_and_layernorm(silu):
    ops, emu = _pair()
    ns, rows, c = 3, 200, 320
    x = bf(ns * rows, c, seed=1, scale=2.0) + 0.5
    s1 = torch.zeros(ns, 32, 2, device=DEV, dtype=torch.float64)
    s2 = torch.zeros_like(s1)
    ops.groupnorm_stats(x, s1, rows, ns, c)
    emu.groupnorm_stats(x, s2, rows, ns, c)
    torch.cuda.synchronize()
    assert torch.allclose(s1, s2, rtol=1e-5, atol=1e-3)          # fp32 partial sums in the kernel
    g, b = f32(c, seed=2) + 1.0, f32(c, seed=3)
    o1, o2 = both("groupnorm_apply", lambda: torch.empty(ns * rows, c, device=DEV, dtype=torch.bfloat16), x, OUT, s2, g,
                  b, rows, ns, c, 1e-5, silu)
    close(o1, o2, "groupnorm apply")
    y1 = torch.empty(ns * rows, c, device=DEV, dtype=torch.bfloat16)
    y2 = torch.empty_like(y1)
    ops.groupnorm(x, y1, g, b, rows, ns, c, 1e-5, silu, ops.groupnorm_workspace(DEV))
    emu.groupnorm(x, y2, g, b, rows, ns, c, 1e-5, silu, None)
    torch.cuda.synchronize()
    close(y1, y2, "groupnorm (one launch)")
    add = f32(ns, c, seed=4)
    ys1 = torch.empty(ns * rows, c, device=DEV, dtype=torch.bfloat16)
    ys2 = torch.empty_like(ys1)
    y1, y2 = torch.empty_like(ys1), torch.empty_like(ys1)
    ops.layernorm(x, y1, g, b, ns * rows, c, add=add, ysum=ys1, rows_per_frame=rows)
    emu.layernorm(x, y2, g, b, ns * rows, c, add=add, ysum=ys2, rows_per_frame=rows)
    torch.cuda.synchronize()
    close(ys1, ys2, "layernorm fused sum")
    close(y1, y2, "layernorm")


def test_attention_family():
    from v3d_b200.viewshard import ViewShard

    nb, ntok, heads = 2, 320, 3
    c = heads * 64
    qkv = bf(nb * ntok, 3 * c, seed=1)
    o1, o2 = both("attention_spatial", lambda: torch.empty(nb * ntok, c, device=DEV, dtype=torch.bfloat16), qkv, OUT, nb,
                  ntok, heads, 0.125)
    close(o1, o2, "attention spatial")
    nb, T, S = 2, 7, 48
    qkv = bf(nb * T * S, 3 * c, seed=2)
    o1, o2 = both("attention_temporal", lambda: torch.empty(nb * T * S, c, device=DEV, dtype=torch.bfloat16), qkv, OUT, nb,
                  T, S, heads, 0.125)
    close(o1, o2, "attention temporal")
    vs = ViewShard(num_frames=T, rank=1, world=3)                       # blocks 3 / 2 / 2, this rank holds frames 3..4
    tl, tmax = vs.tl, vs.tmax
    ql = bf(nb * tl * S, 3 * c, seed=3)
    kvbuf = bf(3 * nb * tmax * S, 2 * c, seed=4)
    row, bstride = vs.kv_table(nb, S)
    o1, o2 = both("attention_temporal_kv", lambda: torch.empty(nb * tl * S, c, device=DEV, dtype=torch.bfloat16), ql,
                  kvbuf, OUT, nb, tl, S, heads, row, bstride, 0.125)
    close(o1, o2, "attention temporal, split K|V")


def test_data_movement_and_small_matrices():
    ops, emu = _pair()
    n, h, w_, c = 2, 6, 10, 64
    x = bf(n * h * w_, c, seed=1)
    o1, o2 = both("upsample_nearest2x", lambda: torch.empty(n * 4 * h * w_, c, device=DEV, dtype=torch.bfloat16), x, OUT,
                  n, h, w_, c)
    assert torch.equal(o1, o2)
    d1 = torch.zeros(n * h * w_, 192, device=DEV, dtype=torch.bfloat16)
    d2 = torch.zeros_like(d1)
    ops.copy_channels(x, c, d1[:, 128:], 192, n * h * w_, c)
    emu.copy_channels(x, c, d2[:, 128:], 192, n * h * w_, c)
    assert torch.equal(d1, d2)
    for cc, stride, pad, ho, wo in ((8, 1, 1, h, w_), (64, 2, 1, 3, 5), (64, 2, 0, 3, 5)):
        xi = bf(n * h * w_, cc, seed=2)
        kpad = (9 * cc + 63) // 64 * 64
        o1, o2 = both("im2col3x3", lambda: torch.empty(n * ho * wo, kpad, device=DEV, dtype=torch.bfloat16), xi, OUT, n, h,
                      w_, cc, stride, pad, ho, wo, kpad)
        assert torch.equal(o1, o2), (cc, stride, pad)
    xf = f32(n, 8, h, w_, seed=3)
    o1, o2 = both("nchw_f32_to_nhwc_bf16", lambda: torch.empty(n * h * w_, 8, device=DEV, dtype=torch.bfloat16), xf, OUT)
    assert torch.equal(o1, o2)
    o1, o2 = both("nhwc_to_nchw_f32", lambda: torch.empty(n, 48, h, w_, device=DEV), x, OUT, n, 48, h * w_, c)
    assert torch.equal(o1, o2)
    xs, wsm, bsm = f32(36, 320, seed=4), bf(1280, 320, seed=5, scale=320 ** -0.5), f32(1280, seed=6)
    o1, o2 = both("small_linear", lambda: torch.zeros(36, 1280, device=DEV), xs, wsm, bsm, OUT, act_in=ops.ACT_SILU)
    close(o1, o2, "small_linear f32")
    t = torch.linspace(-2.0, 700.0, 36, device=DEV)
    o1, o2 = both("timestep_embedding", lambda: torch.empty(36, 320, device=DEV), t, OUT, 320)
    close(o1, o2, "timestep_embedding f32")
    xo = f32(2 * 5 * 40, 16, seed=7)
    o1, o2 = both("time_mix_conv", lambda: torch.empty(2 * 5, 3, 5, 8, device=DEV), xo, 16, f32(3, 3, 3, seed=8),
                  f32(3, seed=9), OUT, 2, 5, 40, 3)
    close(o1, o2, "time_mix_conv f32")


def test_sampler_arithmetic():
    ops, emu = _pair()
    n, per = 6, 4 * 8 * 8
    x, den, den2 = f32(n, 4, 8, 8, seed=1), f32(n, 4, 8, 8, seed=2), f32(n, 4, 8, 8, seed=3)
    sig = torch.rand(n, device=DEV) * 50 + 0.5
    nxt = sig * 0.6
    nxt[-1] = 0.0
    for name, args in (("edm_denoise_combine", (den, x, sig, OUT, n, per)), ("euler_step", (x, den, sig, nxt, OUT, n, per)),
                       ("heun_step", (x, den, den2, den, sig, nxt, OUT, n, per))):
        o1, o2 = both(name, lambda: torch.empty_like(x), *args)
        close(o1, o2, name + " f32")
    cn1, cn2 = torch.empty(n, device=DEV), torch.empty(n, device=DEV)
    y1, y2 = torch.empty_like(x), torch.empty_like(x)
    ops.edm_scale_input(x, sig, y1, cn1, n, per)
    emu.edm_scale_input(x, sig, y2, cn2, n, per)
    close(y1, y2, "edm_scale_input f32")
    close(cn1, cn2, "c_noise f32")
    scale = torch.linspace(1.5, 3.5, 3, device=DEV)
    o1, o2 = both("cfg_combine", lambda: torch.empty(3, 4, 8, 8, device=DEV), x, scale, OUT, 1, 3, per)
    close(o1, o2, "cfg_combine f32")
    img = f32(3, 3, 16, 16, seed=4)
    o1, o2 = both("frames_nchw_to_u8", lambda: torch.empty(3, 16, 16, 3, device=DEV, dtype=torch.uint8), img, OUT)
    assert (o1.int() - o2.int()).abs().max() <= 1
