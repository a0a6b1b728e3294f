"""TEST INFRASTRUCTURE: a torch-on-CPU stand-in for the `v3d_b200.ops` API (the C-ABI wrappers).

The product path has no CPU implementation and must not get one; this module exists so that the HOST SCHEDULES of
the drop-in modules (`VideoUNet._run`, `VideoDecoder._run`: which kernel runs on which buffer with which strides,
offsets, fused-epilogue operands, halo buffers and exchanges) can be executed and checked on a box without a GPU:
  * against the oracle (tests/test_host_schedule_cpu.py): a wrong leading dimension, bias row, residual, packing
    permutation or layout convention in the schedule shows up as a parity failure;
  * frame-sharded against unsharded over gloo (same test file).
Each function restates the CONTRACT documented in include/v3d_b200.h for one entry point, in plain torch: bf16 storage,
arithmetic in fp64 (so that results do not depend on how a CPU BLAS blocks a given shape: a frame-sharded run and an
unsharded run then agree to the last bit wherever the schedule is right), one rounding at the store.  It is only ever
installed by `patched()` inside tests; nothing under v3d_b200/ imports it.  The functions are device-agnostic, so
tests/test_standins_gpu.py can hold each of them against the real kernel on the same CUDA tensors.
"""
from __future__ import annotations

import contextlib
import math

import torch
import torch.nn.functional as F

from v3d_b200 import ops as real_ops

ACT_NONE, ACT_SILU, ACT_GEGLU = 0, 1, 2
BF = torch.bfloat16
_launches = [0]


def _strided(t: torch.Tensor, size, stride) -> torch.Tensor:
    """Typed window into t's storage starting at t's first element (what a kernel sees through a raw pointer)."""
    return torch.as_strided(t, size, stride, t.storage_offset())


def _store(dst: torch.Tensor, val: torch.Tensor) -> None:
    dst.copy_(val.to(dst.dtype))


def launch_count() -> int:
    return _launches[0]


def _tick(n: int = 1) -> None:
    _launches[0] += n


pick_block_n = real_ops.pick_block_n      # host functions of the library (no GPU involved)
geglu_perm = real_ops.geglu_perm


# ---------------------------------------------------------------------------------------------------------------
# v3d_gemm_bf16
# ---------------------------------------------------------------------------------------------------------------
def gemm(a, w, out, *, K, N, rows_per_batch, batch=1, lda=None, ldb=None, ldd=None, a_batch_stride=0,
         b_batch_stride=0, bias=None, fbias=None, rows_per_frame=1, ldfb=0, r1=None, ldr1=0, r2=None, ldr2=0, s0=1.0,
         s1=1.0, s2=1.0, act=ACT_NONE, ntaps=1, tap_shift=0, conv=None, block_n=0, transposed=False, valid_cols=0,
         accumulate=False, a_rows=0, a_row0=0):
    assert a.dtype == BF and w.dtype == BF and not transposed and not accumulate
    assert K % 64 == 0 and N % 16 == 0, (K, N)
    _tick()
    taps = 9 if conv is not None else ntaps
    lda = K if lda is None else lda
    ldb = taps * K if ldb is None else ldb
    n_out = N // 2 if act == ACT_GEGLU else N
    ldd = n_out if ldd is None else ldd
    if conv is not None:
        n, h, wd = conv
        rows = n * h * wd
        x = _strided(a, (n, h, wd, K), (h * wd * lda, wd * lda, lda, 1)).double().permute(0, 3, 1, 2)
        wt = _strided(w, (N, 9 * K), (ldb, 1)).double().reshape(N, 3, 3, K).permute(0, 3, 1, 2)   # (ky, kx, ci)
        acc = F.conv2d(x, wt, padding=1).permute(0, 2, 3, 1).reshape(rows, N)
    else:
        rows = batch * rows_per_batch
        arows = a_rows if a_rows > 0 else rows_per_batch + a_row0
        abs_ = a_batch_stride if batch > 1 else arows * lda
        A = _strided(a, (batch, arows, K), (abs_, lda, 1)).double()
        nbb = batch if b_batch_stride else 1
        Bm = _strided(w, (nbb, N, taps * K), (b_batch_stride, ldb, 1)).double()
        acc = torch.zeros(batch, rows_per_batch, N, dtype=torch.float64, device=a.device)
        for tap in range(taps):
            first = a_row0 + (tap - taps // 2) * tap_shift           # A row read by output row 0
            lo, hi = max(0, -first), min(rows_per_batch, arows - first)
            if hi <= lo:
                continue
            At = torch.zeros(batch, rows_per_batch, K, dtype=torch.float64, device=a.device)
            At[:, lo:hi] = A[:, first + lo: first + hi]               # rows outside [0, a_rows) read as zero
            acc += At @ Bm[:, :, tap * K:(tap + 1) * K].transpose(1, 2)
        acc = acc.reshape(rows, N)
    if bias is not None:
        acc = acc + bias.double()[:N]
    if fbias is not None:
        ldfb_ = ldfb if ldfb > 0 else N
        frames = (rows + rows_per_frame - 1) // rows_per_frame
        fb = _strided(fbias, (frames, N), (ldfb_, 1)).double()
        acc = acc + fb.repeat_interleave(rows_per_frame, dim=0)[:rows]
    if act == ACT_GEGLU:
        bn = block_n or pick_block_n(N, act)
        half = bn // 2
        t = acc.reshape(rows, N // bn, bn)
        acc = (t[..., :half] * F.gelu(t[..., half:])).reshape(rows, n_out)       # value * gelu_erf(gate)
    elif act == ACT_SILU:
        acc = F.silu(acc)
    res = s0 * acc
    if r1 is not None:
        res = res + s1 * _strided(r1, (rows, n_out), (ldr1 or n_out, 1)).double()
    if r2 is not None:
        res = res + s2 * _strided(r2, (rows, n_out), (ldr2 or n_out, 1)).double()
    _store(_strided(out, (rows, n_out), (ldd, 1)), res)
    return out


# ---------------------------------------------------------------------------------------------------------------
# norms
# ---------------------------------------------------------------------------------------------------------------
def groupnorm_stats(x, stats, rows_per_sample, nsamples, c, ldx=None, groups=32, pre_zeroed=False):
    assert x.dtype == BF and stats.dtype == torch.float64
    _tick()
    ldx = c if ldx is None else ldx
    v = _strided(x, (nsamples, rows_per_sample, groups, c // groups), (rows_per_sample * ldx, ldx, c // groups, 1))
    v = v.double()
    if not pre_zeroed:
        stats.zero_()
    st = stats.view(nsamples, groups, 2)
    st[..., 0] += v.sum(dim=(1, 3))
    st[..., 1] += (v * v).sum(dim=(1, 3))
    return stats


def groupnorm_workspace(device):
    return torch.zeros(8, dtype=torch.uint8)


def groupnorm(x, y, gamma, beta, rows_per_sample, nsamples, c, eps, silu, workspace, ldx=None, groups=32):
    """single-launch GroupNorm = statistics + apply (one tick: one launch)"""
    st = torch.zeros(nsamples, groups, 2, dtype=torch.float64, device=x.device)
    groupnorm_stats(x, st, rows_per_sample, nsamples, c, ldx=ldx, groups=groups, pre_zeroed=True)
    groupnorm_apply(x, y, st, gamma, beta, rows_per_sample, nsamples, c, eps, silu, ldx=ldx, groups=groups)
    _tick(-1)
    return y


def groupnorm_apply(x, y, stats, gamma, beta, rows_per_sample, nsamples, c, eps, silu, ldx=None, groups=32):
    _tick()
    ldx = c if ldx is None else ldx
    cpg = c // groups
    n = rows_per_sample * cpg                                       # the kernel divides by the LOCAL element count
    st = stats.reshape(-1)[: nsamples * groups * 2].view(nsamples, groups, 2)
    mean = st[..., 0] / n
    rstd = (st[..., 1] / n - mean * mean + eps).rsqrt()
    v = _strided(x, (nsamples, rows_per_sample, groups, cpg), (rows_per_sample * ldx, ldx, cpg, 1)).double()
    v = (v - mean.double()[:, None, :, None]) * rstd.double()[:, None, :, None]
    v = v.reshape(nsamples, rows_per_sample, c) * gamma.double() + beta.double()
    if silu:
        v = F.silu(v)
    _store(_strided(y, (nsamples * rows_per_sample, c), (c, 1)), v.reshape(-1, c))
    return y


def layernorm(x, y, gamma, beta, rows, c, eps=1e-5, add=None, ysum=None, rows_per_frame=1):
    _tick()
    v = _strided(x, (rows, c), (c, 1)).double()
    if add is not None:
        v = v + add.double().reshape(-1, c).repeat_interleave(rows_per_frame, dim=0)[:rows]
        if ysum is not None:
            v = v.to(BF).double()                                   # rounded once: residual stream == LN input
            _store(_strided(ysum, (rows, c), (c, 1)), v)
    _store(_strided(y, (rows, c), (c, 1)), F.layer_norm(v, (c,), gamma.double(), beta.double(), eps))
    return y


def softmax_rows_f32(x, y, rows, n, scale=1.0):
    _tick()
    _store(y.view(rows, n), torch.softmax(x.view(rows, n).double() * scale, dim=-1))
    return y


# ---------------------------------------------------------------------------------------------------------------
# attention (head dim 64)
# ---------------------------------------------------------------------------------------------------------------
def _heads(t, n, tok, nheads):
    return t.double().reshape(n, tok, nheads, 64).transpose(1, 2)


def attention_spatial(qkv, out, nbatch, ntok, nheads, scale):
    _tick()
    c, ld = nheads * 64, qkv.stride(0)
    q, k, v = (_strided(qkv[:, i * c:], (nbatch * ntok, c), (ld, 1)) for i in range(3))
    o = F.scaled_dot_product_attention(_heads(q, nbatch, ntok, nheads), _heads(k, nbatch, ntok, nheads),
                                       _heads(v, nbatch, ntok, nheads), scale=scale)
    _store(_strided(out, (nbatch * ntok, c), (out.stride(0), 1)), o.transpose(1, 2).reshape(nbatch * ntok, c))
    return out


def attention_temporal(qkv, out, nb, t, s, nheads, scale):
    _tick()
    c, ld = nheads * 64, qkv.stride(0)
    rows = nb * t * s

    def seq(i):  # rows (b, t, s) -> sequences (b, s) of t tokens
        m = _strided(qkv[:, i * c:], (rows, c), (ld, 1)).double().reshape(nb, t, s, nheads, 64)
        return m.permute(0, 2, 3, 1, 4).reshape(nb * s, nheads, t, 64)

    o = F.scaled_dot_product_attention(seq(0), seq(1), seq(2), scale=scale)
    o = o.reshape(nb, s, nheads, t, 64).permute(0, 3, 1, 2, 4).reshape(rows, c)
    _store(_strided(out, (rows, c), (out.stride(0), 1)), o)
    return out


def attention_temporal_kv(q, kv, out, nb, tq, s, nheads, kv_row, kv_bstride, scale):
    _tick()
    c = nheads * 64
    tk = len(kv_row)
    assert 0 < tq <= tk <= 32
    rows = nb * tq * s
    qm = _strided(q, (rows, c), (q.stride(0), 1)).double().reshape(nb, tq, s, nheads, 64)
    qm = qm.permute(0, 2, 3, 1, 4)                                                        # [b, s, h, tq, 64]
    idx = torch.tensor([[[kv_row[f] + b * kv_bstride[f] + p for f in range(tk)] for p in range(s)]
                        for b in range(nb)], device=kv.device)                             # [b, s, tk] buffer rows
    kvm = kv.double()[idx]                                                                  # [b, s, tk, 2c]
    km = kvm[..., :c].reshape(nb, s, tk, nheads, 64).permute(0, 1, 3, 2, 4)
    vm = kvm[..., c:].reshape(nb, s, tk, nheads, 64).permute(0, 1, 3, 2, 4)
    o = F.scaled_dot_product_attention(qm, km, vm, scale=scale)                            # [b, s, h, tq, 64]
    _store(_strided(out, (rows, c), (out.stride(0), 1)), o.permute(0, 3, 1, 2, 4).reshape(rows, c))
    return out


# ---------------------------------------------------------------------------------------------------------------
# data movement / small matrices
# ---------------------------------------------------------------------------------------------------------------
def upsample_nearest2x(x, y, n, h, w, c):
    _tick()
    v = x.view(n, h, w, c)
    _store(y.view(n, 2 * h, 2 * w, c), v.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2))
    return y


def copy_channels(src, ld_src, dst, ld_dst, rows, ncols):
    _tick()
    _strided(dst, (rows, ncols), (ld_dst, 1)).copy_(_strided(src, (rows, ncols), (ld_src, 1)))


def im2col3x3(x, y, n, h, w, c, stride, pad, hout, wout, kpad):
    _tick()
    v = x.view(n, h, w, c).double().permute(0, 3, 1, 2)
    v = F.pad(v, (1, 1, 1, 1)) if pad else F.pad(v, (0, 1, 0, 1))
    cols = F.unfold(v, kernel_size=3, stride=stride)                                       # [n, c*9, L], (c, ky, kx)
    cols = cols.reshape(n, c, 9, hout * wout).permute(0, 3, 2, 1).reshape(n * hout * wout, 9 * c)   # (tap, c)
    full = torch.zeros(n * hout * wout, kpad, dtype=torch.float64, device=x.device)
    full[:, : 9 * c] = cols
    _store(y.view(n * hout * wout, kpad), full)
    return y


def nchw_f32_to_nhwc_bf16(x, y, scale=1.0):
    _tick()
    n, c, h, w = x.shape
    _store(y.view(n, h, w, c), x.double().permute(0, 2, 3, 1) * scale)
    return y


def nhwc_to_nchw_f32(x, y, n, c, hw, ldx, scale=1.0):
    _tick()
    v = _strided(x, (n, hw, c), (hw * ldx, ldx, 1)).double() * scale
    y.view(n, c, hw).copy_(v.transpose(1, 2))
    return y


def _act(v, kind):
    return F.silu(v) if kind == ACT_SILU else v


def small_linear(x, w, bias, y, *, act_in=ACT_NONE, act_out=ACT_NONE, accumulate=False):
    _tick()
    xin = _act(x.double(), act_in)
    if x.shape[1] % 64 == 0 and act_out == ACT_NONE:
        xin = xin.to(BF).double()                                    # tensor-core path: operand converted to bf16
    r = xin @ w.double().t()
    if bias is not None:
        r = r + bias.double()
    r = _act(r, act_out)
    if accumulate:
        y += r
    else:
        y.copy_(r)
    return y


def timestep_embedding(t, out, dim, max_period=10000.0):
    _tick()
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t.float()[:, None] * freqs[None]                         # fp32 like the kernel and the reference
    out.copy_(torch.cat([torch.cos(args), torch.sin(args)], dim=-1))
    return out


def time_mix_conv(x, ldx, w, bias, y, nb, t, hw, c):
    _tick()
    v = _strided(x, (nb, t, hw, c), (t * hw * ldx, hw * ldx, ldx, 1)).double().permute(0, 3, 1, 2)   # [b, c, t, hw]
    o = F.conv3d(v.unsqueeze(-1), w.double().reshape(c, c, 3, 1, 1), bias.double(), padding=(1, 0, 0)).squeeze(-1)
    y.view(nb, t, c, hw).copy_(o.permute(0, 2, 1, 3))
    return y


# ---------------------------------------------------------------------------------------------------------------
# sampler arithmetic (fp32, elementwise; sampler.cu)
# ---------------------------------------------------------------------------------------------------------------
def _per_sample(v, like):
    return v.float().reshape(-1, *([1] * (like.dim() - 1)))


def edm_scale_input(x, sigma, y, c_noise, nsamples, per_sample):
    _tick()
    s = _per_sample(sigma, x)
    y.copy_(x * (1.0 / torch.sqrt(s * s + 1.0)))
    if c_noise is not None:
        c_noise.copy_(0.25 * torch.log(sigma.float()).reshape(c_noise.shape))
    return y


def edm_denoise_combine(net, x, sigma, out, nsamples, per_sample):
    _tick()
    s = _per_sample(sigma, x)
    d = s * s + 1.0
    out.copy_(net * (-s / torch.sqrt(d)) + x * (1.0 / d))
    return out


def cfg_combine(den, scale, out, b, t, per_sample):
    _tick()
    half = b * t
    xu, xc = den[:half], den[half:]
    sc = scale.float().reshape(-1)[:t].repeat(b).reshape(half, *([1] * (den.dim() - 1)))
    out.copy_(xu + sc * (xc - xu))
    return out


def euler_step(x, den, sigma_hat, sigma_next, out, nsamples, per_sample):
    _tick()
    sh, sn = _per_sample(sigma_hat, x), _per_sample(sigma_next, x)
    out.copy_(x + (sn - sh) * ((x - den) / sh))
    return out


def heun_step(x, den, x_euler, den2, sigma_hat, sigma_next, out, nsamples, per_sample):
    _tick()
    sh, sn = _per_sample(sigma_hat, x), _per_sample(sigma_next, x)
    d = (x - den) / sh
    d2 = (x_euler - den2) / sn.clamp_min(1e-30)
    out.copy_(torch.where(sn > 0, x + ((d + d2) / 2.0) * (sn - sh), x_euler))
    return out


def frames_nchw_to_u8(x, y):
    _tick()
    v = torch.clamp((x.float() + 1.0) / 2.0, 0.0, 1.0) * 255.0
    y.copy_(v.permute(0, 2, 3, 1).to(torch.uint8))                  # truncating cast, like numpy astype
    return y


_EMULATED = ["launch_count", "gemm", "groupnorm_stats", "groupnorm_apply", "groupnorm", "groupnorm_workspace", "layernorm", "softmax_rows_f32",
             "attention_spatial", "attention_temporal", "attention_temporal_kv", "upsample_nearest2x", "copy_channels",
             "im2col3x3", "nchw_f32_to_nhwc_bf16", "nhwc_to_nchw_f32", "small_linear", "timestep_embedding",
             "time_mix_conv", "edm_scale_input", "edm_denoise_combine", "cfg_combine", "euler_step", "heun_step",
             "frames_nchw_to_u8"]


@contextlib.contextmanager
def patched():
    """Install the stand-ins on v3d_b200.ops for the duration of a test (every other op keeps its CUDA-only body)."""
    saved = {n: getattr(real_ops, n) for n in _EMULATED}
    try:
        for n in _EMULATED:
            setattr(real_ops, n, globals()[n])
        yield
    finally:
        for n, f in saved.items():
            setattr(real_ops, n, f)
