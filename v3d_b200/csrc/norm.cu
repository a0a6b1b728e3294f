// Memory-bound normalisation kernels on NHWC / token-major bf16 activations (fp32 statistics):
// GroupNorm(32) statistics + apply(+SiLU), LayerNorm (+ fused per-frame add), row softmax.
// All global accesses are 16-byte vectors along the contiguous channel dimension.
#include <cstdlib>

#include "common.cuh"
#include "host_util.cuh"
#include "v3d_b200.h"

namespace v3d {

// ------------------------------------------------------------------------------------------------
// GroupNorm statistics: stats[sample][group] = {sum, sumsq} (double), accumulated with atomics.
// grid = (row chunks, samples); each thread owns one 8-channel vector column and strides over rows.
// ------------------------------------------------------------------------------------------------
constexpr int kGnThreads = 256;
constexpr int kGnUnroll = 8;  // independent 16-byte loads in flight per thread

__global__ void __launch_bounds__(512)
gn_stats_kernel(const bf16* __restrict__ x, double* __restrict__ stats, long long rows_per_sample,
                int C, long long ldx, int groups, int rows_per_cta) {
  extern __shared__ float s_acc[];  // [2][C]
  const int vpr = C >> 3;           // 16-byte vectors per row
  const int sample = blockIdx.y;
  const long long row0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  long long row1 = row0 + rows_per_cta;
  if (row1 > rows_per_sample) row1 = rows_per_sample;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();

  const int lanes_r = blockDim.x / vpr;  // row lanes (block = vpr * lanes_r threads)
  const int vc = threadIdx.x % vpr;
  const int rl = threadIdx.x / vpr;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  const bf16* base = x + (static_cast<long long>(sample) * rows_per_sample) * ldx + vc * 8;
  long long r = row0 + rl;
  for (; r + static_cast<long long>(kGnUnroll - 1) * lanes_r < row1; r += static_cast<long long>(kGnUnroll) * lanes_r) {
    uint4 u[kGnUnroll];
#pragma unroll
    for (int k = 0; k < kGnUnroll; ++k)
      u[k] = __ldg(reinterpret_cast<const uint4*>(base + (r + static_cast<long long>(k) * lanes_r) * ldx));
#pragma unroll
    for (int k = 0; k < kGnUnroll; ++k) {
      const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(w[j]);
        s[2 * j] += f.x; q[2 * j] = fmaf(f.x, f.x, q[2 * j]);
        s[2 * j + 1] += f.y; q[2 * j + 1] = fmaf(f.y, f.y, q[2 * j + 1]);
      }
    }
  }
  for (; r < row1; r += lanes_r) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + r * ldx));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      s[2 * j] += f.x; q[2 * j] = fmaf(f.x, f.x, q[2 * j]);
      s[2 * j + 1] += f.y; q[2 * j + 1] = fmaf(f.y, f.y, q[2 * j + 1]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&s_acc[vc * 8 + j], s[j]);
    atomicAdd(&s_acc[C + vc * 8 + j], q[j]);
  }
  __syncthreads();
  const int cg = C / groups;
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    double ds = 0.0, dq = 0.0;
    for (int c = 0; c < cg; ++c) {
      ds += static_cast<double>(s_acc[g * cg + c]);
      dq += static_cast<double>(s_acc[C + g * cg + c]);
    }
    double* o = stats + (static_cast<long long>(sample) * groups + g) * 2;
    atomicAdd(o, ds);
    atomicAdd(o + 1, dq);
  }
}

// y = act((x - mean) * rstd * gamma + beta). Each thread owns one 8-channel vector column: its scale/shift
// (16 floats) are computed once from the statistics and kept in registers; no smem, no div/mod in the loop.
__global__ void __launch_bounds__(512)
gn_apply_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, const double* __restrict__ stats,
                const float* __restrict__ gamma, const float* __restrict__ beta,
                long long rows_per_sample, int C, long long ldx, int groups, float eps, int silu,
                int rows_per_cta) {
  const int vpr = C >> 3;
  const int sample = blockIdx.y;
  const int lanes_r = blockDim.x / vpr;
  const int vc = threadIdx.x % vpr;
  const int rl = threadIdx.x / vpr;
  const int cg = C / groups;
  const double cnt = static_cast<double>(rows_per_sample) * cg;
  float sc[8], sh[8];
  {
    int g_prev = -1;
    float mean_f = 0.f, rstd_f = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = vc * 8 + j;
      const int g = c / cg;
      if (g != g_prev) {
        const double* st = stats + (static_cast<long long>(sample) * groups + g) * 2;
        const double mean = st[0] / cnt;
        double var = st[1] / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        mean_f = static_cast<float>(mean);
        rstd_f = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
        g_prev = g;
      }
      const float a = rstd_f * __ldg(gamma + c);
      sc[j] = a;
      sh[j] = __ldg(beta + c) - mean_f * a;
    }
  }
  const long long row0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  long long row1 = row0 + rows_per_cta;
  if (row1 > rows_per_sample) row1 = rows_per_sample;
  const long long srow = static_cast<long long>(sample) * rows_per_sample;
  const bf16* xb = x + srow * ldx + vc * 8;
  bf16* yb = y + srow * static_cast<long long>(C) + vc * 8;

  auto emit = [&](const uint4& u, long long r) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      float a = fmaf(f.x, sc[2 * j], sh[2 * j]);
      float b = fmaf(f.y, sc[2 * j + 1], sh[2 * j + 1]);
      if (silu) {
        a = silu_f(a);
        b = silu_f(b);
      }
      o[j] = pack_bf16x2(a, b);
    }
    *reinterpret_cast<uint4*>(yb + r * static_cast<long long>(C)) = make_uint4(o[0], o[1], o[2], o[3]);
  };
  long long r = row0 + rl;
  for (; r + static_cast<long long>(kGnUnroll - 1) * lanes_r < row1; r += static_cast<long long>(kGnUnroll) * lanes_r) {
    uint4 u[kGnUnroll];
#pragma unroll
    for (int k = 0; k < kGnUnroll; ++k)
      u[k] = __ldg(reinterpret_cast<const uint4*>(xb + (r + static_cast<long long>(k) * lanes_r) * ldx));
#pragma unroll
    for (int k = 0; k < kGnUnroll; ++k) emit(u[k], r + static_cast<long long>(k) * lanes_r);
  }
  for (; r < row1; r += lanes_r) emit(__ldg(reinterpret_cast<const uint4*>(xb + r * ldx)), r);
}

// ------------------------------------------------------------------------------------------------
// GroupNorm in ONE launch: statistics, a grid-wide barrier, then the normalisation of the same rows.
//   phase 1  every thread sums (x, x^2) of its 8-channel column over its rows; the CTA folds its threads' partials per
//            group through shared memory + a warp-shuffle tree and writes ONE fp32 (sum, sumsq) pair per group;
//   barrier  sense-reversing counter in the workspace; the grid is sized to be co-resident (occupancy x SMs);
//   phase 2  every CTA adds the partials of its sample in fp64, chunk order fixed by the launch geometry (warp-shuffle
//            tree again) -> mean / rstd, then re-reads its rows in REVERSE order (the rows read last are the ones most
//            likely still in the 126 MB L2) and writes act((x - mean) rstd gamma + beta).
// No atomics on data: the result is bit-reproducible run to run.  One launch and one HBM read less than the
// stats + apply pair, which stays for the frame-sharded path (statistics all-reduce between the two).
// ------------------------------------------------------------------------------------------------
struct GnBarrier {
  unsigned int count;
  unsigned int gen;
};
constexpr int kGnMaxCtas = 4096;                                   // partial slots in the workspace
constexpr size_t kGnWsBytes = 256 + sizeof(float) * 2 * 32 * kGnMaxCtas;

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// MAXT / MINB: launch bounds.  <256, 3> (80 registers) is what every width up to C = 2048 uses: three 240-256-thread
// CTAs per SM instead of two keep enough loads in flight for the streaming phases; wider rows take <512, 1>.
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
gn_fused_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, const float* __restrict__ gamma,
                const float* __restrict__ beta, GnBarrier* __restrict__ bar, float2* __restrict__ partial,
                long long rows_per_sample, int C, long long ldx, int groups, float eps, int silu, int rows_per_cta) {
  extern __shared__ float s_red[];  // [2][lanes_r][C] thread partials, reused as [groups][2] mean / rstd
  const int vpr = C >> 3;
  const int sample = blockIdx.y;
  const int chunks = gridDim.x;
  const int lanes_r = blockDim.x / vpr;
  const int vc = threadIdx.x % vpr;
  const int rl = threadIdx.x / vpr;
  const int cg = C / groups;
  // only FULL warps take part in the shuffle trees (240-thread CTAs end in a 16-lane warp: a full-mask shuffle there
  // would read lanes that do not exist); a trailing partial warp gets warp id >= nwarps and skips those loops
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const long long row0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  long long row1 = row0 + rows_per_cta;
  if (row1 > rows_per_sample) row1 = rows_per_sample;
  const long long srow = static_cast<long long>(sample) * rows_per_sample;
  const bf16* xb = x + srow * ldx + vc * 8;

  // ---- phase 1: thread partials
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  auto accum = [&](const uint4& u) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      s[2 * j] += f.x; q[2 * j] = fmaf(f.x, f.x, q[2 * j]);
      s[2 * j + 1] += f.y; q[2 * j + 1] = fmaf(f.y, f.y, q[2 * j + 1]);
    }
  };
  long long r = row0 + rl;
  for (; r + static_cast<long long>(kGnUnroll - 1) * lanes_r < row1; r += static_cast<long long>(kGnUnroll) * lanes_r) {
    uint4 u[kGnUnroll];
#pragma unroll
    for (int k = 0; k < kGnUnroll; ++k)
      u[k] = __ldg(reinterpret_cast<const uint4*>(xb + (r + static_cast<long long>(k) * lanes_r) * ldx));
#pragma unroll
    for (int k = 0; k < kGnUnroll; ++k) accum(u[k]);
  }
  for (; r < row1; r += lanes_r) accum(__ldg(reinterpret_cast<const uint4*>(xb + r * ldx)));
  {
    float* ss = s_red + (static_cast<size_t>(rl) * C + vc * 8);
    float* qq = ss + static_cast<size_t>(lanes_r) * C;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ss[j] = s[j];
      qq[j] = q[j];
    }
  }
  __syncthreads();
  // CTA partial of group g = sum over its cg channels x lanes_r row lanes: one warp per group, shuffle tree
  const int cta_id = sample * chunks + blockIdx.x;
  for (int g = warp < nwarps ? warp : groups; g < groups; g += nwarps) {
    float a = 0.f, b = 0.f;
    const int n = cg * lanes_r;
    for (int i = lane; i < n; i += 32) {
      const int rr = i / cg, cc = g * cg + (i - rr * cg);
      a += s_red[static_cast<size_t>(rr) * C + cc];
      b += s_red[static_cast<size_t>(lanes_r + rr) * C + cc];
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) partial[static_cast<size_t>(cta_id) * groups + g] = make_float2(a, b);
  }
  // ---- grid-wide barrier (sense reversal: works for any grid size launch after launch on one stream)
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int total = gridDim.x * gridDim.y;
    const unsigned int my_gen = ld_acquire_u32(&bar->gen);
    __threadfence();
    const unsigned int prev = atomicAdd(&bar->count, 1u);
    if (prev == total - 1) {
      bar->count = 0;
      __threadfence();
      st_release_u32(&bar->gen, my_gen + 1);
    } else {
      while (ld_acquire_u32(&bar->gen) == my_gen) __nanosleep(64);
    }
  }
  __syncthreads();
  // ---- phase 2: mean / rstd of this sample's groups (fp64, fixed chunk order), then normalise
  float* s_stat = s_red;  // [groups][2]
  const double cnt = static_cast<double>(rows_per_sample) * cg;
  for (int g = warp < nwarps ? warp : groups; g < groups; g += nwarps) {
    double a = 0.0, b = 0.0;
    for (int k = lane; k < chunks; k += 32) {
      const float2 p = __ldcg(&partial[static_cast<size_t>(sample * chunks + k) * groups + g]);
      a += static_cast<double>(p.x);
      b += static_cast<double>(p.y);
    }
    a = warp_sum_f64(a);
    b = warp_sum_f64(b);
    if (lane == 0) {
      const double mean = a / cnt;
      double var = b / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      s_stat[2 * g] = static_cast<float>(mean);
      s_stat[2 * g + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
  }
  __syncthreads();
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = vc * 8 + j;
    const int g = c / cg;
    const float a = s_stat[2 * g + 1] * __ldg(gamma + c);
    sc[j] = a;
    sh[j] = __ldg(beta + c) - s_stat[2 * g] * a;
  }
  bf16* yb = y + srow * static_cast<long long>(C) + vc * 8;
  auto emit = [&](const uint4& u, long long rr) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      float a = fmaf(f.x, sc[2 * j], sh[2 * j]);
      float b = fmaf(f.y, sc[2 * j + 1], sh[2 * j + 1]);
      if (silu) {
        a = silu_f(a);
        b = silu_f(b);
      }
      o[j] = pack_bf16x2(a, b);
    }
    *reinterpret_cast<uint4*>(yb + rr * static_cast<long long>(C)) = make_uint4(o[0], o[1], o[2], o[3]);
  };
  // reverse order: this thread's rows are row0 + rl + i * lanes_r, i = 0 .. n_i - 1
  const long long span = row1 - row0 - rl;
  long long n_i = span > 0 ? (span + lanes_r - 1) / lanes_r : 0;
  long long i = n_i;
  for (; i >= kGnUnroll; i -= kGnUnroll) {
    uint4 u[kGnUnroll];
#pragma unroll
    for (int k = 0; k < kGnUnroll; ++k)
      u[k] = __ldg(reinterpret_cast<const uint4*>(xb + (row0 + rl + (i - 1 - k) * lanes_r) * ldx));
#pragma unroll
    for (int k = 0; k < kGnUnroll; ++k) emit(u[k], row0 + rl + (i - 1 - k) * lanes_r);
  }
  for (; i > 0; --i) {
    const long long rr = row0 + rl + (i - 1) * lanes_r;
    emit(__ldg(reinterpret_cast<const uint4*>(xb + rr * ldx)), rr);
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over C (<= 2048), one warp per row, row held in registers.
//   z = x[row] + add[row / rows_per_frame]   (add optional, fp32)
//   ysum[row] = bf16(z)                      (optional)
//   y[row] = (z - mean) * rstd * gamma + beta
// ------------------------------------------------------------------------------------------------
constexpr int kLnMaxVec = 8;  // 8 vectors * 8 elements * 32 lanes = 2048 channels

template <int VPL>  // 16-byte vectors per lane: C <= VPL * 256
__global__ void __launch_bounds__(256)
layernorm_kernel(const bf16* __restrict__ x, const float* __restrict__ add, bf16* __restrict__ ysum,
                 bf16* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
                 long long rows, int C, int rows_per_frame, float eps) {
  const int warps = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int vpr = C >> 3;
  for (long long row = static_cast<long long>(blockIdx.x) * warps + (threadIdx.x >> 5); row < rows;
       row += static_cast<long long>(gridDim.x) * warps) {
    float v[VPL][8];
    const bf16* xr = x + row * C;
    const float* ar = add ? add + (row / rows_per_frame) * C : nullptr;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vc = lane + i * 32;
      if (vc < vpr) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(xr + vc * 8));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16x2(w[j]);
          v[i][2 * j] = f.x;
          v[i][2 * j + 1] = f.y;
        }
        if (ar) {
          const float4 a0 = __ldg(reinterpret_cast<const float4*>(ar + vc * 8));
          const float4 a1 = __ldg(reinterpret_cast<const float4*>(ar + vc * 8 + 4));
          v[i][0] += a0.x; v[i][1] += a0.y; v[i][2] += a0.z; v[i][3] += a0.w;
          v[i][4] += a1.x; v[i][5] += a1.y; v[i][6] += a1.z; v[i][7] += a1.w;
          if (ysum) {
            // round once so that the residual stream and the LN input agree bit-for-bit
#pragma unroll
            for (int j = 0; j < 8; ++j) v[i][j] = __bfloat162float(__float2bfloat16_rn(v[i][j]));
            *reinterpret_cast<uint4*>(ysum + row * C + vc * 8) =
                make_uint4(pack_bf16x2(v[i][0], v[i][1]), pack_bf16x2(v[i][2], v[i][3]),
                           pack_bf16x2(v[i][4], v[i][5]), pack_bf16x2(v[i][6], v[i][7]));
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[i][j];
      }
    }
    s = warp_sum(s);
    const float mean = s / static_cast<float>(C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (lane + i * 32 < vpr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[i][j] - mean;
          q += d * d;
        }
      }
    }
    q = warp_sum(q);
    const float rstd = rsqrtf(q / static_cast<float>(C) + eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vc = lane + i * 32;
      if (vc < vpr) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vc * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vc * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vc * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + vc * 8 + 4));
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * gg[j] + bb[j];
        *reinterpret_cast<uint4*>(y + row * C + vc * 8) =
            make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]),
                       pack_bf16x2(o[6], o[7]));
      }
    }
  }
}

// LayerNorm, C = LPR * 40 (320 / 640 / 1280): LPR lanes per row, 5 vectors (40 channels) per lane, 32/LPR rows per
// warp in flight.  Compared with one-warp-per-row this keeps every lane busy at C = 320 (40 vectors), cuts the
// shuffle reductions to log2(LPR) steps and quadruples the bytes in flight per warp.
template <int LPR>
__global__ void __launch_bounds__(256)
layernorm40_kernel(const bf16* __restrict__ x, const float* __restrict__ add, bf16* __restrict__ ysum,
                   bf16* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
                   long long rows, int rows_per_frame, float eps) {
  constexpr int C = LPR * 40;
  constexpr int RPW = 32 / LPR;  // rows per warp per iteration
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR;     // which row of the warp's group
  const int l = lane % LPR;       // lane within the row
  const long long warp_id = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long row0 = warp_id * RPW; row0 < rows; row0 += nwarps * RPW) {
    const long long row = row0 + sub;
    const bool ok = row < rows;
    float v[5][8];
    float s = 0.f;
    if (ok) {
      const bf16* xr = x + row * C;
      uint4 u[5];
#pragma unroll
      for (int i = 0; i < 5; ++i) u[i] = __ldg(reinterpret_cast<const uint4*>(xr + (l + i * LPR) * 8));
      const float* ar = add ? add + (row / rows_per_frame) * C : nullptr;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const uint32_t w[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16x2(w[j]);
          v[i][2 * j] = f.x;
          v[i][2 * j + 1] = f.y;
        }
        if (ar) {
          const int c0 = (l + i * LPR) * 8;
          const float4 a0 = __ldg(reinterpret_cast<const float4*>(ar + c0)), a1 = __ldg(reinterpret_cast<const float4*>(ar + c0 + 4));
          v[i][0] += a0.x; v[i][1] += a0.y; v[i][2] += a0.z; v[i][3] += a0.w;
          v[i][4] += a1.x; v[i][5] += a1.y; v[i][6] += a1.z; v[i][7] += a1.w;
          if (ysum) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[i][j] = __bfloat162float(__float2bfloat16_rn(v[i][j]));
            *reinterpret_cast<uint4*>(ysum + row * C + c0) =
                make_uint4(pack_bf16x2(v[i][0], v[i][1]), pack_bf16x2(v[i][2], v[i][3]), pack_bf16x2(v[i][4], v[i][5]),
                           pack_bf16x2(v[i][6], v[i][7]));
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[i][j];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        q = fmaf(d, d, q);
      }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / C) + eps);
    if (ok) {
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const int c0 = (l + i * LPR) * 8;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf((v[i][j] - mean) * rstd, gg[j], bb[j]);
        *reinterpret_cast<uint4*>(y + row * C + (l + i * LPR) * 8) =
            make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// In-place row softmax over bf16 scores (decoder AttnBlock, one 4096-wide row per CTA).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
softmax_rows_kernel(bf16* __restrict__ x, int n, float scale) {
  __shared__ float red[8];
  __shared__ float bcast;
  bf16* row = x + static_cast<long long>(blockIdx.x) * n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = n >> 3;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const uint4 u = *reinterpret_cast<const uint4*>(row + i * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      m = fmaxf(m, fmaxf(f.x, f.y));
    }
  }
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = red[0];
    for (int i = 1; i < (blockDim.x >> 5); ++i) t = fmaxf(t, red[i]);
    bcast = t;
  }
  __syncthreads();
  m = bcast * scale;
  float s = 0.f;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const uint4 u = *reinterpret_cast<const uint4*>(row + i * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      s += __expf(f.x * scale - m) + __expf(f.y * scale - m);
    }
  }
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += red[i];
    bcast = 1.f / t;
  }
  __syncthreads();
  const float inv = bcast;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const uint4 u = *reinterpret_cast<const uint4*>(row + i * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      o[j] = pack_bf16x2(__expf(f.x * scale - m) * inv, __expf(f.y * scale - m) * inv);
    }
    *reinterpret_cast<uint4*>(row + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// fp32 scores -> bf16 probabilities (separate buffers): keeps the decoder's 4096-wide softmax input in fp32.
__global__ void __launch_bounds__(256)
softmax_rows_f32_kernel(const float* __restrict__ x, bf16* __restrict__ y, int n, float scale) {
  __shared__ float red[8];
  __shared__ float bcast;
  const float* row = x + static_cast<long long>(blockIdx.x) * n;
  bf16* orow = y + static_cast<long long>(blockIdx.x) * n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = n >> 2;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const float4 f = *reinterpret_cast<const float4*>(row + i * 4);
    m = fmaxf(m, fmaxf(fmaxf(f.x, f.y), fmaxf(f.z, f.w)));
  }
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = red[0];
    for (int i = 1; i < (blockDim.x >> 5); ++i) t = fmaxf(t, red[i]);
    bcast = t;
  }
  __syncthreads();
  m = bcast * scale;
  float s = 0.f;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const float4 f = *reinterpret_cast<const float4*>(row + i * 4);
    s += __expf(f.x * scale - m) + __expf(f.y * scale - m) + __expf(f.z * scale - m) + __expf(f.w * scale - m);
  }
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += red[i];
    bcast = 1.f / t;
  }
  __syncthreads();
  const float inv = bcast;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const float4 f = *reinterpret_cast<const float4*>(row + i * 4);
    *reinterpret_cast<uint2*>(orow + i * 4) =
        make_uint2(pack_bf16x2(__expf(f.x * scale - m) * inv, __expf(f.y * scale - m) * inv),
                   pack_bf16x2(__expf(f.z * scale - m) * inv, __expf(f.w * scale - m) * inv));
  }
}

static int pick_rows_per_cta(long long rows_per_sample, int nsamples) {
  // a few CTAs per SM, each with enough rows that the per-CTA prologue / reduction tail stays small next to the
  // streaming loop (V3D_GN_WAVES overrides the CTA-per-SM target: tuning knob)
  static int waves = 0;
  if (waves == 0) {
    const char* v = getenv("V3D_GN_WAVES");
    waves = v ? atoi(v) : 4;
    if (waves < 1) waves = 1;
  }
  const long long target = static_cast<long long>(waves) * num_sms();
  long long chunks = (target + nsamples - 1) / nsamples;
  if (chunks < 1) chunks = 1;
  long long rpc = (rows_per_sample + chunks - 1) / chunks;
  if (rpc < 32) rpc = 32;
  if (rpc > rows_per_sample) rpc = rows_per_sample;
  return static_cast<int>(rpc);
}

}  // namespace v3d

using namespace v3d;

extern "C" {

/* GroupNorm statistics over [nsamples][rows_per_sample][C] (bf16, row stride ldx) -> stats double
 * [nsamples][groups][2] = {sum, sumsq}. rows_per_sample = H*W for 2-D GroupNorm32/Normalize
 * (diffusionmodules/util.py:259-276, attention.py:130-133, model.py:52-55) and T*H*W for the 3-D
 * time_stack ResBlock (openaimodel.py:267-271 with dims=3: reduction over (C/32, T, H, W)). */
int v3d_groupnorm_stats(const void* x, void* stats, int64_t rows_per_sample, int32_t nsamples, int32_t C,
                        int32_t ldx, int32_t groups, int32_t pre_zeroed, void* stream) {
  if (!x || !stats || C % 8 != 0 || C % groups != 0 || C / 8 > 512 || ldx % 8 != 0 ||
      rows_per_sample <= 0 || nsamples <= 0) {
    set_error("v3d_groupnorm_stats: bad args C=%d groups=%d ldx=%d", C, groups, ldx);
    return V3D_ERR_BAD_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!pre_zeroed) {
    cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(double) * 2 * groups * nsamples, st);
    if (e != cudaSuccess) {
      set_error("memset stats: %s", cudaGetErrorString(e));
      return V3D_ERR_CUDA;
    }
  }
  const int rpc = pick_rows_per_cta(rows_per_sample, nsamples);
  dim3 grid(static_cast<unsigned>((rows_per_sample + rpc - 1) / rpc), nsamples);
  const int vpr = C / 8;
  const int lanes_r = kGnThreads / vpr > 0 ? kGnThreads / vpr : 1;
  gn_stats_kernel<<<grid, vpr * lanes_r, 2 * C * sizeof(float), st>>>(
      static_cast<const bf16*>(x), static_cast<double*>(stats), rows_per_sample, C, ldx, groups, rpc);
  V3D_CHECK_LAUNCH("gn_stats_kernel");
  return V3D_OK;
}

/* y[dense, ld=C] = act(GroupNorm(x; stats, gamma, beta, eps)), act = SiLU when silu != 0
 * (fuses normalization() + nn.SiLU of openaimodel.py:267-271,300-303; model.py:131-143 nonlinearity). */
int v3d_groupnorm_apply(const void* x, void* y, const void* stats, const void* gamma, const void* beta,
                        int64_t rows_per_sample, int32_t nsamples, int32_t C, int32_t ldx, int32_t groups,
                        float eps, int32_t silu, void* stream) {
  if (!x || !y || !stats || !gamma || !beta || C % 8 != 0 || C % groups != 0 || ldx % 8 != 0) {
    set_error("v3d_groupnorm_apply: bad args");
    return V3D_ERR_BAD_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (C / 8 > 512) {
    set_error("v3d_groupnorm_apply: C=%d too wide", C);
    return V3D_ERR_BAD_ARG;
  }
  const int rpc = pick_rows_per_cta(rows_per_sample, nsamples);
  dim3 grid(static_cast<unsigned>((rows_per_sample + rpc - 1) / rpc), nsamples);
  const int vpr = C / 8;
  const int lanes_r = kGnThreads / vpr > 0 ? kGnThreads / vpr : 1;
  gn_apply_kernel<<<grid, vpr * lanes_r, 0, st>>>(
      static_cast<const bf16*>(x), static_cast<bf16*>(y), static_cast<const double*>(stats),
      static_cast<const float*>(gamma), static_cast<const float*>(beta), rows_per_sample, C, ldx, groups,
      eps, silu, rpc);
  V3D_CHECK_LAUNCH("gn_apply_kernel");
  return V3D_OK;
}

/* Bytes of the workspace v3d_groupnorm needs (barrier words + per-CTA partials); the caller allocates it once per
 * stream, ZEROED, and passes the same buffer to every call on that stream. */
int64_t v3d_groupnorm_workspace_bytes(void) { return static_cast<int64_t>(kGnWsBytes); }

/* y[dense, ld=C] = act(GroupNorm(x; gamma, beta, eps)) in ONE launch (statistics + grid barrier + apply); same
 * operand meaning as the v3d_groupnorm_stats / v3d_groupnorm_apply pair it replaces on the unsharded path
 * (GroupNorm32 + SiLU openaimodel.py:267-271,300-303, diffusionmodules/util.py:259-276; Normalize model.py:52-55,
 * attention.py:130-133).  Deterministic (no atomics on data).  Calls sharing a workspace must be stream-ordered. */
int v3d_groupnorm(const void* x, void* y, const void* gamma, const void* beta, int64_t rows_per_sample,
                  int32_t nsamples, int32_t C, int32_t ldx, int32_t groups, float eps, int32_t silu,
                  void* workspace, int64_t workspace_bytes, void* stream) {
  if (!x || !y || !gamma || !beta || !workspace || C % 8 != 0 || groups <= 0 || groups > 32 || C % groups != 0 ||
      C / 8 > 512 || ldx % 8 != 0 || rows_per_sample <= 0 || nsamples <= 0 ||
      workspace_bytes < static_cast<int64_t>(kGnWsBytes)) {
    set_error("v3d_groupnorm: bad args C=%d groups=%d ldx=%d workspace=%lld", C, groups, ldx,
              static_cast<long long>(workspace_bytes));
    return V3D_ERR_BAD_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int vpr = C / 8;
  const int lanes_r = kGnThreads / vpr > 0 ? kGnThreads / vpr : 1;
  const int threads = vpr * lanes_r;
  const size_t smem = sizeof(float) * 2 * static_cast<size_t>(lanes_r) * C;
  // co-resident capacity of this launch shape (the barrier needs every CTA on the device at once)
  int per_sm = 0;
  auto kern = threads <= 256 ? gn_fused_kernel<256, 3> : gn_fused_kernel<512, 1>;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
  if (e != cudaSuccess || per_sm < 1) {
    set_error("v3d_groupnorm: occupancy query failed (%s)", cudaGetErrorString(e));
    return V3D_ERR_CUDA;
  }
  if (per_sm > 4) per_sm = 4;
  long long cap = static_cast<long long>(per_sm) * num_sms();
  if (cap > kGnMaxCtas) cap = kGnMaxCtas;
  if (nsamples > cap) {
    // more samples than co-resident CTAs: fall back to the two-kernel form on the caller's side
    set_error("v3d_groupnorm: %d samples exceed the co-resident grid (%lld CTAs)", nsamples, cap);
    return V3D_ERR_UNSUPPORTED;
  }
  long long chunks = cap / nsamples;
  long long max_chunks = rows_per_sample / (16LL * lanes_r);  // >= 16 rows per thread: the barrier and the partials
                                                              // must stay small next to the streaming loops
  if (max_chunks < 1) max_chunks = 1;
  if (chunks > max_chunks) chunks = max_chunks;
  long long rpc = (rows_per_sample + chunks - 1) / chunks;
  chunks = (rows_per_sample + rpc - 1) / rpc;
  dim3 grid(static_cast<unsigned>(chunks), nsamples);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  kern<<<grid, threads, smem, st>>>(
      static_cast<const bf16*>(x), static_cast<bf16*>(y), static_cast<const float*>(gamma),
      static_cast<const float*>(beta), reinterpret_cast<GnBarrier*>(ws), reinterpret_cast<float2*>(ws + 256),
      rows_per_sample, C, ldx, groups, eps, silu, static_cast<int>(rpc));
  V3D_CHECK_LAUNCH("gn_fused_kernel");
  return V3D_OK;
}

/* LayerNorm (attention.py:525-527, video_attention.py:51,79,93-94) with the optional fused
 * "x_mix = x + emb" of video_attention.py:286-287 (add = per-frame fp32 vectors, ysum receives x+emb). */
int v3d_layernorm(const void* x, const void* add, void* ysum, void* y, const void* gamma, const void* beta,
                  int64_t rows, int32_t C, int32_t rows_per_frame, float eps, void* stream) {
  if (!x || !y || !gamma || !beta || C % 8 != 0 || C > kLnMaxVec * 256 || rows <= 0) {
    set_error("v3d_layernorm: bad args C=%d", C);
    return V3D_ERR_BAD_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int warps = 8;
  long long blocks = (rows + warps - 1) / warps;
  const long long cap = 32LL * num_sms();
  if (blocks > cap) blocks = cap;
  if (C == 320 || C == 640 || C == 1280) {
    const int lpr = C / 40;
    const long long rows_per_cta = 8LL * (32 / lpr);
    long long nb = (rows + rows_per_cta - 1) / rows_per_cta;
    // grid-stride over rows: V3D_LN_CTAS_PER_SM (tuning knob) bounds the grid in CTAs per SM
    static int ln_ctas = 0;
    if (ln_ctas == 0) {
      const char* v = getenv("V3D_LN_CTAS_PER_SM");
      ln_ctas = v ? atoi(v) : 6;   // 6 resident-sized waves measured best on B200 (16: -5..10 %)
      if (ln_ctas < 1) ln_ctas = 1;
    }
    const long long cap2 = static_cast<long long>(ln_ctas) * num_sms();
    if (nb > cap2) nb = cap2;
#define V3D_LN40(L)                                                                                              \
  layernorm40_kernel<L><<<static_cast<unsigned>(nb), 256, 0, st>>>(                                             \
      static_cast<const bf16*>(x), static_cast<const float*>(add), static_cast<bf16*>(ysum), static_cast<bf16*>(y), \
      static_cast<const float*>(gamma), static_cast<const float*>(beta), rows, rows_per_frame > 0 ? rows_per_frame : 1, eps)
    if (lpr == 8) V3D_LN40(8);
    else if (lpr == 16) V3D_LN40(16);
    else V3D_LN40(32);
#undef V3D_LN40
    V3D_CHECK_LAUNCH("layernorm40_kernel");
    return V3D_OK;
  }
  const int vpl = (C / 8 + 31) / 32;
#define V3D_LN_LAUNCH(V)                                                                                   \
  layernorm_kernel<V><<<static_cast<unsigned>(blocks), warps * 32, 0, st>>>(                                \
      static_cast<const bf16*>(x), static_cast<const float*>(add), static_cast<bf16*>(ysum),               \
      static_cast<bf16*>(y), static_cast<const float*>(gamma), static_cast<const float*>(beta), rows, C,   \
      rows_per_frame > 0 ? rows_per_frame : 1, eps)
  switch (vpl) {
    case 1: V3D_LN_LAUNCH(1); break;
    case 2: V3D_LN_LAUNCH(2); break;
    case 3: V3D_LN_LAUNCH(3); break;
    case 4: V3D_LN_LAUNCH(4); break;
    case 5: V3D_LN_LAUNCH(5); break;
    default: V3D_LN_LAUNCH(8); break;
  }
#undef V3D_LN_LAUNCH
  V3D_CHECK_LAUNCH("layernorm_kernel");
  return V3D_OK;
}

/* In-place softmax(scale * x) over rows of n bf16 scores (decoder AttnBlock, model.py:190-192). */
int v3d_softmax_rows(void* x, int64_t rows, int32_t n, float scale, void* stream) {
  if (!x || n % 8 != 0 || rows <= 0 || rows > 0x7fffffffLL) {
    set_error("v3d_softmax_rows: bad args");
    return V3D_ERR_BAD_ARG;
  }
  softmax_rows_kernel<<<static_cast<unsigned>(rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<bf16*>(x), n, scale);
  V3D_CHECK_LAUNCH("softmax_rows_kernel");
  return V3D_OK;
}

/* softmax(scale * x) from fp32 scores to bf16 probabilities (decoder AttnBlock, model.py:190-192). */
int v3d_softmax_rows_f32(const void* x, void* y, int64_t rows, int32_t n, float scale, void* stream) {
  if (!x || !y || n % 4 != 0 || rows <= 0 || rows > 0x7fffffffLL) {
    set_error("v3d_softmax_rows_f32: bad args");
    return V3D_ERR_BAD_ARG;
  }
  softmax_rows_f32_kernel<<<static_cast<unsigned>(rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float*>(x), static_cast<bf16*>(y), n, scale);
  V3D_CHECK_LAUNCH("softmax_rows_f32_kernel");
  return V3D_OK;
}

}  // extern "C"
