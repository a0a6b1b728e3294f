// Memory-bound data-movement and small-matrix kernels around the tensor-core path:
// layout conversion at the sgm boundary (NCHW fp32 <-> NHWC bf16), nearest-2x upsample, channel-slice
// copies (skip concat), explicit im2row for the few convs TMA cannot gather (Cin % 64 != 0, stride 2),
// the M<=64 "embedding" linears, sinusoidal timestep embeddings.
#include "common.cuh"
#include "host_util.cuh"
#include "v3d_b200.h"

namespace v3d {

static inline unsigned blocks_for(long long n, int threads, int max_waves = 32) {
  long long b = (n + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms()) * max_waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<unsigned>(b);
}

// ---------------- nearest 2x upsample, NHWC bf16, 16-byte vectors ----------------
__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H, int W,
                                  int vpc) {
  const long long total = static_cast<long long>(N) * (2 * H) * (2 * W) * vpc;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vpc);
    long long p = i / vpc;
    const int ow = static_cast<int>(p % (2 * W));
    p /= (2 * W);
    const int oh = static_cast<int>(p % (2 * H));
    const long long n = p / (2 * H);
    y[i] = __ldg(&x[((n * H + (oh >> 1)) * W + (ow >> 1)) * vpc + v]);
  }
}

// ---------------- strided 2-D copy of [rows][ncols] bf16 (16-byte vectors) ----------------
__global__ void copy_channels_kernel(const bf16* __restrict__ src, long long lds, bf16* __restrict__ dst,
                                     long long ldd, long long rows, int vpr) {
  const long long total = rows * vpr;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / vpr;
    const int v = static_cast<int>(i % vpr);
    *reinterpret_cast<uint4*>(dst + r * ldd + v * 8) = __ldg(reinterpret_cast<const uint4*>(src + r * lds + v * 8));
  }
}

// ---------------- explicit im2row for 3x3 convs: y[n,oh,ow][tap*C + c], zero padded to Kpad ----------------
__global__ void im2col3x3_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int N, int H, int W, int C,
                                 int stride, int pad, int Ho, int Wo, int Kpad) {
  const long long total = static_cast<long long>(N) * Ho * Wo * Kpad;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % Kpad);
    long long p = i / Kpad;
    const int ow = static_cast<int>(p % Wo);
    p /= Wo;
    const int oh = static_cast<int>(p % Ho);
    const long long n = p / Ho;
    bf16 v = __float2bfloat16(0.f);
    if (k < 9 * C) {
      const int tap = k / C, c = k - tap * C;
      const int ih = oh * stride + tap / 3 - pad;
      const int iw = ow * stride + tap % 3 - pad;
      if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = x[((n * H + ih) * W + iw) * C + c];
    }
    y[i] = v;
  }
}

// vectorised variant for C % 8 == 0 and Kpad == 9*C: one 16-byte chunk per thread
__global__ void im2col3x3_vec_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H, int W, int vpc,
                                     int stride, int pad, int Ho, int Wo) {
  const long long total = static_cast<long long>(N) * Ho * Wo * 9 * vpc;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vpc);
    long long p = i / vpc;
    const int tap = static_cast<int>(p % 9);
    p /= 9;
    const int ow = static_cast<int>(p % Wo);
    p /= Wo;
    const int oh = static_cast<int>(p % Ho);
    const long long n = p / Ho;
    const int ih = oh * stride + tap / 3 - pad;
    const int iw = ow * stride + tap % 3 - pad;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) val = __ldg(&x[((n * H + ih) * W + iw) * vpc + v]);
    y[i] = val;
  }
}

// ---------------- NCHW fp32 -> NHWC bf16 (optional scale), via smem transpose ----------------
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, bf16* __restrict__ y, int C, int HW, float scale) {
  // grid: (ceil(HW/32), N); block 32 x 8. Each block transposes a [C][32-pixel] slab, C in chunks of 32.
  __shared__ float tile[32][33];
  const long long n = blockIdx.y;
  const int p0 = blockIdx.x * 32;
  for (int c0 = 0; c0 < C; c0 += 32) {
    for (int cy = threadIdx.y; cy < 32; cy += 8) {
      const int c = c0 + cy, p = p0 + threadIdx.x;
      tile[cy][threadIdx.x] = (c < C && p < HW) ? x[(n * C + c) * HW + p] : 0.f;
    }
    __syncthreads();
    for (int py = threadIdx.y; py < 32; py += 8) {
      const int p = p0 + py, c = c0 + threadIdx.x;
      if (p < HW && c < C) y[(n * HW + p) * C + c] = __float2bfloat16_rn(tile[threadIdx.x][py] * scale);
    }
    __syncthreads();
  }
}

// ---------------- NHWC (bf16 or fp32, row stride ldx, first C channels) -> NCHW fp32 ----------------
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ x, float* __restrict__ y, int C, int HW, long long ldx,
                                    float scale) {
  __shared__ float tile[32][33];
  const long long n = blockIdx.y;
  const int p0 = blockIdx.x * 32;
  for (int c0 = 0; c0 < C; c0 += 32) {
    for (int py = threadIdx.y; py < 32; py += 8) {
      const int p = p0 + py, c = c0 + threadIdx.x;
      tile[py][threadIdx.x] = (p < HW && c < C) ? static_cast<float>(x[(n * HW + p) * ldx + c]) : 0.f;
    }
    __syncthreads();
    for (int cy = threadIdx.y; cy < 32; cy += 8) {
      const int c = c0 + cy, p = p0 + threadIdx.x;
      if (c < C && p < HW) y[(n * C + c) * HW + p] = tile[threadIdx.x][cy] * scale;
    }
    __syncthreads();
  }
}

// ---------------- small-M linear: y[m,n] (+)= act_out( sum_k act_in(x[m,k]) W[n,k] + b[n] ) ----------------
// x fp32 [M,K] (M <= 64), W bf16 [N,K], y fp32 [M,N].  A CTA of 8 warps owns 32 output columns (4 per warp).
// x is staged once per K-chunk in shared memory with act_in already applied (so SiLU is evaluated once per CTA,
// not once per output column); each lane streams 16-byte weight vectors and FMAs them against smem rows.
constexpr int kSlMaxM = 64;
constexpr int kSlCols = 4;     // output columns per warp
constexpr int kSlKc = 256;     // K-chunk staged in smem (one 8-element vector per lane)
constexpr int kSlMc = 16;      // rows of x accumulated per pass

__global__ void __launch_bounds__(256)
small_linear_kernel(const float* __restrict__ x, const bf16* __restrict__ W, const float* __restrict__ bias,
                    float* __restrict__ y, int M, int K, int N, int act_in, int act_out, int accumulate,
                    long long ldx, long long ldy) {
  __shared__ __align__(16) float sx[kSlMc][kSlKc + 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n0 = (blockIdx.x * 8 + warp) * kSlCols;
  // per-lane partial sums live in registers per pass; final sums are reduced across lanes
  for (int m0 = 0; m0 < M; m0 += kSlMc) {
    float part[kSlCols][kSlMc];
#pragma unroll
    for (int cidx = 0; cidx < kSlCols; ++cidx)
#pragma unroll
      for (int i = 0; i < kSlMc; ++i) part[cidx][i] = 0.f;
    for (int k0 = 0; k0 < K; k0 += kSlKc) {
      __syncthreads();
      // stage x[m0..m0+Mc)[k0..k0+Kc) with act_in applied
      for (int i = threadIdx.x; i < kSlMc * (kSlKc / 4); i += blockDim.x) {
        const int r = i / (kSlKc / 4), c4 = (i % (kSlKc / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + r < M && k0 + c4 < K) v = __ldg(reinterpret_cast<const float4*>(x + static_cast<long long>(m0 + r) * ldx + k0 + c4));
        if (act_in == V3D_ACT_SILU) {
          v.x = v.x / (1.0f + expf(-v.x)); v.y = v.y / (1.0f + expf(-v.y));
          v.z = v.z / (1.0f + expf(-v.z)); v.w = v.w / (1.0f + expf(-v.w));
        }
        *reinterpret_cast<float4*>(&sx[r][c4]) = v;
      }
      __syncthreads();
      const int k = k0 + lane * 8;
      if (k < K) {
        float w[kSlCols][8];
#pragma unroll
        for (int cidx = 0; cidx < kSlCols; ++cidx) {
          const int n = n0 + cidx;
          uint4 u = make_uint4(0, 0, 0, 0);
          if (n < N) u = __ldg(reinterpret_cast<const uint4*>(W + static_cast<long long>(n) * K + k));
          const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(wv[j]);
            w[cidx][2 * j] = f.x;
            w[cidx][2 * j + 1] = f.y;
          }
        }
#pragma unroll
        for (int i = 0; i < kSlMc; ++i) {
          const float4 a = *reinterpret_cast<const float4*>(&sx[i][lane * 8]);
          const float4 b = *reinterpret_cast<const float4*>(&sx[i][lane * 8 + 4]);
          const float xv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
          for (int cidx = 0; cidx < kSlCols; ++cidx)
#pragma unroll
            for (int j = 0; j < 8; ++j) part[cidx][i] = fmaf(xv[j], w[cidx][j], part[cidx][i]);
        }
      }
    }
#pragma unroll
    for (int cidx = 0; cidx < kSlCols; ++cidx) {
      const int n = n0 + cidx;
      const float bv = (bias != nullptr && n < N) ? bias[n] : 0.f;
#pragma unroll
      for (int i = 0; i < kSlMc; ++i) {
        const float sres = warp_sum(part[cidx][i]);
        if (lane == 0 && n < N && m0 + i < M) {
          float v = sres + bv;
          if (act_out == V3D_ACT_SILU) v = v / (1.0f + expf(-v));
          float* o = y + static_cast<long long>(m0 + i) * ldy + n;
          *o = accumulate ? *o + v : v;
        }
      }
    }
  }
}

// ---------------- sinusoidal embedding: out[i] = [cos(t_i f_k) | sin(t_i f_k)], f_k = exp(-ln(P) k / half) ----
__global__ void timestep_embedding_kernel(const float* __restrict__ t, float* __restrict__ out, int n, int dim,
                                          float max_period) {
  const int half = dim / 2;
  const int total = n * dim;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / dim, c = i - r * dim;
    float v = 0.f;
    if (c < 2 * half) {
      const int k = c < half ? c : c - half;
      const float f = expf(-logf(max_period) * static_cast<float>(k) / static_cast<float>(half));
      const float a = t[r] * f;
      v = c < half ? cosf(a) : sinf(a);
    }
    out[i] = v;
  }
}

// ---------------- operand prep for the small-M tensor-core path: y[64][K] bf16 = act_in(x[M][ldx]) (rows >= M zero)
__global__ void prep_small_x_kernel(const float* __restrict__ x, long long ldx, bf16* __restrict__ y, int M, int K,
                                    int act_in) {
  const int total = 64 * K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / K, c = i - r * K;
    float v = 0.f;
    if (r < M) {
      v = x[static_cast<long long>(r) * ldx + c];
      if (act_in == V3D_ACT_SILU) v = v / (1.0f + expf(-v));
    }
    y[i] = __float2bfloat16_rn(v);
  }
}

__global__ void add_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o,
                                long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    o[i] = a[i] + b[i];
}

// ---------------- AE3DConv time_mix_conv: Conv3d C->C (C <= 4), k=(3,1,1), pad (1,0,0), NHWC fp32 in, NCHW fp32 out
__global__ void time_mix_conv_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ w,
                                     const float* __restrict__ bias, float* __restrict__ y, int nb, int T,
                                     long long HW, int C) {
  const long long total = static_cast<long long>(nb) * T * HW;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i % HW;
    const long long f = i / HW;  // frame index b*T + t
    const int t = static_cast<int>(f % T);
    float acc[4];
    for (int co = 0; co < C; ++co) acc[co] = bias[co];
    for (int tap = 0; tap < 3; ++tap) {
      const int tt = t + tap - 1;
      if (tt < 0 || tt >= T) continue;
      const float* xp = x + ((f + tap - 1) * HW + p) * ldx;
      for (int ci = 0; ci < C; ++ci) {
        const float v = xp[ci];
        for (int co = 0; co < C; ++co) acc[co] = fmaf(w[(co * C + ci) * 3 + tap], v, acc[co]);
      }
    }
    for (int co = 0; co < C; ++co) y[(f * C + co) * HW + p] = acc[co];
  }
}

}  // namespace v3d

using namespace v3d;

extern "C" {

/* F.interpolate(scale_factor=2, mode="nearest") on NHWC bf16 (openaimodel.py:164; model.py:68). */
int v3d_upsample_nearest2x(const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, void* stream) {
  if (!x || !y || C % 8 != 0 || N <= 0 || H <= 0 || W <= 0) {
    set_error("v3d_upsample_nearest2x: bad args");
    return V3D_ERR_BAD_ARG;
  }
  const long long total = 4LL * N * H * W * (C / 8);
  upsample2x_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(y), N, H, W, C / 8);
  V3D_CHECK_LAUNCH("upsample2x_kernel");
  return V3D_OK;
}

/* Copy a [rows][ncols] bf16 block between row-strided buffers: the two halves of
 * th.cat([h, hs.pop()], dim=1) (video_model.py:483) written into one NHWC buffer. */
int v3d_copy_channels(const void* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows, int32_t ncols,
                      void* stream) {
  if (!src || !dst || ncols % 8 != 0 || ld_src % 8 != 0 || ld_dst % 8 != 0 || rows <= 0) {
    set_error("v3d_copy_channels: bad args");
    return V3D_ERR_BAD_ARG;
  }
  copy_channels_kernel<<<blocks_for(rows * (ncols / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(src), ld_src, static_cast<bf16*>(dst), ld_dst, rows, ncols / 8);
  V3D_CHECK_LAUNCH("copy_channels_kernel");
  return V3D_OK;
}

/* Explicit im2row (tap-major, then channel; zero-padded to Kpad columns) for 3x3 convs the TMA gather
 * does not cover: Cin % 64 != 0 (video_model.py:189 8->320; model.py:651 4->512) and stride 2
 * (openaimodel.py:202-209 Downsample; model.py:82-90 with pad=0). */
int v3d_im2col3x3(const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, int32_t stride, int32_t pad,
                  int32_t Hout, int32_t Wout, int32_t Kpad, void* stream) {
  if (!x || !y || Kpad < 9 * C || stride <= 0) {
    set_error("v3d_im2col3x3: bad args");
    return V3D_ERR_BAD_ARG;
  }
  const long long total = static_cast<long long>(N) * Hout * Wout * Kpad;
  if (C % 8 == 0 && Kpad == 9 * C) {
    im2col3x3_vec_kernel<<<blocks_for(total / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(x), static_cast<uint4*>(y), N, H, W, C / 8, stride, pad, Hout, Wout);
    V3D_CHECK_LAUNCH("im2col3x3_vec_kernel");
    return V3D_OK;
  }
  im2col3x3_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(x), static_cast<bf16*>(y), N, H, W, C, stride, pad, Hout, Wout, Kpad);
  V3D_CHECK_LAUNCH("im2col3x3_kernel");
  return V3D_OK;
}

/* sgm boundary: the reference hands the network NCHW fp32 (wrappers.py:27, video_diffusion.py:184). */
int v3d_nchw_f32_to_nhwc_bf16(const void* x, void* y, int32_t N, int32_t C, int32_t H, int32_t W, float scale,
                              void* stream) {
  if (!x || !y || N <= 0 || C <= 0 || N > 65535) {
    set_error("v3d_nchw_f32_to_nhwc_bf16: bad args");
    return V3D_ERR_BAD_ARG;
  }
  dim3 grid((H * W + 31) / 32, N), block(32, 8);
  nchw_to_nhwc_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float*>(x), static_cast<bf16*>(y), C, H * W, scale);
  V3D_CHECK_LAUNCH("nchw_to_nhwc_kernel");
  return V3D_OK;
}

int v3d_nhwc_to_nchw_f32(const void* x, void* y, int32_t N, int32_t C, int32_t HW, int64_t ldx, int32_t src_fp32,
                         float scale, void* stream) {
  if (!x || !y || N <= 0 || C <= 0 || N > 65535) {
    set_error("v3d_nhwc_to_nchw_f32: bad args");
    return V3D_ERR_BAD_ARG;
  }
  dim3 grid((HW + 31) / 32, N), block(32, 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (src_fp32)
    nhwc_to_nchw_kernel<float><<<grid, block, 0, st>>>(static_cast<const float*>(x), static_cast<float*>(y), C, HW,
                                                       ldx, scale);
  else
    nhwc_to_nchw_kernel<bf16><<<grid, block, 0, st>>>(static_cast<const bf16*>(x), static_cast<float*>(y), C, HW,
                                                      ldx, scale);
  V3D_CHECK_LAUNCH("nhwc_to_nchw_kernel");
  return V3D_OK;
}

/* The M<=64 linears on embeddings: time_embed / label_emb (video_model.py:151-182,456-461), ResBlock
 * emb_layers (openaimodel.py:291-297), time_pos_embed (video_attention.py:220-224,275), and the
 * single-token cross-attention collapse to_out(to_v(ctx)) (attention.py:277-283 with one key). */
int v3d_small_linear(const void* x, const void* W, const void* bias, void* y, int32_t M, int32_t K, int32_t N,
                     int32_t act_in, int32_t act_out, int32_t accumulate, int64_t ldx, int64_t ldy, void* stream) {
  if (ldx <= 0) ldx = K;
  if (ldy <= 0) ldy = N;
  if (!x || !W || !y || M <= 0 || M > kSlMaxM || K % 8 != 0 || N <= 0 || ldx % 4 != 0) {
    set_error("v3d_small_linear: bad args M=%d K=%d N=%d", M, K, N);
    return V3D_ERR_BAD_ARG;
  }
  const int cols_per_cta = 8 * kSlCols;
  small_linear_kernel<<<(N + cols_per_cta - 1) / cols_per_cta, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float*>(x), static_cast<const bf16*>(W), static_cast<const float*>(bias),
      static_cast<float*>(y), M, K, N, act_in, act_out, accumulate, ldx, ldy);
  V3D_CHECK_LAUNCH("small_linear_kernel");
  return V3D_OK;
}

/* timestep_embedding (diffusionmodules/util.py:207-231): cos | sin halves, fp32. */
int v3d_timestep_embedding(const void* t, void* out, int32_t n, int32_t dim, float max_period, void* stream) {
  if (!t || !out || n <= 0 || dim <= 0) {
    set_error("v3d_timestep_embedding: bad args");
    return V3D_ERR_BAD_ARG;
  }
  timestep_embedding_kernel<<<blocks_for(static_cast<long long>(n) * dim, 256), 256, 0,
                              static_cast<cudaStream_t>(stream)>>>(static_cast<const float*>(t),
                                                                   static_cast<float*>(out), n, dim, max_period);
  V3D_CHECK_LAUNCH("timestep_embedding_kernel");
  return V3D_OK;
}

/* AE3DConv.time_mix_conv (temporal_ae.py:94-107): Conv3d(C, C, (3,1,1), pad (1,0,0)) across the frames of each
 * video, C <= 4. x: NHWC fp32 [nb*T*HW][ldx] (first C channels), w: [C][C][3] fp32, y: NCHW fp32 [nb*T][C][HW]. */
int v3d_time_mix_conv(const void* x, int64_t ldx, const void* w, const void* bias, void* y, int32_t nb, int32_t T,
                      int64_t HW, int32_t C, void* stream) {
  if (!x || !w || !bias || !y || C <= 0 || C > 4 || nb <= 0 || T <= 0 || HW <= 0) {
    set_error("v3d_time_mix_conv: bad args");
    return V3D_ERR_BAD_ARG;
  }
  time_mix_conv_kernel<<<blocks_for(static_cast<long long>(nb) * T * HW, 256), 256, 0,
                         static_cast<cudaStream_t>(stream)>>>(static_cast<const float*>(x), ldx,
                                                              static_cast<const float*>(w),
                                                              static_cast<const float*>(bias),
                                                              static_cast<float*>(y), nb, T, HW, C);
  V3D_CHECK_LAUNCH("time_mix_conv_kernel");
  return V3D_OK;
}

/* Operand prep for running an M<=64 linear on the tensor cores as W[N,K] x X[64,K]^T (v3d_gemm_bf16 with
 * out_transposed): y bf16 [64][K] = act_in(x fp32 [M][ldx]), rows M..63 zero. */
int v3d_prep_small_x(const void* x, int64_t ldx, void* y, int32_t M, int32_t K, int32_t act_in, void* stream) {
  if (!x || !y || M <= 0 || M > 64 || K <= 0) {
    set_error("v3d_prep_small_x: bad args");
    return V3D_ERR_BAD_ARG;
  }
  prep_small_x_kernel<<<blocks_for(64LL * K, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float*>(x), ldx > 0 ? ldx : K, static_cast<bf16*>(y), M, K, act_in);
  V3D_CHECK_LAUNCH("prep_small_x_kernel");
  return V3D_OK;
}

int v3d_add_rows(const void* a, const void* b, void* out, int32_t rows, int32_t cols, void* stream) {
  if (!a || !b || !out) {
    set_error("v3d_add_rows: bad args");
    return V3D_ERR_BAD_ARG;
  }
  const long long n = static_cast<long long>(rows) * cols;
  add_rows_kernel<<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float*>(a), static_cast<const float*>(b), static_cast<float*>(out), n);
  V3D_CHECK_LAUNCH("add_rows_kernel");
  return V3D_OK;
}

}  // extern "C"
