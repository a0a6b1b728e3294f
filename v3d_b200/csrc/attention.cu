// Attention kernels, head dim 64, bf16 in / fp32 softmax / bf16 out.
//   spatial : flash attention over N tokens per (sample, head); streaming K/V tiles through shared
//             memory with cp.async double buffering, online softmax in registers.
//             (mma.sync m16n8k16 bring-up kernel, exported as v3d_attention_spatial_mma for cross-checks;
//             the product kernel is the tcgen05 one in attention_tc.cu)
//   temporal: sequences of T <= 32 view-frames per (pixel, head); one warp per sequence, whole
//             problem on chip, no online softmax. Reads q/k/v with the frame stride directly from the
//             frame-major token matrix, so "(b t) s c -> (b s) t c" is never materialised.
#include "common.cuh"
#include "host_util.cuh"
#include "v3d_b200.h"

namespace v3d {

constexpr int HD = 64;  // head dim

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                          uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// byte offset of 16B chunk `chunk` of row `row` in a [rows][64] bf16 tile with XOR swizzle
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
  return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}

constexpr int FA_BN = 64;

template <int BM>
__global__ void __launch_bounds__(BM * 2)
attn_spatial_kernel(const bf16* __restrict__ Q, const bf16* __restrict__ K, const bf16* __restrict__ V,
                    bf16* __restrict__ O, long long ld, long long ldo, int ntok, float scale_log2e) {
  constexpr int NT = BM * 2;  // threads: one warp per 16 query rows
  extern __shared__ __align__(128) uint8_t fa_smem[];
  uint8_t* sQ = fa_smem;                      // BM x 128 B
  uint8_t* sK = sQ + BM * 128;                // 2 x 64 x 128 B
  uint8_t* sV = sK + 2 * FA_BN * 128;         // 2 x 64 x 128 B

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * BM;
  const int head = blockIdx.y;
  const long long tok0 = static_cast<long long>(blockIdx.z) * ntok;
  const bf16* Qb = Q + tok0 * ld + head * HD;
  const bf16* Kb = K + tok0 * ld + head * HD;
  const bf16* Vb = V + tok0 * ld + head * HD;

  // ---- async loads: Q tile + first K/V tile
  for (int i = tid; i < BM * 8; i += NT) {
    const int r = i >> 3, c = i & 7;
    const bool ok = q0 + r < ntok;
    cp_async16(sQ + tile_off(r, c), Qb + static_cast<long long>(ok ? q0 + r : 0) * ld + c * 8, ok);
  }
  auto load_kv = [&](int tile, int buf) {
    const int k0 = tile * FA_BN;
    for (int i = tid; i < FA_BN * 8; i += NT) {
      const int r = i >> 3, c = i & 7;
      const bool ok = k0 + r < ntok;
      const long long off = static_cast<long long>(ok ? k0 + r : 0) * ld + c * 8;
      cp_async16(sK + buf * FA_BN * 128 + tile_off(r, c), Kb + off, ok);
      cp_async16(sV + buf * FA_BN * 128 + tile_off(r, c), Vb + off, ok);
    }
  };
  load_kv(0, 0);
  cp_async_commit();

  const int nkv = (ntok + FA_BN - 1) / FA_BN;
  uint32_t qf[4][4];
  float o_acc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) o_acc[j][e] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};

  for (int it = 0; it < nkv; ++it) {
    const int buf = it & 1;
    if (it + 1 < nkv) {
      load_kv(it + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (it == 0) {
      // Q fragments for this warp's 16 rows: 4 k-steps of 16 along d
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int r = warp * 16 + (lane & 15);
        const int c = kk * 2 + (lane >> 4);
        ldsm_x4(smem_u32(sQ + tile_off(r, c)), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
      }
    }
    const uint8_t* kt = sK + buf * FA_BN * 128;
    const uint8_t* vt = sV + buf * FA_BN * 128;

    // ---- S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[j][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {  // pairs of key n-tiles
        const int mi = lane >> 3;
        const int key = jp * 16 + (mi >> 1) * 8 + (lane & 7);
        const int c = kk * 2 + (mi & 1);
        uint32_t b0, b1, b2, b3;
        ldsm_x4(smem_u32(kt + tile_off(key, c)), b0, b1, b2, b3);
        mma_bf16_16816(s[2 * jp], qf[kk], b0, b1);
        mma_bf16_16816(s[2 * jp + 1], qf[kk], b2, b3);
      }
    }
    // ---- mask keys beyond ntok (only possible in the last tile)
    const int kbase = it * FA_BN;
    if (kbase + FA_BN > ntok) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int key = kbase + j * 8 + 2 * t;
        if (key >= ntok) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
        if (key + 1 >= ntok) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      }
    }
    // ---- online softmax (rows g and g+8 of this warp's 16)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
    }
    float corr[2], msc[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float mnew = fmaxf(m_run[h], mx[h]);
      corr[h] = exp2f((m_run[h] - mnew) * scale_log2e);  // exp2(-inf) = 0 on the first tile
      m_run[h] = mnew;
      msc[h] = mnew * scale_log2e;
      l_run[h] *= corr[h];
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = exp2f(s[j][0] * scale_log2e - msc[0]);
      const float p1 = exp2f(s[j][1] * scale_log2e - msc[0]);
      const float p2 = exp2f(s[j][2] * scale_log2e - msc[1]);
      const float p3 = exp2f(s[j][3] * scale_log2e - msc[1]);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      const int kk = j >> 1;
      if ((j & 1) == 0) {
        pf[kk][0] = pack_bf16x2(p0, p1);
        pf[kk][1] = pack_bf16x2(p2, p3);
      } else {
        pf[kk][2] = pack_bf16x2(p0, p1);
        pf[kk][3] = pack_bf16x2(p2, p3);
      }
    }
    l_run[0] += rs[0];
    l_run[1] += rs[1];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o_acc[j][0] *= corr[0]; o_acc[j][1] *= corr[0];
      o_acc[j][2] *= corr[1]; o_acc[j][3] *= corr[1];
    }
    // ---- O += P V   (16 x 64 per warp; k = keys, n = d)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {  // pairs of d n-tiles
        const int mi = lane >> 3;
        const int key = kk * 16 + (mi & 1) * 8 + (lane & 7);
        const int c = jp * 2 + (mi >> 1);
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(smem_u32(vt + tile_off(key, c)), b0, b1, b2, b3);
        mma_bf16_16816(o_acc[2 * jp], pf[kk], b0, b1);
        mma_bf16_16816(o_acc[2 * jp + 1], pf[kk], b2, b3);
      }
    }
    __syncthreads();
  }

  // ---- finalize: quad-reduce row sums, normalise, store
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const int r0 = q0 + warp * 16 + g;
  bf16* Ob = O + tok0 * ldo + head * HD;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = j * 8 + 2 * t;
    if (r0 < ntok)
      *reinterpret_cast<uint32_t*>(Ob + static_cast<long long>(r0) * ldo + col) =
          pack_bf16x2(o_acc[j][0] * inv0, o_acc[j][1] * inv0);
    if (r0 + 8 < ntok)
      *reinterpret_cast<uint32_t*>(Ob + static_cast<long long>(r0 + 8) * ldo + col) =
          pack_bf16x2(o_acc[j][2] * inv1, o_acc[j][3] * inv1);
  }
}

// ------------------------------------------------------------------------------------------------
// temporal attention: grid = (pixels, halves b); block = nheads warps. Row of token (b, t, s) is
// (b*T + t)*S + s. Lane t < T owns query frame t.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxT = 32;
constexpr int kTaWarps = 4;
constexpr int kTaRowB = 144;  // 128 B of data + 16 B pad per staged row

__global__ void __launch_bounds__(kTaWarps * 32)
attn_temporal_kernel(const bf16* __restrict__ Q, const bf16* __restrict__ K, const bf16* __restrict__ V,
                     bf16* __restrict__ O, long long ld, long long ldo, int T, int S, int nheads,
                     float scale) {
  __shared__ __align__(16) uint8_t ta_smem[kTaWarps * 2 * kMaxT * kTaRowB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x, b = blockIdx.y;
  uint8_t* wk = ta_smem + static_cast<size_t>(warp) * 2 * kMaxT * kTaRowB;
  uint8_t* wv = wk + kMaxT * kTaRowB;
  const long long row_base = (static_cast<long long>(b) * T) * S + s;
  const bool active = lane < T;

  for (int head = warp; head < nheads; head += kTaWarps) {
    const long long col = static_cast<long long>(head) * HD;
    __syncwarp();
    for (int i = lane; i < T * 8; i += 32) {
      const int r = i >> 3, c = i & 7;
      const long long off = (row_base + static_cast<long long>(r) * S) * ld + col + c * 8;
      *reinterpret_cast<uint4*>(wk + r * kTaRowB + c * 16) = __ldg(reinterpret_cast<const uint4*>(K + off));
      *reinterpret_cast<uint4*>(wv + r * kTaRowB + c * 16) = __ldg(reinterpret_cast<const uint4*>(V + off));
    }
    float p[kMaxT];
    {
      float q[HD];
      const long long off = (row_base + static_cast<long long>(active ? lane : 0) * S) * ld + col;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(Q + off + c * 8));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16x2(w[j]);
          q[c * 8 + 2 * j] = f.x * scale;
          q[c * 8 + 2 * j + 1] = f.y * scale;
        }
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < kMaxT; ++j) {
        float acc = -INFINITY;
        if (j < T) {
          acc = 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 u = *reinterpret_cast<const uint4*>(wk + j * kTaRowB + c * 16);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = unpack_bf16x2(w[e]);
              acc = fmaf(q[c * 8 + 2 * e], f.x, acc);
              acc = fmaf(q[c * 8 + 2 * e + 1], f.y, acc);
            }
          }
        }
        p[j] = acc;
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kMaxT; ++j) mx = fmaxf(mx, p[j]);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxT; ++j) {
      p[j] = __expf(p[j] - mx);  // exp(-inf) = 0 for j >= T
      sum += p[j];
    }
    const float inv = 1.f / sum;
    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxT; ++j) {
      if (j < T) {
        // P is rounded to bf16 before P.V, matching the tensor-core attention kernels
        const float pj = __bfloat162float(__float2bfloat16_rn(p[j] * inv));
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 u = *reinterpret_cast<const uint4*>(wv + j * kTaRowB + c * 16);
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = unpack_bf16x2(w[e]);
            o[c * 8 + 2 * e] = fmaf(pj, f.x, o[c * 8 + 2 * e]);
            o[c * 8 + 2 * e + 1] = fmaf(pj, f.y, o[c * 8 + 2 * e + 1]);
          }
        }
      }
    }
    if (active) {
      bf16* op = O + (row_base + static_cast<long long>(lane) * S) * ldo + col;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        *reinterpret_cast<uint4*>(op + c * 8) =
            make_uint4(pack_bf16x2(o[c * 8], o[c * 8 + 1]), pack_bf16x2(o[c * 8 + 2], o[c * 8 + 3]),
                       pack_bf16x2(o[c * 8 + 4], o[c * 8 + 5]), pack_bf16x2(o[c * 8 + 6], o[c * 8 + 7]));
      }
    }
  }
}

}  // namespace v3d

using namespace v3d;

extern "C" {

/* Bring-up / cross-check implementation of v3d_attention_spatial on mma.sync (kept for validation of the
 * tcgen05 kernel in tests; the product path calls v3d_attention_spatial).
 * softmax(q k^T * scale) v per (sample, head), head dim 64; q/k/v are column slices of one packed
 * projection matrix (row stride ld_qkv), token rows sample-major. Replaces
 * F.scaled_dot_product_attention / xformers FMHA at sgm/modules/attention.py:337-341,432-444. */
int v3d_attention_spatial_mma(const void* q, const void* k, const void* v, void* o, int64_t ld_qkv, int64_t ld_o,
                          int32_t nbatch, int32_t ntok, int32_t nheads, float scale, void* stream) {
  if (!q || !k || !v || !o || ld_qkv % 8 != 0 || ld_o % 8 != 0 || nbatch <= 0 || ntok <= 0 || nheads <= 0 ||
      nheads > 65535 || nbatch > 65535) {
    set_error("v3d_attention_spatial_mma: bad args");
    return V3D_ERR_BAD_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float sl2 = scale * 1.44269504088896340736f;
  if (ntok >= 512) {
    constexpr int BM = 128;
    const int smem = BM * 128 + 4 * FA_BN * 128;
    static bool cfg = false;
    if (!cfg) {
      cudaFuncSetAttribute(attn_spatial_kernel<BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cfg = true;
    }
    dim3 grid((ntok + BM - 1) / BM, nheads, nbatch);
    attn_spatial_kernel<BM><<<grid, BM * 2, smem, st>>>(
        static_cast<const bf16*>(q), static_cast<const bf16*>(k), static_cast<const bf16*>(v),
        static_cast<bf16*>(o), ld_qkv, ld_o, ntok, sl2);
  } else {
    constexpr int BM = 64;
    const int smem = BM * 128 + 4 * FA_BN * 128;
    dim3 grid((ntok + BM - 1) / BM, nheads, nbatch);
    attn_spatial_kernel<BM><<<grid, BM * 2, smem, st>>>(
        static_cast<const bf16*>(q), static_cast<const bf16*>(k), static_cast<const bf16*>(v),
        static_cast<bf16*>(o), ld_qkv, ld_o, ntok, sl2);
  }
  V3D_CHECK_LAUNCH("attn_spatial_kernel");
  return V3D_OK;
}

/* Temporal self-attention across the T view-frames of each pixel (video_attention.py:114,125 around
 * attention.py:337-341), on the frame-major token matrix: row(b,t,s) = (b*T + t)*S + s. */
int v3d_attention_temporal(const void* q, const void* k, const void* v, void* o, int64_t ld_qkv, int64_t ld_o,
                           int32_t nb, int32_t T, int32_t S, int32_t nheads, float scale, void* stream) {
  if (!q || !k || !v || !o || ld_qkv % 8 != 0 || ld_o % 8 != 0 || T <= 0 || T > kMaxT || nheads <= 0 ||
      nb <= 0 || nb > 65535 || S <= 0) {
    set_error("v3d_attention_temporal: bad args (T=%d must be <= %d)", T, kMaxT);
    return V3D_ERR_BAD_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid(S, nb);
  attn_temporal_kernel<<<grid, kTaWarps * 32, 0, st>>>(
      static_cast<const bf16*>(q), static_cast<const bf16*>(k), static_cast<const bf16*>(v),
      static_cast<bf16*>(o), ld_qkv, ld_o, T, S, nheads, scale);
  V3D_CHECK_LAUNCH("attn_temporal_kernel");
  return V3D_OK;
}

}  // extern "C"
