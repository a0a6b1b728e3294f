// Spatial self-attention on the 5th-gen tensor cores (head dim 64, bf16 operands, fp32 softmax).
//
// One CTA = one (sample, head) x 128 queries; it streams 128-key K/V tiles:
//   warp 0      : TMA producer (Q once; K_j, V_j through a 2-stage mbarrier ring, 128B swizzle)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer
//                   S_j = Q K_j^T      (128x128x64,  A,B K-major)          -> TMEM cols [0,128)
//                   O_j = P_j V_j      (128x64x128,  A = P K-major in smem, B = V MN-major) -> TMEM cols 128+64*(j&1)
//   warps 2..5  : softmax: thread r owns query row r = its TMEM lane.  Two passes over the S row straight from
//                 TMEM (row max, then exp2 / row sum), P written as bf16 into the swizzled smem operand tile,
//                 O_{j-1} pulled from TMEM and folded into a register accumulator with the running-max correction.
// TMEM per CTA: 256 columns; shared memory ~112 KB -> two CTAs per SM interleave (one in softmax while the
// other's MMAs run), which is what hides the single-CTA S -> softmax -> PV dependency chain.
#include <type_traits>

#include "common.cuh"
#include "host_util.cuh"
#include "v3d_b200.h"

namespace v3d {

constexpr int AT_BM = 128;
constexpr int AT_BN = 128;
constexpr int AT_THREADS = 192;
constexpr int AT_TILE = 128 * 128;                 // bytes of a [128][64] bf16 tile
constexpr int AT_OFF_Q = 0;
constexpr int AT_OFF_K = AT_OFF_Q + AT_TILE;       // 2 stages
constexpr int AT_OFF_V = AT_OFF_K + 2 * AT_TILE;   // 2 stages
constexpr int AT_OFF_P = AT_OFF_V + 2 * AT_TILE;   // 2 atoms of 64 keys
constexpr int AT_OFF_BAR = AT_OFF_P + 2 * AT_TILE;
constexpr int AT_SMEM = AT_OFF_BAR + 256;
constexpr uint32_t AT_TMEM_COLS = 256;

__global__ void __launch_bounds__(AT_THREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
               const __grid_constant__ CUtensorMap mapV, bf16* __restrict__ O, long long ldo, int ntok,
               float scale_log2e) {
  extern __shared__ __align__(1024) uint8_t at_smem[];
  uint8_t* smem = at_smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AT_OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_ready = bars + 6;
  uint64_t* o_full = bars + 7;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BM;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int nkv = (ntok + AT_BN - 1) / AT_BN;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();  // the swizzled operand tiles need a 1024-byte aligned base
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&o_full[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, AT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;            // 128 columns
  const uint32_t tO = tmem_base + 128;      // 2 x 64 columns

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, AT_TILE);
      tma_load_3d(smem + AT_OFF_Q, &mapQ, q_full, head * 64, q0, b);
      for (int j = 0; j < nkv; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1u);
        mbar_arrive_expect_tx(&kv_full[st], 2 * AT_TILE);
        tma_load_3d(smem + AT_OFF_K + st * AT_TILE, &mapK, &kv_full[st], head * 64, j * AT_BN, b);
        tma_load_3d(smem + AT_OFF_V + st * AT_TILE, &mapV, &kv_full[st], head * 64, j * AT_BN, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);  // B (= V) is MN-major: d contiguous
      const uint64_t dq = umma_desc_k_sw128(smem_u32(smem + AT_OFF_Q));
      mbar_wait(q_full, 0);
      for (int j = 0; j < nkv; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&kv_full[st], ph);
        tc_fence_after();
        const uint64_t dk = umma_desc_k_sw128(smem_u32(smem + AT_OFF_K + st * AT_TILE));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_f16(tS, dq + static_cast<uint64_t>(2 * k), dk + static_cast<uint64_t>(2 * k), idesc_s, k != 0);
        tc_commit(s_full);
        mbar_wait(p_ready, j & 1);
        tc_fence_after();
        const uint32_t vbase = smem_u32(smem + AT_OFF_V + st * AT_TILE);
        const uint32_t d_o = tO + static_cast<uint32_t>((j & 1) * 64);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t dp = umma_desc_k_sw128(smem_u32(smem + AT_OFF_P + (kk >> 2) * AT_TILE)) +
                              static_cast<uint64_t>(2 * (kk & 3));
          const uint64_t dv = umma_desc_mn_sw128(vbase + kk * 2048, AT_TILE);
          tc_mma_f16(d_o, dp, dv, idesc_pv, kk != 0);
        }
        tc_commit(&o_full[j & 1]);
        tc_commit(&kv_empty[st]);
      }
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    uint8_t* prow = smem + AT_OFF_P + r * 128;
    const int sw = r & 7;
    float m_run = -INFINITY, l_run = 0.f;
    uint64_t acc[32];  // O accumulator, 64 fp32 as 32 packed pairs (FFMA2/FADD2/FMUL2 halve the issue slots)
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0ull;
    const uint64_t sl2 = pack2(scale_log2e, scale_log2e);

    // One 128-key tile. TAIL = the tile crosses ntok: keys beyond it are masked (only the last tile can be one).
    auto softmax_tile = [&](int j, auto tail_tag) {
      constexpr bool TAIL = decltype(tail_tag)::value;
      const int kbase = j * AT_BN;
      uint32_t va[32], vb[32];  // double-buffered TMEM chunks
      // ---- pass 1: row max (FMNMX3: two elements per instruction)
      float mx = -INFINITY;
      auto scan_max = [&](const uint32_t (&v)[32], int c4) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float a0 = __uint_as_float(v[i]), a1 = __uint_as_float(v[i + 1]);
          if (TAIL) {
            if (kbase + c4 * 32 + i >= ntok) a0 = -INFINITY;
            if (kbase + c4 * 32 + i + 1 >= ntok) a1 = -INFINITY;
          }
          mx = fmaxf(fmaxf(a0, a1), mx);
        }
      };
      tmem_ld32(tS + lane_base, va);
      tmem_ld_wait();
      tmem_ld32(tS + lane_base + 32u, vb);
      scan_max(va, 0);
      tmem_ld_wait();
      tmem_ld32(tS + lane_base + 64u, va);
      scan_max(vb, 1);
      tmem_ld_wait();
      tmem_ld32(tS + lane_base + 96u, vb);
      scan_max(va, 2);
      tmem_ld_wait();
      tmem_ld32(tS + lane_base, va);  // first chunk of pass 2, requested while the last max chunk is reduced
      scan_max(vb, 3);
      const float m_new = fmaxf(m_run, mx);
      const float corr = ex2_approx((m_run - m_new) * scale_log2e);  // 0 on the first tile
      const float msc = m_new * scale_log2e;
      m_run = m_new;
      tmem_ld_wait();
      // ---- fold O_{j-1} into the register accumulator (PV_{j-1} was issued before S_j, so it has completed)
      if (j > 0) {
        const int pj = j - 1;
        mbar_wait(&o_full[pj & 1], (pj >> 1) & 1);
        tc_fence_after();
        const uint64_t corr2 = pack2(corr, corr);
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          tmem_ld32(tO + lane_base + static_cast<uint32_t>((pj & 1) * 64 + c2 * 32), vb);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i)
            acc[c2 * 16 + i] = mul2(add2(acc[c2 * 16 + i], pack2(__uint_as_float(vb[2 * i]), __uint_as_float(vb[2 * i + 1]))), corr2);
        }
      }
      // ---- pass 2: P = exp2(s*scale - m), row sum, bf16 P into the swizzled smem A-operand tile
      const uint64_t nm2 = pack2(-msc, -msc);
      uint64_t lsum = 0ull;
      auto emit_p = [&](const uint32_t (&v)[32], int c4) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float x0, x1;
          unpack2(fma2(pack2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sl2, nm2), x0, x1);
          float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
          if (TAIL) {
            if (kbase + c4 * 32 + i >= ntok) p0 = 0.f;
            if (kbase + c4 * 32 + i + 1 >= ntok) p1 = 0.f;
          }
          lsum = add2(lsum, pack2(p0, p1));
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
        // keys c4*32 .. +31 -> atom (c4 >> 1), 16-byte chunks ((c4 & 1) * 4 + t), t = 0..3
        uint8_t* pa = prow + (c4 >> 1) * AT_TILE;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int chunk = (c4 & 1) * 4 + t;
          *reinterpret_cast<uint4*>(pa + ((chunk ^ sw) << 4)) =
              make_uint4(pk[4 * t], pk[4 * t + 1], pk[4 * t + 2], pk[4 * t + 3]);
        }
      };
      tmem_ld32(tS + lane_base + 32u, vb);
      emit_p(va, 0);
      tmem_ld_wait();
      tmem_ld32(tS + lane_base + 64u, va);
      emit_p(vb, 1);
      tmem_ld_wait();
      tmem_ld32(tS + lane_base + 96u, vb);
      emit_p(va, 2);
      tmem_ld_wait();
      emit_p(vb, 3);
      float l0, l1;
      unpack2(lsum, l0, l1);
      l_run = l_run * corr + (l0 + l1);
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core's async proxy
      tc_fence_before();
      mbar_arrive(p_ready);
    };

    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      if (j * AT_BN + AT_BN > ntok) softmax_tile(j, std::true_type{});
      else softmax_tile(j, std::false_type{});
    }
    // ---- last tile's O, normalise, store
    {
      const int pj = nkv - 1;
      mbar_wait(&o_full[pj & 1], (pj >> 1) & 1);
      tc_fence_after();
      const float inv = 1.0f / l_run;
      const int row = q0 + r;
      bf16* op = O + (static_cast<long long>(b) * ntok + row) * ldo + head * 64;
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        uint32_t v[32];
        tmem_ld32(tO + lane_base + static_cast<uint32_t>((pj & 1) * 64 + c2 * 32), v);
        tmem_ld_wait();
        if (row < ntok) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
              float a0, a1;
              unpack2(acc[(c2 * 32 + i + e) >> 1], a0, a1);
              o[e] = (a0 + __uint_as_float(v[i + e])) * inv;
              o[e + 1] = (a1 + __uint_as_float(v[i + e + 1])) * inv;
            }
            *reinterpret_cast<uint4*>(op + c2 * 32 + i) =
                make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]),
                           pack_bf16x2(o[6], o[7]));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, AT_TMEM_COLS);
  }
}

}  // namespace v3d

using namespace v3d;

extern "C" {

/* softmax(q k^T * scale) v per (sample, head), head dim 64, on tcgen05 tensor cores with TMA-staged operands.
 * q/k/v are column slices of one packed projection matrix (row stride ld_qkv elements), token rows sample-major.
 * Replaces F.scaled_dot_product_attention / xformers FMHA at sgm/modules/attention.py:337-341,432-444. */
int v3d_attention_spatial(const void* q, const void* k, const void* v, void* o, int64_t ld_qkv, int64_t ld_o,
                          int32_t nbatch, int32_t ntok, int32_t nheads, float scale, void* stream) {
  if (!q || !k || !v || !o || ld_qkv % 8 != 0 || ld_o % 8 != 0 || nbatch <= 0 || ntok <= 0 || nheads <= 0 ||
      nheads > 65535 || nbatch > 65535) {
    set_error("v3d_attention_spatial: bad args");
    return V3D_ERR_BAD_ARG;
  }
  CUtensorMap mq, mk, mv;
  const uint64_t dims[3] = {static_cast<uint64_t>(nheads) * 64, static_cast<uint64_t>(ntok),
                            static_cast<uint64_t>(nbatch)};
  const uint64_t str[2] = {static_cast<uint64_t>(ld_qkv) * 2, static_cast<uint64_t>(ld_qkv) * 2 * ntok};
  const uint32_t box[3] = {64, 128, 1};
  int rc;
  if ((rc = make_tmap_bf16(&mq, q, 3, dims, str, box))) return rc;
  if ((rc = make_tmap_bf16(&mk, k, 3, dims, str, box))) return rc;
  if ((rc = make_tmap_bf16(&mv, v, 3, dims, str, box))) return rc;
  static bool cfg = false;
  if (!cfg) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (e != cudaSuccess) {
      set_error("attn_tc smem attr: %s", cudaGetErrorString(e));
      return V3D_ERR_CUDA;
    }
    cfg = true;
  }
  dim3 grid((ntok + AT_BM - 1) / AT_BM, nheads, nbatch);
  attn_tc_kernel<<<grid, AT_THREADS, AT_SMEM, static_cast<cudaStream_t>(stream)>>>(
      mq, mk, mv, static_cast<bf16*>(o), ld_o, ntok, scale * 1.44269504088896340736f);
  V3D_CHECK_LAUNCH("attn_tc_kernel");
  return V3D_OK;
}

}  // extern "C"
