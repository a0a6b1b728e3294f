// tcgen05 GEMM / temporal-conv / implicit-GEMM 3x3 conv for sm_100a.
//
// One persistent, warp-specialised kernel:
//   warp 0      : TMA producer  (A tile 128x64 bf16 + B tile BNx64 bf16 per pipeline stage, 128B swizzle)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (128 x BN x 16 per instruction)
//   warps 2..9  : epilogue (tcgen05.ld 32x32b -> bias / per-frame bias / SiLU / GEGLU / residual blend ->
//                 bf16 or fp32 stores). Two TMEM accumulator stages so the epilogue of tile i overlaps
//                 the main loop of tile i+1.
// Operand gather modes (see include/v3d_b200.h): linear rows, 3-tap temporal shift, 3x3 spatial taps.
// The im2row never exists in memory: each (tap, 64-channel) K-block is one TMA box whose out-of-bounds
// part is zero-filled by the hardware, which is exactly the conv zero padding.
#include <cstdlib>

#include "common.cuh"
#include "host_util.cuh"
#include "v3d_b200.h"

namespace v3d {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int kSmemBudget = 227 * 1024;

struct GemmEpi {
  void* D;
  const float* bias;
  const float* fbias;
  const bf16* R1;
  const bf16* R2;
  long long ldd, ldr1, ldr2, ldfb;
  int batch, rows_per_batch, tiles_per_batch;
  int N, kpt, ntaps, tap_shift;  // kpt = K-blocks per tap
  int rows_per_frame, act, out_fp32;
  int b_batched;
  int fb_uniform;  // every 32-row warp slice of a tile lies in one frame: per-frame bias folds into the bias registers
  long long* trace;  // diagnostics only: CTA 0 writes role timelines (clock64) here when non-null
  // diagnostics only (V3D_GEMM_DEBUG): 1 = skip the store path, 2 = skip TMA store issue, 4 = skip TMEM loads,
  // 8 = skip the drain wait, 16 = skip the proxy fence, 32 = skip the smem writes.  Results are wrong with any bit set.
  int dbg;
  int transposed, valid_cols, accumulate;  // small-M mode: D is fp32 [cols][ldd], D[col][row]; bias per row
  // conv3x3 geometry
  int cn, ch, cw, bw, bh, bn, tiles_w, tiles_h, bw_shift, bh_shift, tw_shift, th_shift;  // t*_shift < 0: not pow2
  int num_m_tiles, num_n_tiles;
  float s0, s1, s2;
  int tap0;  // A row coordinate of tap 0 relative to the output row: a_row0 - (ntaps / 2) * tap_shift
};

// Fused GEMM -> all-gather over peer memory: output sub-tiles whose first column is >= col0 are stored a second time
// through each of these tensor maps (destination matrices that may live on other GPUs: IPC-mapped arenas), by the same
// store leader with the same cp.async.bulk.tensor instruction.  n = 0: off (every launch but the frame-sharded K|V
// projection).
struct alignas(64) KvMaps {
  CUtensorMap m[8];
  int n;
  int col0;
};

// Diagnostics (role timelines through v3d_debug_set_trace, V3D_GEMM_DEBUG stage-skipping switches) are compiled in
// only with -DV3D_GEMM_DIAG: the hooks cost registers and ~6% of the epilogue's issue slots.
#ifdef V3D_GEMM_DIAG
#define V3D_DIAG(x) (x)
#define V3D_DBG(bit) (p.dbg & (bit))
#else
#define V3D_DIAG(x) (false)
#define V3D_DBG(bit) (0)
#endif

// NCTA = 1: one CTA per 128-row tile.  NCTA = 2: a CTA pair (cluster of 2, cta_group::2) computes a 256 x BN tile;
// each CTA stages its own 128 rows of A and HALF of the B tile, the leader's tensor core reads both halves.
template <int BN, int NCTA = 1, bool RT = false>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / NCTA) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // 4 KB per epilogue warp: 32 rows x 128 B, swizzled; RT adds a ring of RDEPTH residual sub-tiles (8 KB each) per group
  static constexpr int OUT_STAGING_BYTES = 8 * 4096;
  static constexpr int RDEPTH = 3;
  static constexpr int STAGING_BYTES = OUT_STAGING_BYTES + (RT ? 2 * RDEPTH * 8192 : 0);
  static constexpr int NSTAGE_RAW = (kSmemBudget - 2048 - STAGING_BYTES) / STAGE_BYTES;
  static constexpr int NSTAGE = NSTAGE_RAW > 8 ? 8 : NSTAGE_RAW;
  // accumulator ring in TMEM: as many stages as fit in the 512 columns (max 4). Deeper rings let the MMA issuer run
  // further ahead of the epilogue, which hides the tile-to-tile signalling round trip on short-K problems.
  static constexpr int ACC_STRIDE = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : (BN <= 160 ? 160 : 256)));
  static constexpr int ACC_STAGES = (512 / ACC_STRIDE) > 4 ? 4 : (512 / ACC_STRIDE);
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + STAGING_BYTES + 1024 /*align slack*/ + (RT ? 512 : 256) /*barriers*/;
};

// epilogue variants (compile-time): what is stored and how
constexpr int EPI_BF16 = 0;   // bf16 sub-tiles staged in shared memory, written by TMA; at most one residual (R1)
constexpr int EPI_BF16R2 = 4; // as EPI_BF16 with both residuals (blend epilogue); the only variant that carries R2
constexpr int EPI_BF16RT = 5; // as EPI_BF16 with R1 always present and staged by TMA: the residual sub-tile (128 rows x 32
                              // columns) lands in shared memory two sub-tiles ahead of its use instead of arriving
                              // through per-lane global loads one load group ahead (opt-in: V3D_GEMM_RTMA=1)
constexpr int EPI_GEGLU = 1;  // value * gelu(gate) then as EPI_BF16 (linear mode only)
constexpr int EPI_F32 = 2;    // fp32 direct stores (optional SiLU)
constexpr int EPI_TRANS = 3;  // fp32 transposed store with per-row bias (small-M mode, linear only)

// Persistent tile walk without per-tile divisions: tile = first + i * step, n fastest.  With CTA pairs the walk is
// over pair tiles (m_tile = pair index; the CTA's own 128-row tile is 2 * m_tile + rank).
struct TileWalk {
  int n_tile, m_tile, step_n, step_m, num_n;
  __device__ __forceinline__ TileWalk(const GemmEpi& p, int ncta) {
    num_n = p.num_n_tiles;
    n_tile = (static_cast<int>(blockIdx.x) / ncta) % num_n;
    m_tile = (static_cast<int>(blockIdx.x) / ncta) / num_n;
    step_n = (static_cast<int>(gridDim.x) / ncta) % num_n;
    step_m = (static_cast<int>(gridDim.x) / ncta) / num_n;
  }
  __device__ __forceinline__ void next() {
    n_tile += step_n;
    m_tile += step_m;
    if (n_tile >= num_n) {
      n_tile -= num_n;
      ++m_tile;
    }
  }
};

// m_tile -> tile origin. conv: (w0, h0, img0); linear: (batch index, first row, 0)
template <bool CONV>
__device__ __forceinline__ void tile_origin(const GemmEpi& p, int m_tile, int& t0, int& t1, int& t2) {
  if (CONV) {
    int tw, th, tn;
    if (p.tw_shift >= 0) {
      tw = m_tile & (p.tiles_w - 1);
      th = (m_tile >> p.tw_shift) & (p.tiles_h - 1);
      tn = m_tile >> (p.tw_shift + p.th_shift);
    } else {
      tw = m_tile % p.tiles_w;
      th = (m_tile / p.tiles_w) % p.tiles_h;
      tn = m_tile / (p.tiles_w * p.tiles_h);
    }
    t0 = tw * p.bw;
    t1 = th * p.bh;
    t2 = tn * p.bn;
  } else {
    if (p.batch == 1) {
      t0 = 0;
      t1 = m_tile * BM;
    } else {
      t0 = m_tile / p.tiles_per_batch;
      t1 = (m_tile - t0 * p.tiles_per_batch) * BM;
    }
    t2 = 0;
  }
}

template <int BN, bool CONV, int EPI, int NCTA = 1>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapD, const GemmEpi p, const __grid_constant__ CUtensorMap mapR,
               const __grid_constant__ KvMaps kv) {
  constexpr bool RTMA = EPI == EPI_BF16RT;
  static_assert(!RTMA || NCTA == 1, "the TMA-staged residual variant is single-CTA");
  using C = Cfg<BN, NCTA, RTMA>;
  static_assert(NCTA == 1 || NCTA == 2, "one CTA or a CTA pair");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* staging = smem + C::NSTAGE * C::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + C::STAGING_BYTES);
  uint64_t* empty_bar = full_bar + C::NSTAGE;
  uint64_t* tfull_bar = empty_bar + C::NSTAGE;
  uint64_t* tempty_bar = tfull_bar + 4;
  uint64_t* rfull_bar = tempty_bar + 4;  // RTMA only: [group][ring slot], one TMA transaction each
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 4 + (RTMA ? 2 * C::RDEPTH : 0));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    if (EPI == EPI_BF16 || EPI == EPI_BF16R2 || EPI == EPI_GEGLU || RTMA) tma_prefetch_desc(&mapD);
    if (RTMA) {
      tma_prefetch_desc(&mapR);
      for (int i = 0; i < 2 * C::RDEPTH; ++i) mbar_init(&rfull_bar[i], 1);
    }
    for (int i = 0; i < C::NSTAGE; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < C::ACC_STAGES; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8 * NCTA);  // pair: the leader's barrier collects both CTAs' epilogue warps
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (NCTA == 2) {
      tmem_alloc_2cta(tmem_slot, C::TMEM_COLS);
      tmem_relinquish_2cta();
    } else {
      tmem_alloc(tmem_slot, C::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (NCTA == 2) cluster_sync_all();  // both CTAs' barriers exist before any remote arrive / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // rank inside the pair; the persistent walk is over pair tiles: first = blockIdx.x / NCTA, step = gridDim.x / NCTA
  const int cta_rank = NCTA == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  const int total_tiles = ((p.num_m_tiles + NCTA - 1) / NCTA) * p.num_n_tiles;
  const int num_kb = p.ntaps * p.kpt;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int tr_n = 0;
      TileWalk tw(p, NCTA);
      for (int tile = blockIdx.x / NCTA; tile < total_tiles; tile += gridDim.x / NCTA, tw.next()) {
        const int n_tile = tw.n_tile;
        const int m_tile = tw.m_tile * NCTA + cta_rank;
        int c1, c2, c3;
        if (CONV) {
          tile_origin<true>(p, m_tile, c1, c2, c3);
        } else {
          tile_origin<false>(p, m_tile, c2, c1, c3);
        }
        const int bz = p.b_batched ? c2 : 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (V3D_DIAG(p.trace && blockIdx.x == 0 && tr_n < 1024)) p.trace[tr_n++] = clock64();
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          // pair: only the leader arms its barrier, with the bytes of both CTAs; the peer's loads credit it remotely
          if (NCTA == 1 || cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], NCTA * C::STAGE_BYTES);
          const int tap = kb / p.kpt;
          const int kc = (kb - tap * p.kpt) * BK;
          if (NCTA == 2) {
            if (CONV) {
              const int ky = tap / 3, kx = tap - ky * 3;
              tma_load_4d_2cta(sa, &mapA, &full_bar[stage], kc, c1 + kx - 1, c2 + ky - 1, c3);
            } else {
              tma_load_3d_2cta(sa, &mapA, &full_bar[stage], kc, c1 + tap * p.tap_shift + p.tap0, c2);
            }
            // this CTA's half of the B tile: rows [rank * BN/2, (rank + 1) * BN/2) of the N tile
            tma_load_3d_2cta(sb, &mapB, &full_bar[stage], kb * BK, n_tile * BN + cta_rank * (BN / 2), bz);
          } else {
            if (CONV) {
              const int ky = tap / 3, kx = tap - ky * 3;
              tma_load_4d(sa, &mapA, &full_bar[stage], kc, c1 + kx - 1, c2 + ky - 1, c3);
            } else {
              tma_load_3d(sa, &mapA, &full_bar[stage], kc, c1 + tap * p.tap_shift + p.tap0, c2);
            }
            tma_load_3d(sb, &mapB, &full_bar[stage], kb * BK, n_tile * BN, bz);
          }
          if (++stage == C::NSTAGE) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0 && (NCTA == 1 || cta_rank == 0)) {  // pair: the leader issues for both SMs
      constexpr uint32_t idesc = umma_idesc_bf16(BM * NCTA, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int tr_m = 0;
      for (int tile = blockIdx.x / NCTA; tile < total_tiles; tile += gridDim.x / NCTA) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        if (V3D_DIAG(p.trace && blockIdx.x == 0 && tr_m < 1020)) p.trace[1024 + tr_m++] = -clock64();  // negative: tile start
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * C::ACC_STRIDE);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (V3D_DIAG(p.trace && blockIdx.x == 0 && tr_m < 1020)) p.trace[1024 + tr_m++] = clock64();
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t adesc = umma_desc_k_sw128(sa);
          const uint64_t bdesc = umma_desc_k_sw128(sa + C::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 bytes per 16-element K step inside the 128B swizzle atom (encoded >>4)
            if (NCTA == 2)
              tc_mma_f16_2cta(d_tmem, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2),
                              idesc, (kb | k) != 0 ? 1u : 0u);
            else
              tc_mma_f16(d_tmem, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2),
                         idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if (NCTA == 2) tc_commit_2cta(&empty_bar[stage], 3);  // frees the stage in both CTAs
          else tc_commit(&empty_bar[stage]);
          if (++stage == C::NSTAGE) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (NCTA == 2) tc_commit_2cta(&tfull_bar[acc], 3);  // each CTA's epilogue reads its own 128 rows
        else tc_commit(&tfull_bar[acc]);
        if (++acc == C::ACC_STAGES) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ------------------------------ epilogue (warps 2..9) ------------------------------
    // Two warps per TMEM lane quarter; each takes half of the tile's 16-column chunks.  The variant (EPI) is a
    // template parameter and the chunk loop is fully unrolled, so the per-chunk code is straight-line: TMEM load,
    // bias/residual loads issued underneath it, math, then either the swizzled smem slab (32 rows x 128 B per
    // warp, flushed as row-contiguous 16-byte stores: 4 full 128-byte lines per warp instruction) or direct fp32.
    constexpr bool GEGLU = EPI == EPI_GEGLU;
    constexpr bool STAGED = EPI == EPI_BF16 || EPI == EPI_BF16R2 || EPI == EPI_GEGLU || RTMA;
    constexpr bool DUAL = EPI == EPI_BF16R2 || EPI == EPI_F32;  // variants that may take a second residual
    constexpr int OUT_COLS = GEGLU ? BN / 2 : BN;
    constexpr int NCH = OUT_COLS / 16;
    // bf16 output leaves as whole-tile-height TMA stores: a sub-tile is 128 rows x SUBW columns, staged by the four
    // warps of a group (one per TMEM lane quarter) in one of the group's two buffers and written by one
    // cp.async.bulk.tensor per sub-tile.  Group 0 (warps 2..5) takes the first half of the sub-tiles, group 1 the rest.
    constexpr int SUBW = (STAGED && OUT_COLS % 32 == 0) ? 32 : 16;
    constexpr int CPS = SUBW / 16;  // 16-column chunks per sub-tile
    constexpr int NSUB = OUT_COLS / SUBW;
    constexpr int CH_HALF = ((NSUB + 1) / 2) * CPS;
    constexpr int SUB_BYTES = BM * SUBW * 2;
    static_assert(2 * 2 * SUB_BYTES <= C::OUT_STAGING_BYTES, "staging too small");
    static_assert(!RTMA || SUBW == 32, "the TMA-staged residual needs 32-column sub-tiles");
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;   // 0: warps 2..5, 1: warps 6..9
    const int r = q * 32 + lane;
    const uint32_t stg = smem_u32(staging) + static_cast<uint32_t>(half * 2 * SUB_BYTES);  // this group's buffers
    // RTMA: this group's ring of residual buffers (same 128 x 64 B, 64B-swizzled layout as the output buffers) and
    // their barriers; sub-tile number n (the group's running count `nstore`) lives in slot n % RD, phase (n / RD) & 1
    constexpr uint32_t RD = C::RDEPTH;
    uint8_t* const rbuf = staging + C::OUT_STAGING_BYTES + half * static_cast<int>(RD) * SUB_BYTES;
    uint64_t* const rfull = rfull_bar + half * RD;
    const bool store_leader = q == 0 && lane == 0;
    const uint32_t row_off = static_cast<uint32_t>(r * SUBW * 2);
    const uint32_t swz = SUBW == 32 ? static_cast<uint32_t>((r >> 1) & 3) : 0u;  // 64B-swizzle XOR of the 16-byte chunk
    const bool has_bias = p.bias != nullptr;
    // the GEGLU variant carries no residual / per-frame bias code (rejected on the host): registers
    const bool has_fb = !GEGLU && p.fbias != nullptr;
    const bool has_r1 = !GEGLU && p.R1 != nullptr;
    const bool has_r2 = DUAL && p.R2 != nullptr;
    const bool do_silu = p.act == V3D_ACT_SILU;
    const float s0 = p.s0, s1 = p.s1, s2 = p.s2;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t nstore = 0;  // sub-tiles this group has stored (buffer ring position)
    int tr_e = 0;
    const bool tracing = V3D_DIAG(p.trace && blockIdx.x == 0 && warp == 4 && lane == 0);  // group 0's store leader
#define V3D_ETRACE() do { if (V3D_DIAG(tracing && tr_e < 1020)) p.trace[2048 + tr_e++] = clock64(); } while (0)
    // TMEM is read in load groups, double-buffered: while one group is processed the next is in flight.  A group is
    // a whole sub-tile (32 columns, one x32 load) for bf16 output, value+gate (2 x16) for GEGLU, one x16 otherwise.
    constexpr int LG = GEGLU ? 1 : CPS;          // 16-column output chunks per load group
    constexpr int LREGS = GEGLU ? 32 : 16 * LG;  // registers per buffer
    constexpr int NG_HALF = CH_HALF / LG;
    constexpr int NBR = (CH_HALF * 16 + 31) / 32;  // bias registers per lane
    // tile row -> (valid, global row). bw/bh are powers of two (checked on the host).
    auto map_row_of = [&](int tr, int o0, int o1, int o2, long long& grow) -> bool {
      if (CONV) {
        const int w = tr & (p.bw - 1);
        const int h = (tr >> p.bw_shift) & (p.bh - 1);
        const int img = o2 + (tr >> (p.bw_shift + p.bh_shift));
        grow = (static_cast<long long>(img) * p.ch + (o1 + h)) * p.cw + (o0 + w);
        return img < p.cn;
      } else {
        const int m = o1 + tr;
        grow = static_cast<long long>(o0) * p.rows_per_batch + m;
        // pair: with an odd number of 128-row tiles the peer's last tile lies past the last batch item (its TMA
        // loads are zero-filled and its stores clipped; its residual / per-frame-bias rows must not be touched)
        return m < p.rows_per_batch && (NCTA == 1 || o0 < p.batch);
      }
    };
    TileWalk tw(p, NCTA);
    int iter = 0;
    for (int tile = blockIdx.x / NCTA; tile < total_tiles; tile += gridDim.x / NCTA, ++iter) {
      V3D_ETRACE();  // [0] tile prologue start
      const int n_tile = tw.n_tile;
      int t0, t1, t2;
      tile_origin<CONV>(p, tw.m_tile * NCTA + cta_rank, t0, t1, t2);
      tw.next();  // tw now names the next tile (used by the residual prefetch below)
      // with an odd number of sub-tiles the two groups swap the larger share every tile, so neither paces the other
      const int hsel = (NSUB & 1) ? (half ^ (iter & 1)) : half;
      const int c_begin = hsel ? CH_HALF : 0;
      const int my_n = hsel ? NCH - CH_HALF : CH_HALF;
      const int c_begin_next = (NSUB & 1) ? (hsel ? 0 : CH_HALF) : c_begin;

      // accumulator first: the first TMEM load is in flight while the row bookkeeping below is computed
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      V3D_ETRACE();  // [1] accumulator ready
      const uint32_t t_acc =
          tmem_base + static_cast<uint32_t>(acc * C::ACC_STRIDE) + (static_cast<uint32_t>(q * 32) << 16);
      uint32_t vb[2][LREGS];
      auto issue_group = [&](int g, uint32_t* dst) {
        if (V3D_DBG(4)) return;
        const uint32_t c = static_cast<uint32_t>((c_begin + g * LG) * 16);
        if (GEGLU) {
          tmem_ld16p(t_acc + c, dst);
          tmem_ld16p(t_acc + static_cast<uint32_t>(BN / 2) + c, dst + 16);
        } else if (LG == 2) {
          tmem_ld32p(t_acc + c, dst);
        } else {
          tmem_ld16p(t_acc + c, dst);
        }
      };
      if (my_n > 0) issue_group(0, vb[0]);
      // bias for this warp's columns, one coalesced load per tile, spread over the lanes (column i*32+lane of the
      // warp's slice) and handed out by shuffles in the chunk loop: no memory latency inside the loop.  A per-frame
      // bias that is uniform over the warp's rows is folded in here.
      float breg[NBR], greg[GEGLU ? NBR : 1];
      {
        const float* fbw = nullptr;
        if (has_fb && p.fb_uniform) {
          long long rw;
          if (map_row_of(q * 32, t0, t1, t2, rw))
            fbw = p.fbias + static_cast<long long>(static_cast<int>(rw) / p.rows_per_frame) * p.ldfb + n_tile * BN;
        }
#pragma unroll
        for (int i = 0; i < NBR; ++i) {
          const int col = c_begin * 16 + i * 32 + lane;
          const bool ok = i * 32 + lane < my_n * 16;
          float bvv = 0.f;
          if (EPI != EPI_TRANS && has_bias && ok) bvv = __ldg(p.bias + n_tile * BN + col);
          if (fbw && ok) bvv += __ldg(fbw + col);
          breg[i] = bvv;
          if (GEGLU) greg[i] = (has_bias && ok) ? __ldg(p.bias + n_tile * BN + BN / 2 + col) : 0.f;
        }
      }

      // tile row -> (valid, global row). bw/bh are powers of two (checked on the host).
      auto map_row = [&](int o0, int o1, int o2, long long& grow) -> bool {
        return map_row_of(r, o0, o1, o2, grow);
      };
      const int obase = n_tile * OUT_COLS;  // first output column of this tile
      const int nbase = n_tile * BN;        // first column in the (packed) N space
      long long row = 0;
      bool valid = true;
      const float* fb = nullptr;
      const bf16* r1p = nullptr;
      const bf16* r2p = nullptr;
      float row_bias = 0.f;
      if (!STAGED || (has_fb && !p.fb_uniform) || has_r1 || has_r2) {  // plain bf16 tiles never need per-row addresses
        valid = map_row(t0, t1, t2, row);
        if (has_fb && !p.fb_uniform && valid) fb = p.fbias + static_cast<long long>(static_cast<int>(row) / p.rows_per_frame) * p.ldfb + nbase;
        if (has_r1 && valid) r1p = p.R1 + row * p.ldr1 + obase;
        if (has_r2 && valid) r2p = p.R2 + row * p.ldr2 + obase;
        if (EPI == EPI_TRANS && has_bias && valid) row_bias = __ldg(p.bias + row);
      }
      if (has_r1 || has_r2) {
        // pull the residual row segments of the tile after next into L2 now (two tile-times of lead: these rows
        // come from HBM); the first iteration also covers the next tile
        auto prefetch_tile = [&](const TileWalk& w, int cb) {
          int u0, u1, u2;
          tile_origin<CONV>(p, w.m_tile * NCTA + cta_rank, u0, u1, u2);
          long long nrow;
          if (map_row(u0, u1, u2, nrow)) {
            const int ncol0 = w.n_tile * OUT_COLS + cb * 16;
#pragma unroll
            for (int off = 0; off < CH_HALF * 16; off += 64) {
              if (has_r1) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.R1 + nrow * p.ldr1 + ncol0 + off));
              if (has_r2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.R2 + nrow * p.ldr2 + ncol0 + off));
            }
          }
        };
        const int step = static_cast<int>(gridDim.x) / NCTA;
        if (iter == 0 && tile + step < total_tiles) prefetch_tile(tw, c_begin_next);
        if (tile + 2 * step < total_tiles) {
          TileWalk tw2 = tw;
          tw2.next();
          prefetch_tile(tw2, c_begin);
        }
      }

      // residuals travel one load group ahead: group g+1's rows are requested before group g is processed, so they
      // have landed by the time the proxy fence (a full CTA memory barrier) of group g is reached
      uint4 na[LG][2];
#pragma unroll
      for (int hh = 0; hh < LG; ++hh) na[hh][0] = na[hh][1] = make_uint4(0, 0, 0, 0);
      auto load_res = [&](int g) {
        if (RTMA) return;  // the residual sub-tiles arrive through TMA (below)
#pragma unroll
        for (int hh = 0; hh < LG; ++hh) {
          const int c = (c_begin + g * LG + hh) * 16;
          if (r1p) { na[hh][0] = __ldg(reinterpret_cast<const uint4*>(r1p + c)); na[hh][1] = __ldg(reinterpret_cast<const uint4*>(r1p + c) + 1); }
        }
      };
      // RTMA: the group's store leader requests residual sub-tile `sub` (index within this tile's share) into the
      // buffer the sub-tile counter selects; the TMA clips at row / image tails like the output store (zero fill)
      auto request_res = [&](int col0, int o0, int o1, int o2, uint32_t slot) {
        const uint32_t sl = slot % RD;
        uint8_t* dst = rbuf + sl * SUB_BYTES;
        mbar_arrive_expect_tx(&rfull[sl], SUB_BYTES);
        if (CONV) tma_load_4d(dst, &mapR, &rfull[sl], col0, o0, o1, o2);
        else tma_load_3d(dst, &mapR, &rfull[sl], col0, o1, o0);
      };
      // the first RD sub-tiles of a tile are requested at the end of the previous tile (below); only the very first
      // tile of this CTA requests its own
      if (RTMA && store_leader && iter == 0) {
#pragma unroll
        for (uint32_t j = 0; j < RD; ++j)
          if (static_cast<int>(j) * LG < my_n)
            request_res(obase + (c_begin + static_cast<int>(j) * LG) * 16, t0, t1, t2, nstore + j);
      }
      if (my_n > 0) load_res(0);
#pragma unroll
      for (int g = 0; g < NG_HALF; ++g) {
        if (g * LG < my_n) {
          tmem_ld_wait();
          V3D_ETRACE();  // group: TMEM data landed
          if ((g + 1) * LG < my_n) issue_group(g + 1, vb[(g + 1) & 1]);
          uint4 ca[LG][2], cb[LG][2];
          if (RTMA) mbar_wait(&rfull[nstore % RD], (nstore / RD) & 1u);  // this sub-tile's residual rows have landed
#pragma unroll
          for (int hh = 0; hh < LG; ++hh) {
            if (RTMA) {
              const uint32_t rrow = smem_u32(rbuf) + (nstore % RD) * SUB_BYTES + row_off;
              ca[hh][0] = lds128(rrow + (((hh * 2) ^ swz) << 4));
              ca[hh][1] = lds128(rrow + (((hh * 2 + 1) ^ swz) << 4));
            } else {
              ca[hh][0] = na[hh][0];
              ca[hh][1] = na[hh][1];
            }
            cb[hh][0] = cb[hh][1] = make_uint4(0, 0, 0, 0);
            // the second residual (blend epilogue only) is fetched for the current group: its rows were pulled
            // into L2 two tiles ago, and holding a second prefetched set would spill
            if (r2p) {
              const int c = (c_begin + g * LG + hh) * 16;
              cb[hh][0] = __ldg(reinterpret_cast<const uint4*>(r2p + c));
              cb[hh][1] = __ldg(reinterpret_cast<const uint4*>(r2p + c) + 1);
            }
          }
          if ((g + 1) * LG < my_n) load_res(g + 1);
#pragma unroll
          for (int hh = 0; hh < LG; ++hh) {
            const int k = g * LG + hh;
            const int c = (c_begin + k) * 16;
            const uint32_t* v = &vb[g & 1][GEGLU ? 0 : hh * 16];
            const uint32_t* gv = &vb[g & 1][GEGLU ? 16 : 0];
            const uint4 ra0 = ca[hh][0], ra1 = ca[hh][1], rb0 = cb[hh][0], rb1 = cb[hh][1];
            // all epilogue arithmetic runs on packed f32x2 pairs (FADD2 / FMUL2 / FFMA2: half the issue slots)
            uint64_t f2[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f2[j] = pack2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
            // bias vectors are shared by every row: after the first warp touched them they are L1 hits
            if (EPI == EPI_TRANS) {
              const uint64_t rb2 = pack2(row_bias, row_bias);
#pragma unroll
              for (int j = 0; j < 8; ++j) f2[j] = add2(f2[j], rb2);
            } else if (has_bias || (has_fb && p.fb_uniform)) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float b0 = __shfl_sync(0xffffffffu, breg[(k * 16 + 2 * j) >> 5], (k * 16 + 2 * j) & 31);
                const float b1 = __shfl_sync(0xffffffffu, breg[(k * 16 + 2 * j + 1) >> 5], (k * 16 + 2 * j + 1) & 31);
                f2[j] = add2(f2[j], pack2(b0, b1));
              }
            }
            if (fb) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(fb + c) + j);
                f2[2 * j] = add2(f2[2 * j], pack2(b4.x, b4.y));
                f2[2 * j + 1] = add2(f2[2 * j + 1], pack2(b4.z, b4.w));
              }
            }
            if (GEGLU) {
              uint64_t g2[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) g2[j] = pack2(__uint_as_float(gv[2 * j]), __uint_as_float(gv[2 * j + 1]));
              if (has_bias) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float b0 = __shfl_sync(0xffffffffu, greg[(k * 16 + 2 * j) >> 5], (k * 16 + 2 * j) & 31);
                  const float b1 = __shfl_sync(0xffffffffu, greg[(k * 16 + 2 * j + 1) >> 5], (k * 16 + 2 * j + 1) & 31);
                  g2[j] = add2(g2[j], pack2(b0, b1));
                }
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) f2[j] = mul2(f2[j], gelu_erf_fast2(g2[j]));
            } else if (EPI == EPI_F32 && do_silu) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float a, b;
                unpack2(f2[j], a, b);
                f2[j] = pack2(silu_f(a), silu_f(b));
              }
            }
            if (r1p || r2p) {
              const uint64_t s0v = pack2(s0, s0), s1v = pack2(s1, s1), s2v = pack2(s2, s2);
              if (r1p) {
                const uint32_t u[8] = {ra0.x, ra0.y, ra0.z, ra0.w, ra1.x, ra1.y, ra1.z, ra1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float2 x2 = unpack_bf16x2(u[j]);
                  f2[j] = fma2(f2[j], s0v, mul2(pack2(x2.x, x2.y), s1v));
                }
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) f2[j] = mul2(f2[j], s0v);
              }
              if (r2p) {
                const uint32_t u[8] = {rb0.x, rb0.y, rb0.z, rb0.w, rb1.x, rb1.y, rb1.z, rb1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float2 x2 = unpack_bf16x2(u[j]);
                  f2[j] = fma2(pack2(x2.x, x2.y), s2v, f2[j]);
                }
              }
            } else if (s0 != 1.0f) {
              const uint64_t s0v = pack2(s0, s0);
#pragma unroll
              for (int j = 0; j < 8; ++j) f2[j] = mul2(f2[j], s0v);
            }
            float f[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) unpack2(f2[j], f[2 * j], f[2 * j + 1]);
            if (EPI == EPI_TRANS) {
              // D[col][row]: lanes hold consecutive rows -> each store instruction is one contiguous 128-byte line
              if (valid) {
                float* dp = static_cast<float*>(p.D) + row;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  if (obase + c + j < p.valid_cols) {
                    float* o = dp + static_cast<long long>(obase + c + j) * p.ldd;
                    *o = p.accumulate ? *o + f[j] : f[j];
                  }
                }
              }
            } else if (EPI == EPI_F32) {
              if (valid) {
                float4* dp = reinterpret_cast<float4*>(static_cast<float*>(p.D) + row * p.ldd + obase + c);
#pragma unroll
                for (int j = 0; j < 4; ++j) dp[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
              }
            } else if (!V3D_DBG(1)) {
              // bf16 rows leave through TMA.  The chunk goes into this row's slot of the group's current buffer
              // (64-byte rows, 64B-swizzled: conflict-free 16-byte stores); when the sub-tile is complete the
              // group meets on its named barrier and one thread issues the 128-row store, which the hardware
              // clips at the tensor edge (row tails, image tails).
              const uint32_t buf = stg + static_cast<uint32_t>((nstore & 1) * SUB_BYTES) + row_off;
              const int part = k % CPS;
              if (!V3D_DBG(32)) {
                sts128(buf + (((part * 2) ^ swz) << 4), pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                       pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
                sts128(buf + (((part * 2 + 1) ^ swz) << 4), pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]),
                       pack_bf16x2(f[12], f[13]), pack_bf16x2(f[14], f[15]));
              }
              if (part == CPS - 1) {
                if (!V3D_DBG(16)) fence_proxy_async_smem();
                V3D_ETRACE();  // sub-tile rows written + fenced
                // the previous store (other buffer) must have drained before anyone refills it after the barrier
                if (store_leader && !V3D_DBG(8)) bulk_wait_read<0>();
                V3D_ETRACE();  // previous store drained
                asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
                V3D_ETRACE();  // group met
                if (store_leader && !V3D_DBG(2)) {
                  const uint32_t src = stg + static_cast<uint32_t>((nstore & 1) * SUB_BYTES);
                  const int col0 = obase + c - (CPS - 1) * 16;
                  if (CONV) tma_store_4d(&mapD, src, col0, t0, t1, t2);
                  else tma_store_3d(&mapD, src, col0, t1, t0);
                  if (!CONV && EPI == EPI_BF16 && kv.n > 0 && col0 >= kv.col0) {
                    // fused all-gather: the same sub-tile into every rank's gather buffer (peer memory over NVLink)
                    for (int i = 0; i < kv.n; ++i) tma_store_3d(&kv.m[i], src, col0 - kv.col0, t1, t0);
                  }
                  bulk_commit();
                }
                V3D_ETRACE();  // sub-tile handed to TMA
                // RTMA: every thread of the group read this sub-tile's residual slot before the barrier above, so
                // the slot can take the sub-tile RD further on
                if (RTMA && store_leader && (g + static_cast<int>(RD)) * LG < my_n)
                  request_res(obase + (c_begin + (g + static_cast<int>(RD)) * LG) * 16, t0, t1, t2, nstore);
                ++nstore;
              }
            }
          }
        }
      }
      if (RTMA && store_leader && tile + static_cast<int>(gridDim.x) / NCTA < total_tiles) {
        // every residual slot is free now (every thread of the group passed the last sub-tile's barrier): request
        // the head of the NEXT tile's share, so that it is in shared memory before that tile's accumulator is ready
        int u0, u1, u2;
        tile_origin<CONV>(p, tw.m_tile * NCTA + cta_rank, u0, u1, u2);
        const int hsel_next = (NSUB & 1) ? (half ^ ((iter + 1) & 1)) : half;
        const int my_n_next = hsel_next ? NCH - CH_HALF : CH_HALF;
        const int col_next = tw.n_tile * OUT_COLS + c_begin_next * 16;
#pragma unroll
        for (uint32_t j = 0; j < RD; ++j)
          if (static_cast<int>(j) * LG < my_n_next)
            request_res(col_next + static_cast<int>(j) * LG * 16, u0, u1, u2, nstore + j);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (NCTA == 2) mbar_arrive_leader(&tempty_bar[acc]);
        else mbar_arrive(&tempty_bar[acc]);
      }
      V3D_ETRACE();  // tile done
      if (++acc == C::ACC_STAGES) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
#undef V3D_ETRACE
    if (STAGED && lane == 0) bulk_wait_all();  // smem slabs must outlive the last TMA store reads
  }

  tc_fence_before();
  __syncthreads();
  if (NCTA == 2) cluster_sync_all();  // the peer's shared memory and TMEM stay alive until the leader's MMAs retired
  if (warp == 1) {
    tc_fence_after();
    if (NCTA == 2) tmem_dealloc_2cta(tmem_base, C::TMEM_COLS);
    else tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static long long* g_trace = nullptr;  // diagnostics: set through v3d_debug_set_trace
static int pick_block_n(int N, int act) {
  static const int cands[] = {256, 160, 128, 64, 32, 16};
  for (int c : cands) {
    if (N % c != 0) continue;
    if (act == V3D_ACT_GEGLU && (c / 2) % 16 != 0) continue;
    return c;
  }
  return 0;
}

static KvMaps g_kv_none;  // zero-initialised: n = 0
static const KvMaps* g_kv = &g_kv_none;  // set by v3d_gemm_bf16 around the one launch that scatters K|V

template <int BN, bool CONV, int EPI, int NCTA = 1>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& md, const GemmEpi& epi,
                  cudaStream_t st, const CUtensorMap* mr = nullptr) {
  using C = Cfg<BN, NCTA, EPI == EPI_BF16RT>;
  const CUtensorMap& mres = mr ? *mr : md;  // only the EPI_BF16RT instantiations read it
  static bool configured = false;
  auto kern = gemm_tc_kernel<BN, CONV, EPI, NCTA>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%d) failed: %s", C::SMEM_BYTES, cudaGetErrorString(e));
      return V3D_ERR_CUDA;
    }
    configured = true;
  }
  if (NCTA == 1) {
    const int total = epi.num_m_tiles * epi.num_n_tiles;
    const int grid = total < num_sms() ? total : num_sms();
    kern<<<grid, kThreads, C::SMEM_BYTES, st>>>(ma, mb, md, epi, mres, *g_kv);
  } else {
    // one cluster of two CTAs per 256-row pair tile, persistent over at most num_sms / 2 clusters
    const int total = ((epi.num_m_tiles + 1) / 2) * epi.num_n_tiles;
    const int clusters = total < num_sms() / 2 ? total : num_sms() / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(static_cast<unsigned>(2 * clusters));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ma, mb, md, epi, mres, *g_kv);
    if (e != cudaSuccess) {
      set_error("gemm_tc_kernel (CTA pair) launch failed: %s", cudaGetErrorString(e));
      return V3D_ERR_CUDA;
    }
  }
  V3D_CHECK_LAUNCH("gemm_tc_kernel");
  return V3D_OK;
}

// CTA-pair variants exist for the wide bf16-output tiles only (the shapes that are ingest- or shared-memory-bound
// with one CTA: 160-wide UNet layers, the 128-wide decoder levels)
template <int BN, bool CONV>
static int dispatch_epi_pair(int epi_kind, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& md,
                             const GemmEpi& epi, cudaStream_t st) {
  if constexpr (BN == 128 || BN == 160 || BN == 256) {
    switch (epi_kind) {
      case EPI_BF16: return launch<BN, CONV, EPI_BF16, 2>(ma, mb, md, epi, st);
      case EPI_BF16R2: return launch<BN, CONV, EPI_BF16R2, 2>(ma, mb, md, epi, st);
      case EPI_GEGLU:
        if constexpr (!CONV && BN == 256) return launch<256, false, EPI_GEGLU, 2>(ma, mb, md, epi, st);
        break;
    }
  }
  set_error("no CTA-pair variant for epilogue %d / block_n %d", epi_kind, BN);
  return V3D_ERR_BAD_ARG;
}

template <int BN, bool CONV>
static int dispatch_epi(int epi_kind, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& md,
                        const GemmEpi& epi, cudaStream_t st, const CUtensorMap* mr = nullptr) {
  switch (epi_kind) {
    case EPI_BF16: return launch<BN, CONV, EPI_BF16>(ma, mb, md, epi, st);
    case EPI_BF16RT:
      if constexpr (BN % 32 == 0 && BN >= 64) {
        if (mr != nullptr) return launch<BN, CONV, EPI_BF16RT>(ma, mb, md, epi, st, mr);
      }
      break;
    case EPI_BF16R2: return launch<BN, CONV, EPI_BF16R2>(ma, mb, md, epi, st);
    case EPI_F32: return launch<BN, CONV, EPI_F32>(ma, mb, md, epi, st);
    case EPI_GEGLU:
      if constexpr (!CONV && (BN / 2) % 16 == 0) return launch<BN, false, EPI_GEGLU>(ma, mb, md, epi, st);
      break;
    case EPI_TRANS:
      if constexpr (!CONV) return launch<BN, false, EPI_TRANS>(ma, mb, md, epi, st);
      break;
  }
  set_error("unsupported epilogue %d for this mode/tile", epi_kind);
  return V3D_ERR_BAD_ARG;
}

template <bool CONV>
static int dispatch_bn(int bn, int epi_kind, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& md,
                       const GemmEpi& epi, cudaStream_t st, bool pair = false, const CUtensorMap* mr = nullptr) {
  if (pair) {
    switch (bn) {
      case 256: return dispatch_epi_pair<256, CONV>(epi_kind, ma, mb, md, epi, st);
      case 160: return dispatch_epi_pair<160, CONV>(epi_kind, ma, mb, md, epi, st);
      case 128: return dispatch_epi_pair<128, CONV>(epi_kind, ma, mb, md, epi, st);
      default: set_error("no CTA-pair variant for block_n %d", bn); return V3D_ERR_BAD_ARG;
    }
  }
  switch (bn) {
    case 256: return dispatch_epi<256, CONV>(epi_kind, ma, mb, md, epi, st, mr);
    case 160: return dispatch_epi<160, CONV>(epi_kind, ma, mb, md, epi, st, mr);
    case 128: return dispatch_epi<128, CONV>(epi_kind, ma, mb, md, epi, st, mr);
    case 64: return dispatch_epi<64, CONV>(epi_kind, ma, mb, md, epi, st, mr);
    case 32: return dispatch_epi<32, CONV>(epi_kind, ma, mb, md, epi, st, mr);
    case 16: return dispatch_epi<16, CONV>(epi_kind, ma, mb, md, epi, st, mr);
    default: set_error("unsupported block_n %d", bn); return V3D_ERR_BAD_ARG;
  }
}

}  // namespace v3d

using namespace v3d;

/* diagnostics: device buffer of >= 3072 int64 that CTA 0 of every subsequent GEMM launch fills with clock64()
 * timelines (producer k-blocks | MMA issuer | epilogue warp 2); NULL disables. Not part of the product path. */
extern "C" int v3d_debug_set_trace(void* buf) {
#ifdef V3D_GEMM_DIAG
  g_trace = static_cast<long long*>(buf);
  return V3D_OK;
#else
  if (buf != nullptr) {
    set_error("v3d_debug_set_trace: library built without V3D_GEMM_DIAG");
    return V3D_ERR_UNSUPPORTED;
  }
  g_trace = nullptr;
  return V3D_OK;
#endif
}

extern "C" int v3d_gemm_pick_block_n(int32_t N, int32_t act) { return pick_block_n(N, act); }

extern "C" int v3d_geglu_pack_rows(int32_t n_out, int32_t block_n, int32_t* perm) {
  if (n_out <= 0 || block_n <= 0 || (block_n & 1) || (2 * n_out) % block_n != 0 || perm == nullptr) {
    set_error("v3d_geglu_pack_rows: bad n_out=%d block_n=%d", n_out, block_n);
    return V3D_ERR_BAD_ARG;
  }
  const int half = block_n / 2;
  const int tiles = 2 * n_out / block_n;
  for (int t = 0; t < tiles; ++t) {
    for (int j = 0; j < half; ++j) {
      perm[t * block_n + j] = t * half + j;                 // value rows
      perm[t * block_n + half + j] = n_out + t * half + j;  // gate rows
    }
  }
  return V3D_OK;
}

extern "C" int v3d_gemm_args_size(void) { return static_cast<int>(sizeof(v3d_gemm_args)); }

extern "C" int v3d_gemm_bf16(const v3d_gemm_args* a, void* stream) {
  if (a == nullptr || a->A == nullptr || a->B == nullptr || a->D == nullptr) {
    set_error("v3d_gemm_bf16: null pointer");
    return V3D_ERR_BAD_ARG;
  }
  const bool conv = a->conv_w > 0;
  const int ntaps = conv ? 9 : (a->ntaps > 0 ? a->ntaps : 1);
  if (a->K <= 0 || a->K % BK != 0 || a->N <= 0 || a->N % 16 != 0) {
    set_error("v3d_gemm_bf16: K=%d must be a multiple of 64 and N=%d of 16", a->K, a->N);
    return V3D_ERR_BAD_ARG;
  }
  if (a->out_transposed && (conv || ntaps != 1 || !a->out_fp32 || a->batch > 1 || a->act == V3D_ACT_GEGLU || a->R1 || a->R2 || a->fbias)) {
    set_error("v3d_gemm_bf16: out_transposed needs a plain fp32 linear GEMM (no taps/conv/batch/residual/geglu)");
    return V3D_ERR_BAD_ARG;
  }
  if (a->act == V3D_ACT_GEGLU && (a->R1 || a->R2 || a->fbias)) {
    set_error("v3d_gemm_bf16: GEGLU takes no residual or per-frame bias");
    return V3D_ERR_BAD_ARG;
  }
  if (!conv && ntaps != 1 && ntaps != 3) {
    set_error("v3d_gemm_bf16: ntaps must be 1 or 3");
    return V3D_ERR_BAD_ARG;
  }
  if ((a->lda % 8) || (a->ldb % 8) || (a->ldd % 8) || (a->R1 && a->ldr1 % 8) || (a->R2 && a->ldr2 % 8) ||
      (a->ldfb % 4)) {
    set_error("v3d_gemm_bf16: leading dimensions must be multiples of 8 elements");
    return V3D_ERR_BAD_ARG;
  }
  int bn = a->block_n > 0 ? a->block_n : pick_block_n(a->N, a->act);
  if (a->block_n <= 0 && a->act != V3D_ACT_GEGLU && bn >= 128) {
    // wave quantisation: among the wide tiles that divide N, take the one that fills the SMs best
    // (e.g. M = 2304, N = 1280: 90 tiles of 128x256 leave 58 SMs idle, 144 tiles of 128x160 do not)
    long long mt;
    if (a->conv_w > 0) {
      const int bw = a->conv_w < BM ? a->conv_w : BM;
      int bh = BM / bw;
      if (bh > a->conv_h) bh = a->conv_h;
      const int bnimg = BM / (bw * (bh > 0 ? bh : 1));
      mt = static_cast<long long>(a->conv_w / bw) * (a->conv_h / (bh > 0 ? bh : 1)) * ((a->conv_n + bnimg - 1) / bnimg);
    } else {
      mt = static_cast<long long>(a->batch > 0 ? a->batch : 1) * ((a->rows_per_batch + BM - 1) / BM);
    }
    static const int cands[] = {256, 160, 128};
    static const double mma_eff[] = {1.0, 0.97, 0.80};  // 128-wide tiles are shared-memory-bandwidth bound
    double best = -1.0;
    const int sms = num_sms();
    for (int i = 0; i < 3; ++i) {
      if (a->N % cands[i] != 0) continue;
      const long long tiles = mt * (a->N / cands[i]);
      const long long waves = (tiles + sms - 1) / sms;
      const double score = mma_eff[i] * static_cast<double>(tiles) / static_cast<double>(waves * sms);
      if (score > best + 1e-9) {
        best = score;
        bn = cands[i];
      }
    }
  }
  if (bn == 0 || a->N % bn != 0 || (a->act == V3D_ACT_GEGLU && (bn / 2) % 16 != 0)) {
    set_error("v3d_gemm_bf16: no valid N tile for N=%d act=%d block_n=%d", a->N, a->act, a->block_n);
    return V3D_ERR_BAD_ARG;
  }

  GemmEpi e;
  memset(&e, 0, sizeof(e));
  e.D = a->D;
  e.bias = a->bias;
  e.fbias = a->fbias;
  e.R1 = static_cast<const bf16*>(a->R1);
  e.R2 = static_cast<const bf16*>(a->R2);
  e.ldd = a->ldd;
  e.ldr1 = a->ldr1;
  e.ldr2 = a->ldr2;
  e.ldfb = a->ldfb > 0 ? a->ldfb : a->N;
  e.N = a->N;
  e.kpt = a->K / BK;
  e.ntaps = ntaps;
  e.tap_shift = a->tap_shift;
  e.tap0 = a->a_row0 - (ntaps >> 1) * a->tap_shift;
  e.rows_per_frame = a->rows_per_frame > 0 ? a->rows_per_frame : 1;
  e.act = a->act;
  e.out_fp32 = a->out_fp32;
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* v = getenv("V3D_GEMM_DEBUG");
      dbg = v ? atoi(v) : 0;
    }
    e.dbg = dbg;
    e.trace = g_trace;
  }
  if (a->fbias != nullptr) {
    if (conv) {
      const int bw_ = a->conv_w < BM ? a->conv_w : BM;
      int bh_ = BM / bw_;
      if (bh_ > a->conv_h) bh_ = a->conv_h;
      e.fb_uniform = (bw_ * bh_) % 32 == 0 && e.rows_per_frame % (a->conv_w * a->conv_h) == 0;
    } else {
      e.fb_uniform = e.rows_per_frame % 32 == 0 && (a->batch <= 1 || a->rows_per_batch % 32 == 0);
    }
  }
  e.transposed = a->out_transposed;
  e.valid_cols = a->valid_cols > 0 ? a->valid_cols : a->N;
  e.accumulate = a->accumulate;
  e.s0 = a->s0;
  e.s1 = a->s1;
  e.s2 = a->s2;
  e.num_n_tiles = a->N / bn;

  CUtensorMap ma, mb;
  int rc;
  int b_batch = 1;
  if (conv) {
    const int W = a->conv_w, H = a->conv_h, NI = a->conv_n;
    if (H <= 0 || NI <= 0) {
      set_error("v3d_gemm_bf16: bad conv geometry");
      return V3D_ERR_BAD_ARG;
    }
    const int bw = W < BM ? W : BM;
    if (BM % bw != 0 || W % bw != 0) {
      set_error("conv3x3: width %d does not tile into 128-pixel boxes", W);
      return V3D_ERR_UNSUPPORTED;
    }
    int bh = BM / bw;
    if (bh > H) bh = H;
    if (H % bh != 0 || (BM / bw) % bh != 0) {
      set_error("conv3x3: height %d does not tile into 128-pixel boxes (bw=%d)", H, bw);
      return V3D_ERR_UNSUPPORTED;
    }
    const int bnimg = BM / (bw * bh);
    if ((bw & (bw - 1)) != 0 || (bh & (bh - 1)) != 0) {
      set_error("conv3x3: tile box %dx%d is not a power of two", bw, bh);
      return V3D_ERR_UNSUPPORTED;
    }
    e.cn = NI; e.ch = H; e.cw = W; e.bw = bw; e.bh = bh; e.bn = bnimg;
    e.bw_shift = __builtin_ctz(bw);
    e.bh_shift = __builtin_ctz(bh);
    e.tiles_w = W / bw;
    e.tiles_h = H / bh;
    {
      const bool p2 = (e.tiles_w & (e.tiles_w - 1)) == 0 && (e.tiles_h & (e.tiles_h - 1)) == 0;
      e.tw_shift = p2 ? __builtin_ctz(e.tiles_w) : -1;
      e.th_shift = p2 ? __builtin_ctz(e.tiles_h) : -1;
    }
    e.num_m_tiles = e.tiles_w * e.tiles_h * ((NI + bnimg - 1) / bnimg);
    e.batch = 1; e.rows_per_batch = NI * H * W; e.tiles_per_batch = 1;
    const uint64_t dims[4] = {(uint64_t)a->K, (uint64_t)W, (uint64_t)H, (uint64_t)NI};
    const uint64_t str[3] = {(uint64_t)a->lda * 2, (uint64_t)a->lda * 2 * W, (uint64_t)a->lda * 2 * W * H};
    const uint32_t box[4] = {BK, (uint32_t)bw, (uint32_t)bh, (uint32_t)bnimg};
    rc = make_tmap_bf16(&ma, a->A, 4, dims, str, box);
    if (rc) return rc;
  } else {
    if (a->batch <= 0 || a->rows_per_batch <= 0) {
      set_error("v3d_gemm_bf16: bad batch/rows");
      return V3D_ERR_BAD_ARG;
    }
    e.batch = a->batch;
    e.rows_per_batch = a->rows_per_batch;
    e.tiles_per_batch = (a->rows_per_batch + BM - 1) / BM;
    e.num_m_tiles = e.batch * e.tiles_per_batch;
    const uint64_t abs_ = a->batch > 1 ? (uint64_t)a->a_batch_stride
                                       : (uint64_t)(a->a_rows > 0 ? a->a_rows : a->rows_per_batch) * a->lda;
    if (abs_ % 8 != 0) {
      set_error("v3d_gemm_bf16: a_batch_stride must be a multiple of 8");
      return V3D_ERR_BAD_ARG;
    }
    // halo'd operand (frame-sharded temporal convs): the tensor map spans a_rows >= rows_per_batch rows per batch
    // item and output row r reads A rows a_row0 + r (+ tap shifts); rows outside [0, a_rows) are zero-filled
    if (a->a_rows < 0 || a->a_row0 < 0 || (a->a_rows > 0 && a->a_rows < a->a_row0 + a->rows_per_batch)) {
      set_error("v3d_gemm_bf16: bad halo geometry a_rows=%d a_row0=%d rows_per_batch=%d", a->a_rows, a->a_row0,
                a->rows_per_batch);
      return V3D_ERR_BAD_ARG;
    }
    const uint64_t a_rows = a->a_rows > 0 ? (uint64_t)a->a_rows : (uint64_t)a->rows_per_batch + (uint64_t)a->a_row0;
    const uint64_t dims[3] = {(uint64_t)a->K, a_rows, (uint64_t)a->batch};
    const uint64_t str[2] = {(uint64_t)a->lda * 2, abs_ * 2};
    const uint32_t box[3] = {BK, BM, 1};
    rc = make_tmap_bf16(&ma, a->A, 3, dims, str, box);
    if (rc) return rc;
    if (a->b_batch_stride != 0) {
      b_batch = a->batch;
      e.b_batched = 1;
    }
  }
  // CTA-pair (cta_group::2) tiles, wide bf16-output tiles only.  V3D_GEMM_2CTA: unset / "auto" = per-shape choice
  // from the round-2 hardware timings (profiles/microbench_r2_pair_rtma.md: +5..8 % on the 256-wide GEGLU and decoder
  // conv tiles, +12 % on the decoder's 128-wide top level, +2..3 % on the residual-free projections; temporal convs,
  // residual epilogues and short grids lose), "1" = every eligible tile, "0" = never.
  bool pair = false;
  {
    static int want_pair = -2;
    if (want_pair == -2) {
      const char* v = getenv("V3D_GEMM_2CTA");
      want_pair = (v == nullptr || v[0] == 'a') ? -1 : (atoi(v) != 0 ? 1 : 0);
    }
    const bool staged_out = !a->out_transposed && !a->out_fp32;
    const bool eligible = staged_out && e.num_m_tiles >= 2 &&
                          (bn == 256 || ((bn == 160 || bn == 128) && a->act != V3D_ACT_GEGLU));
    if (want_pair == 1) {
      pair = eligible;
    } else if (want_pair == -1 && eligible && a->R1 == nullptr && a->R2 == nullptr) {
      if (a->act == V3D_ACT_GEGLU) pair = e.num_m_tiles >= 16;
      else if (conv) pair = e.num_m_tiles >= 64 && (bn != 160 || a->K >= 640);
      else if (ntaps == 1) pair = e.num_m_tiles >= 64 && b_batch == 1;
    }
  }
  {
    const uint64_t ktot = (uint64_t)ntaps * a->K;
    const uint64_t bbs = b_batch > 1 ? (uint64_t)a->b_batch_stride : (uint64_t)a->N * a->ldb;
    const uint64_t dims[3] = {ktot, (uint64_t)a->N, (uint64_t)b_batch};
    const uint64_t str[2] = {(uint64_t)a->ldb * 2, bbs * 2};
    const uint32_t box[3] = {BK, (uint32_t)(pair ? bn / 2 : bn), 1};  // a pair stages half of the B tile per CTA
    rc = make_tmap_bf16(&mb, a->B, 3, dims, str, box);
    if (rc) return rc;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int epi_kind = EPI_BF16;
  if (a->out_transposed) epi_kind = EPI_TRANS;
  else if (a->out_fp32) epi_kind = EPI_F32;
  else if (a->act == V3D_ACT_GEGLU) epi_kind = EPI_GEGLU;
  if ((a->act == V3D_ACT_GEGLU && epi_kind != EPI_GEGLU) || (a->act == V3D_ACT_SILU && epi_kind != EPI_F32)) {
    set_error("v3d_gemm_bf16: act=%d is not available with this output mode", a->act);
    return V3D_ERR_UNSUPPORTED;
  }
  CUtensorMap md;
  memset(&md, 0, sizeof(md));
  if (e.R2 != nullptr && e.R1 == nullptr) {  // a lone residual always travels as R1
    e.R1 = e.R2; e.ldr1 = e.ldr2; e.s1 = e.s2;
    e.R2 = nullptr;
  }
  if (epi_kind == EPI_BF16 && e.R2 != nullptr) epi_kind = EPI_BF16R2;
  // TMA-staged residual (EPI_BF16RT): single residual, 32-column sub-tiles, one CTA.  V3D_GEMM_RTMA: unset / "auto" =
  // the short-K linear projections (K <= 640: +30 % at K = 320, +10 % at K = 640 on hardware; longer K loses a pipeline
  // stage to the residual ring and gets slower), "1" = every eligible launch, "0" = never.
  bool rtma = false;
  {
    static int want_rtma = -2;
    if (want_rtma == -2) {
      const char* v = getenv("V3D_GEMM_RTMA");
      want_rtma = (v == nullptr || v[0] == 'a') ? -1 : (atoi(v) != 0 ? 1 : 0);
    }
    const bool eligible = epi_kind == EPI_BF16 && e.R1 != nullptr && !pair && bn % 32 == 0 && bn >= 64 &&
                          (reinterpret_cast<uintptr_t>(e.R1) & 15u) == 0;
    rtma = eligible && (want_rtma == 1 || (want_rtma == -1 && !conv && ntaps == 1 && a->K <= 640 && e.num_m_tiles >= 64));
  }
  if (rtma) epi_kind = EPI_BF16RT;
  CUtensorMap mr;
  memset(&mr, 0, sizeof(mr));
  if (epi_kind == EPI_BF16 || epi_kind == EPI_BF16R2 || epi_kind == EPI_GEGLU || epi_kind == EPI_BF16RT) {
    const uint64_t out_cols = static_cast<uint64_t>(e.num_n_tiles) * (epi_kind == EPI_GEGLU ? bn / 2 : bn);
    const uint32_t ocols_tile = epi_kind == EPI_GEGLU ? bn / 2 : bn;
    const uint32_t subw = ocols_tile % 32 == 0 ? 32 : 16;  // must match SUBW in the kernel
    const int swz = subw == 32 ? 64 : 0;
    if (conv) {
      const uint64_t dims[4] = {out_cols, (uint64_t)e.cw, (uint64_t)e.ch, (uint64_t)e.cn};
      const uint64_t str[3] = {(uint64_t)a->ldd * 2, (uint64_t)a->ldd * 2 * e.cw, (uint64_t)a->ldd * 2 * e.cw * e.ch};
      const uint32_t box[4] = {subw, (uint32_t)e.bw, (uint32_t)e.bh, (uint32_t)e.bn};
      rc = make_tmap_bf16(&md, a->D, 4, dims, str, box, swz);
    } else {
      const uint64_t dims[3] = {out_cols, (uint64_t)e.rows_per_batch, (uint64_t)e.batch};
      const uint64_t str[2] = {(uint64_t)a->ldd * 2, (uint64_t)a->ldd * 2 * e.rows_per_batch};
      const uint32_t box[3] = {subw, BM, 1};
      rc = make_tmap_bf16(&md, a->D, 3, dims, str, box, swz);
    }
    if (rc) return rc;
    if (rtma) {  // the residual through the same geometry as the output: [cols][rows...] with row stride ldr1
      if (conv) {
        const uint64_t dims[4] = {out_cols, (uint64_t)e.cw, (uint64_t)e.ch, (uint64_t)e.cn};
        const uint64_t str[3] = {(uint64_t)e.ldr1 * 2, (uint64_t)e.ldr1 * 2 * e.cw, (uint64_t)e.ldr1 * 2 * e.cw * e.ch};
        const uint32_t box[4] = {subw, (uint32_t)e.bw, (uint32_t)e.bh, (uint32_t)e.bn};
        rc = make_tmap_bf16(&mr, e.R1, 4, dims, str, box, swz);
      } else {
        const uint64_t dims[3] = {out_cols, (uint64_t)e.rows_per_batch, (uint64_t)e.batch};
        const uint64_t str[2] = {(uint64_t)e.ldr1 * 2, (uint64_t)e.ldr1 * 2 * e.rows_per_batch};
        const uint32_t box[3] = {subw, BM, 1};
        rc = make_tmap_bf16(&mr, e.R1, 3, dims, str, box, swz);
      }
      if (rc) return rc;
    }
  }
  // fused GEMM -> all-gather (K|V scatter into peer memory): plain linear bf16-output launches only
  KvMaps kvm;
  memset(&kvm, 0, sizeof(kvm));
  if (a->kv_n > 0) {
    if (conv || ntaps != 1 || epi_kind != EPI_BF16 || a->batch > 1 || a->kv_n > 8 || a->kv_col0 <= 0 ||
        a->kv_col0 % 32 != 0 || a->kv_col0 >= a->N || a->kv_ld < a->N - a->kv_col0 || a->kv_ld % 8 != 0 || bn % 32 != 0) {
      set_error("v3d_gemm_bf16: K|V scatter needs a plain linear bf16 GEMM without a second residual (kv_col0=%d kv_n=%d "
                "N=%d block_n=%d epilogue=%d)", a->kv_col0, a->kv_n, a->N, bn, epi_kind);
      return V3D_ERR_BAD_ARG;
    }
    kvm.n = a->kv_n;
    kvm.col0 = a->kv_col0;
    const uint64_t dims[3] = {(uint64_t)(a->N - a->kv_col0), (uint64_t)e.rows_per_batch, 1};
    const uint64_t str[2] = {(uint64_t)a->kv_ld * 2, (uint64_t)a->kv_ld * 2 * e.rows_per_batch};
    const uint32_t box[3] = {32, BM, 1};
    for (int i = 0; i < a->kv_n; ++i) {
      if (a->kv_dst[i] == nullptr) {
        set_error("v3d_gemm_bf16: kv_dst[%d] is null", i);
        return V3D_ERR_BAD_ARG;
      }
      rc = make_tmap_bf16(&kvm.m[i], a->kv_dst[i], 3, dims, str, box, 64);
      if (rc) return rc;
    }
    g_kv = &kvm;
  }
  rc = conv ? dispatch_bn<true>(bn, epi_kind, ma, mb, md, e, st, pair, rtma ? &mr : nullptr)
            : dispatch_bn<false>(bn, epi_kind, ma, mb, md, e, st, pair, rtma ? &mr : nullptr);
  g_kv = &g_kv_none;
  return rc;
}
