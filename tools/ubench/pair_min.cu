// Bring-up probe for the CTA-pair (cta_group::2) GEMM tiles: the smallest program that uses every primitive the pair
// instantiations of gemm_tc_kernel rely on, ONE STAGE AT A TIME, with bounded waits - a protocol mistake shows up as
// "stage N timed out" instead of a hung device.  Uses the library's own PTX wrappers (csrc/common.cuh), so what is
// probed is what the product kernel executes.
//
//   stage 1  cluster of 2 launched; rank and shared-window address of each CTA (bit 24 must distinguish the ranks)
//   stage 2  tcgen05.alloc.cta_group::2 in both CTAs (same column base expected), relinquish
//   stage 3  mbarrier arrive from the PEER on the LEADER's barrier through the peer-bit-masked address
//   stage 4  TMA loads (.cta_group::2) of both CTAs crediting the LEADER's full barrier (A: own 128 rows, B: own half)
//   stage 5  four tcgen05.mma.cta_group::2 (M = 256, N = 128, K = 64) + multicast commit to BOTH CTAs' barriers
//   stage 6  each CTA reads its 128 accumulator rows from its own TMEM; host compares with A B^T computed on the CPU
//   stage 7  cluster barrier, tcgen05.dealloc.cta_group::2
//
// Build (from the repo root):
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 --expt-relaxed-constexpr -I include -I v3d_b200/csrc \
//        -o tools/ubench/pair_min tools/ubench/pair_min.cu v3d_b200/csrc/host_util.cu
// Run on a B200:  timeout 60 tools/ubench/pair_min
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_bf16.h>

#include "common.cuh"
#include "host_util.cuh"
#include "v3d_b200.h"

using namespace v3d;

constexpr int PM = 128, PN = 128, PK = 64;          // per-CTA A rows, full N, K
constexpr int A_BYTES = PM * PK * 2;                // 16 KB
constexpr int B_HALF_BYTES = (PN / 2) * PK * 2;     // 8 KB: this CTA's half of B
constexpr int NSTATUS = 16;

__device__ __forceinline__ bool bounded_wait(uint64_t* bar, uint32_t parity, int iters = 2000000) {
  for (int i = 0; i < iters; ++i) {
    if (mbar_try_wait(bar, parity)) return true;
    __nanosleep(64);
  }
  return false;
}

// status[cta][k]: 0 rank, 1 smem address of the barrier block, 2 tmem base, 3.. stage results (1 = ok, -1 = timeout)
__global__ void __launch_bounds__(128, 1)
pair_probe(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* out,
           int* status) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sa = smem;
  uint8_t* sb = smem + A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A_BYTES + B_HALF_BYTES);
  uint64_t* hello_bar = bars + 0;   // stage 3: leader's barrier, 2 arrivals (leader + peer)
  uint64_t* full_bar = bars + 1;    // stage 4: leader's barrier, tx bytes of both CTAs
  uint64_t* done_bar = bars + 2;    // stage 5: multicast commit lands in BOTH CTAs' copies
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  int* st = status + rank * NSTATUS;

  if (threadIdx.x == 0) {
    st[0] = rank;
    st[1] = static_cast<int>(smem_u32(bars));
    mbar_init(hello_bar, 2);
    mbar_init(full_bar, 1);
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc_2cta(tmem_slot, 128);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) st[2] = static_cast<int>(tmem_base);

  // ---- stage 3: remote arrive on the leader's barrier
  if (threadIdx.x == 0) {
    mbar_arrive_leader(hello_bar);                                   // both CTAs: masked address = the leader's copy
    if (rank == 0) st[3] = bounded_wait(hello_bar, 0) ? 1 : -1;
    else st[3] = 1;
  }
  __syncthreads();
  cluster_sync_all();

  // ---- stage 4: TMA of both CTAs credits the leader's full barrier
  if (threadIdx.x == 0) {
    if (rank == 0) mbar_arrive_expect_tx(full_bar, 2 * (A_BYTES + B_HALF_BYTES));
    tma_load_3d_2cta(sa, &mapA, full_bar, 0, rank * PM, 0);          // this CTA's 128 rows of A
    tma_load_3d_2cta(sb, &mapB, full_bar, 0, rank * (PN / 2), 0);    // this CTA's half of B's rows
    if (rank == 0) st[4] = bounded_wait(full_bar, 0) ? 1 : -1;
    else st[4] = 1;
  }
  __syncthreads();
  cluster_sync_all();   // peer's tile is in its shared memory (the leader waited for all bytes)

  // ---- stage 5: the leader issues the pair MMA; the commit is multicast to both CTAs
  bool mma_ok = true;
  if (rank == 0 && threadIdx.x == 0 && st[4] == 1) {
    tc_fence_after();
    constexpr uint32_t idesc = umma_idesc_bf16(2 * PM, PN, 0, 0);
    const uint64_t adesc = umma_desc_k_sw128(smem_u32(sa));
    const uint64_t bdesc = umma_desc_k_sw128(smem_u32(sb));
#pragma unroll
    for (int k = 0; k < PK / 16; ++k)
      tc_mma_f16_2cta(tmem_base, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc,
                      k != 0 ? 1u : 0u);
    tc_commit_2cta(done_bar, 3);
  }
  if (threadIdx.x == 0) {
    mma_ok = bounded_wait(done_bar, 0);
    st[5] = mma_ok ? 1 : -1;
  }
  __syncthreads();
  mma_ok = st[5] == 1;

  // ---- stage 6: every CTA reads its own 128 rows x 128 columns
  if (mma_ok) {
    tc_fence_after();
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const int row = rank * PM + warp * 32 + lane;
#pragma unroll
    for (int c = 0; c < PN / 16; ++c) {
      uint32_t v[16];
      tmem_ld16p(tmem_base + lane_base + static_cast<uint32_t>(c * 16), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) out[static_cast<size_t>(row) * PN + c * 16 + j] = __uint_as_float(v[j]);
    }
    if (threadIdx.x == 0) st[6] = 1;
  }

  // ---- stage 7
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 128);
  }
  if (threadIdx.x == 0) st[7] = 1;
}

int main() {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    printf("no CUDA device\n");
    return 2;
  }
  printf("device: %s (sm_%d%d)\n", prop.name, prop.major, prop.minor);
  std::vector<__nv_bfloat16> hA(2 * PM * PK), hB(PN * PK);
  srand(7);
  for (auto& x : hA) x = __float2bfloat16((rand() % 17 - 8) / 8.0f);
  for (auto& x : hB) x = __float2bfloat16((rand() % 13 - 6) / 4.0f);
  __nv_bfloat16 *dA, *dB;
  float* dOut;
  int* dStatus;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dOut, 2 * PM * PN * sizeof(float));
  cudaMalloc(&dStatus, 2 * NSTATUS * sizeof(int));
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dOut, 0xff, 2 * PM * PN * sizeof(float));
  cudaMemset(dStatus, 0, 2 * NSTATUS * sizeof(int));

  CUtensorMap ma, mb;
  {
    const uint64_t dims[3] = {PK, 2 * PM, 1};
    const uint64_t str[2] = {PK * 2, static_cast<uint64_t>(PK) * 2 * 2 * PM};
    const uint32_t box[3] = {PK, PM, 1};
    if (make_tmap_bf16(&ma, dA, 3, dims, str, box)) { printf("tensor map A failed\n"); return 2; }
  }
  {
    const uint64_t dims[3] = {PK, PN, 1};
    const uint64_t str[2] = {PK * 2, static_cast<uint64_t>(PK) * 2 * PN};
    const uint32_t box[3] = {PK, PN / 2, 1};
    if (make_tmap_bf16(&mb, dB, 3, dims, str, box)) { printf("tensor map B failed\n"); return 2; }
  }
  const int smem = A_BYTES + B_HALF_BYTES + 1024 + 256;
  cudaFuncSetAttribute(pair_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, pair_probe, ma, mb, dOut, dStatus);
  if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 2; }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 2; }

  int hs[2 * NSTATUS];
  cudaMemcpy(hs, dStatus, sizeof(hs), cudaMemcpyDeviceToHost);
  const char* names[8] = {"", "", "", "peer -> leader mbarrier arrive", "2-CTA TMA -> leader barrier",
                          "pair MMA + multicast commit", "TMEM read-back", "cluster sync + dealloc"};
  for (int r = 0; r < 2; ++r) {
    printf("CTA rank %d: barrier smem addr 0x%08x (bit 24 = %d), tmem base 0x%08x\n", hs[r * NSTATUS],
           hs[r * NSTATUS + 1], (hs[r * NSTATUS + 1] >> 24) & 1, hs[r * NSTATUS + 2]);
    for (int sidx = 3; sidx <= 7; ++sidx)
      printf("  stage %d %-32s %s\n", sidx, names[sidx],
             hs[r * NSTATUS + sidx] == 1 ? "ok" : (hs[r * NSTATUS + sidx] == -1 ? "TIMED OUT" : "not reached"));
  }
  std::vector<float> hOut(2 * PM * PN);
  cudaMemcpy(hOut.data(), dOut, hOut.size() * 4, cudaMemcpyDeviceToHost);
  double max_err = 0.0;
  int bad = 0;
  for (int m = 0; m < 2 * PM; ++m)
    for (int n = 0; n < PN; ++n) {
      float ref = 0.f;
      for (int k = 0; k < PK; ++k) ref += __bfloat162float(hA[m * PK + k]) * __bfloat162float(hB[n * PK + k]);
      const double d = fabs(static_cast<double>(hOut[m * PN + n]) - ref);
      if (!(d <= 1e-3)) ++bad;   // small integers over 8 and 4: products and sums are exact in fp32
      if (d > max_err) max_err = d;
    }
  printf("D = A B^T (256 x 128 x 64): max |err| %.3g, %d mismatches -> %s\n", max_err, bad, bad == 0 ? "PASS" : "FAIL");
  return bad == 0 ? 0 : 1;
}
