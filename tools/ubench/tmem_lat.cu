// Micro-benchmark: latency of tcgen05.ld / tcgen05.wait::ld and of the epilogue's fences, on an idle SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_lat tmem_lat.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

#define LD_X(N, REGS) \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x" #N ".b32 {" REGS "}, [%" #N "];"

template <int N> struct Ld;
template <> struct Ld<16> {
  static __device__ __forceinline__ void go(uint32_t a, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(a));
  }
};
template <> struct Ld<32> {
  static __device__ __forceinline__ void go(uint32_t a, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(a));
  }
};
template <> struct Ld<64> {
  static __device__ __forceinline__ void go(uint32_t a, uint32_t* v) {
    Ld<32>::go(a, v);
    Ld<32>::go(a + 32, v + 32);
  }
};

__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// res[warp][test] = cycles (lane 0)
template <int N>
__device__ __forceinline__ long long time_ld(uint32_t taddr, int reps, uint32_t& sink) {
  uint32_t v[N];
  long long t0 = clock64();
  for (int i = 0; i < reps; ++i) {
    Ld<N>::go(taddr, v);
    wait_ld();
#pragma unroll
    for (int j = 0; j < N; ++j) sink ^= v[j];
  }
  return (clock64() - t0) / reps;
}

__global__ void __launch_bounds__(320, 1) k(long long* res, int active_warps) {
  __shared__ uint32_t slot;
  __shared__ __align__(1024) uint8_t buf[8192];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot;
  uint32_t sink = 0;
  if (warp >= 2 && warp < 2 + active_warps) {
    const uint32_t taddr = base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    long long r[10];
    // 0: empty wait
    {
      long long t0 = clock64();
      for (int i = 0; i < 16; ++i) wait_ld();
      r[0] = (clock64() - t0) / 16;
    }
    r[1] = time_ld<16>(taddr, 16, sink);
    r[2] = time_ld<32>(taddr, 16, sink);
    r[3] = time_ld<64>(taddr, 16, sink);
    // 4: fence.proxy.async after one st.shared
    {
      long long t0 = clock64();
      for (int i = 0; i < 16; ++i) {
        asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(smem_u32(buf) + lane * 16 + (warp - 2) * 512), "r"(i));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      r[4] = (clock64() - t0) / 16;
    }
    // 5: st.shared alone
    {
      long long t0 = clock64();
      for (int i = 0; i < 16; ++i)
        asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(smem_u32(buf) + lane * 16 + (warp - 2) * 512), "r"(i));
      r[5] = (clock64() - t0) / 16;
    }
    // 6: named barrier among the 4 warps of a group (only if all 8 active)
    r[6] = 0;
    if (active_warps == 8) {
      long long t0 = clock64();
      for (int i = 0; i < 16; ++i) asm volatile("bar.sync %0, 128;" ::"r"(1 + ((warp - 2) >> 2)) : "memory");
      r[6] = (clock64() - t0) / 16;
    }
    // 7: two x32 loads in flight then one wait (pipelined pair)
    {
      uint32_t v[64];
      long long t0 = clock64();
      for (int i = 0; i < 16; ++i) {
        Ld<32>::go(taddr, v);
        Ld<32>::go(taddr + 32, v + 32);
        wait_ld();
#pragma unroll
        for (int j = 0; j < 64; ++j) sink ^= v[j];
      }
      r[7] = (clock64() - t0) / 16;
    }
    // 8: clock64 pair overhead
    {
      long long t0 = clock64();
      long long t1 = clock64();
      r[8] = t1 - t0;
    }
    // 9: __syncwarp + lane0 branch
    {
      long long t0 = clock64();
      for (int i = 0; i < 16; ++i) __syncwarp();
      r[9] = (clock64() - t0) / 16;
    }
    if (lane == 0)
      for (int i = 0; i < 10; ++i) res[(warp - 2) * 10 + i] = r[i];
  }
  if (sink == 0x12345678u) res[99] = sink;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512));
}

int main() {
  long long* d;
  cudaMalloc(&d, 100 * sizeof(long long));
  const char* names[10] = {"wait::ld (nothing outstanding)", "ld.x16 + wait", "ld.x32 + wait", "2 x ld.x32 (64 cols) serial issue + wait",
                           "st.shared.v4 + fence.proxy.async", "st.shared.v4", "bar.sync 128 (4 warps)", "2 x ld.x32 + 1 wait",
                           "clock64 pair", "__syncwarp"};
  for (int aw : {1, 8}) {
    cudaMemset(d, 0, 100 * sizeof(long long));
    k<<<1, 320>>>(d, aw);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[100];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("== %d epilogue warp(s) active (idle SM, no MMA)\n", aw);
    for (int i = 0; i < 10; ++i) {
      printf("%-44s", names[i]);
      for (int w = 0; w < aw; ++w) printf(" %5lld", h[w * 10 + i]);
      printf("\n");
    }
  }
  return 0;
}
