// Throughput of the exp2 forms a softmax can use, per SM: ex2.approx.ftz.f32 (1 result / lane-op),
// ex2.approx.f16x2 and ex2.approx.ftz.bf16x2 (2 results / lane-op), plus the conversions around them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/mufu tools/ubench/mufu.cu && tools/ubench/mufu
// Finding without running it (cuobjdump -sass): the packed forms compile to TWO MUFU ops on sm_100a
// (MUFU.EX2.F16 Rd, Ra.H1 and MUFU.EX2.F16 Rd', Ra), so they do not raise the exponential rate of a softmax.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[8];
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = -0.001f * (threadIdx.x + i);
    h[i] = 0xB800B400u + threadIdx.x + i;  // two small negative halves
  }
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      } else if (MODE == 1) {
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      } else if (MODE == 2) {
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      } else if (MODE == 3) {  // fp32 pair -> f16x2 -> exp2 (what a softmax would issue per 2 elements)
        uint32_t p;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(a[i]), "f"(a[(i + 1) & 7]));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(p));
        h[i] ^= p;
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_op) {
  float* out;
  long long* cyc;
  const int blocks = 148, threads = 1024, iters = 2000;
  cudaMalloc(&out, blocks * threads * sizeof(float));
  cudaMalloc(&cyc, blocks * sizeof(long long));
  k<MODE><<<blocks, threads>>>(out, cyc, 10);
  k<MODE><<<blocks, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < blocks; ++i) avg += h[i];
  avg /= blocks;
  const double ops = double(iters) * 8 * threads;  // lane-ops per SM (one CTA per SM)
  printf("%-34s %8.1f cycles  %6.2f lane-ops/clk/SM  %6.2f results/clk/SM  (%s)\n", name, avg, ops / avg,
         per_op * ops / avg, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("cvt.f16x2.f32 + ex2.f16x2", 2);
  return 0;
}
