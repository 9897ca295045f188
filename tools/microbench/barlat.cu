// Micro-benchmark (development tool): cost of one CTA barrier of 128 threads between a shared-memory store and
// a dependent load -- the pattern at both barriers of the FAST kernel's event loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o barlat barlat.cu && ./barlat
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  __shared__ float sm[256];
  sm[threadIdx.x] = threadIdx.x;
  sm[threadIdx.x + 128] = 0;
  __syncthreads();
  float v = threadIdx.x;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (MODE >= 1) sm[(threadIdx.x + it) & 127] = v;                               // store
    if (MODE >= 2) asm volatile("bar.sync 0, 128;" ::: "memory");                  // barrier
    v += sm[(threadIdx.x * 7 + it + (int)v) & 127];                                // dependent load
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = v;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 1 << 22);
  cudaMalloc(&cyc, 64);
  const int iters = 4000;
  const char* names[3] = {"dependent LDS only", "STS + dependent LDS", "STS + bar.sync + dependent LDS"};
  for (int ctas = 1; ctas <= 4; ctas += 3)
    for (int mode = 0; mode < 3; ++mode) {
      if (mode == 0) k<0><<<148 * ctas, 128>>>(out, cyc, iters);
      if (mode == 1) k<1><<<148 * ctas, 128>>>(out, cyc, iters);
      if (mode == 2) k<2><<<148 * ctas, 128>>>(out, cyc, iters);
      cudaDeviceSynchronize();
      long long c;
      cudaMemcpy(&c, cyc, sizeof c, cudaMemcpyDeviceToHost);
      printf("%d CTA/SM, %-32s: %.1f cycles per iteration\n", ctas, names[mode], (double)c / iters);
    }
  return 0;
}
