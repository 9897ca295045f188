#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                  "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__global__ void __launch_bounds__(128) k(float* out, long long* cyc, int iters) {
  __shared__ uint32_t base_s;
  const int w = threadIdx.x >> 5;
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&base_s)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = base_s;
  const uint32_t ta = base + ((uint32_t)(w & 3) * 32u << 16);
  float v[8];
  for (int j = 0; j < 6; ++j) {
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 100.0f + j * 8 + i;
    tmem_st8(ta + 8 * j, v);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  float acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
    for (int j = 0; j < 6; ++j) {
      tmem_ld8(ta + 8 * j, v);
      for (int i = 0; i < 8; ++i) v[i] += 1.0f;
      acc += v[3];
      tmem_st8(ta + 8 * j, v);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  const long long t1 = clock64();
  tmem_ld8(ta + 8 * 5, v);
  out[blockIdx.x * 128 + threadIdx.x] = v[7] + acc * 0.0f;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "r"(64u));
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 5 * 128 * 4); cudaMalloc(&cyc, 8);
  const int iters = 1000;
  for (int ctas = 1; ctas <= 5; ctas += 4) {
    k<<<148 * ctas, 128>>>(out, cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    float h[256]; long long c;
    cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%d CTA/SM: thread 0 -> %.0f (expect %d), thread 77 -> %.0f (expect %d); %.1f cycles per ld+add+st+wait of 8 columns\n", ctas,
           h[0], 0 * 100 + 47 + iters, h[77], 77 * 100 + 47 + iters, (double)c / iters / 6);
  }
  return 0;
}
