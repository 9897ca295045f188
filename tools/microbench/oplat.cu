// Micro-benchmark (development tool): dependent-issue latency and per-SMSP throughput of the instructions the
// FAST kernel is made of, on the GPU it runs on.  One CTA; W warps; every warp runs K independent chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o oplat oplat.cu && ./oplat
#include <cstdio>
#include <cuda_runtime.h>

enum { OP_FFMA, OP_FFMA2, OP_FADD2, OP_PRMT, OP_SHFL, OP_REDUX, OP_LDS, OP_FADD, OP_MIX };

template <int OP, int K>
__global__ void k(float* out, long long* cyc, int iters, float a, float b) {
  __shared__ float sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (float)(i & 31);
  __syncthreads();
  float2 v[K];
  unsigned u[K];
#pragma unroll
  for (int j = 0; j < K; ++j) { v[j] = make_float2(threadIdx.x + j, 1.0f + j); u[j] = threadIdx.x * 7 + j; }
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int j = 0; j < K; ++j) {
        if (OP == OP_FFMA) v[j].x = __fmaf_rn(v[j].x, a, b);
        if (OP == OP_FADD) v[j].x = __fadd_rn(v[j].x, a);
        if (OP == OP_FFMA2) v[j] = __ffma2_rn(v[j], a2, b2);
        if (OP == OP_FADD2) v[j] = __fadd2_rn(v[j], a2);
        if (OP == OP_PRMT) u[j] = __byte_perm(u[j], 0x4B000000u, 0x7440 + (r & 1));
        if (OP == OP_SHFL) v[j].x = __shfl_xor_sync(0xffffffffu, v[j].x, 1 + (r & 3));
        if (OP == OP_REDUX) u[j] = __reduce_max_sync(0xffffffffu, u[j] + threadIdx.x);
        if (OP == OP_LDS) u[j] = __float_as_uint(sm[(u[j] & 1023)]);
        if (OP == OP_MIX) {  // two scalar FFMAs instead of one packed
          v[j].x = __fmaf_rn(v[j].x, a, b);
          v[j].y = __fmaf_rn(v[j].y, a, b);
        }
      }
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < K; ++j) s += v[j].x + v[j].y + __uint_as_float(u[j]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP, int K>
void run(const char* name, int warps, float* out, long long* cyc) {
  const int iters = 2000;
  k<OP, K><<<1, 32 * warps>>>(out, cyc, iters, 1.0001f, 0.5f);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, sizeof c, cudaMemcpyDeviceToHost);
  const double n = (double)iters * 8 * K;  // instructions per warp
  printf("%-6s K=%d warps=%2d: %6.2f cycles per instruction per warp; %6.2f cycles per instruction per SMSP\n", name, K, warps,
         c / n, c / (n * ((warps + 3) / 4)));
}

#define ALLW(OP, NAME)                  \
  run<OP, 1>(NAME, 1, out, cyc);        \
  run<OP, 8>(NAME, 1, out, cyc);        \
  run<OP, 8>(NAME, 4, out, cyc);        \
  run<OP, 8>(NAME, 16, out, cyc);

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 1 << 20);
  cudaMalloc(&cyc, 64);
  ALLW(OP_FFMA, "FFMA")
  ALLW(OP_FFMA2, "FFMA2")
  ALLW(OP_MIX, "2xFFMA")
  ALLW(OP_FADD, "FADD")
  ALLW(OP_FADD2, "FADD2")
  ALLW(OP_PRMT, "PRMT")
  ALLW(OP_SHFL, "SHFL")
  ALLW(OP_REDUX, "REDUX")
  ALLW(OP_LDS, "LDS")
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
