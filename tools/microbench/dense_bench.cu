// Development micro-benchmark: the agents' dense pass / evaluation phase of dca_fast.cuh in isolation
// (cycles per pass per warp at 1..4 CTAs per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -DDCA_WG_REGS -I../../haskelldca_b200/csrc -o dense_bench dense_bench.cu
// (wGradCorr in registers: the tensor-memory variant needs the allocation the run kernel does)
#include <cstdio>
#include <cuda_runtime.h>
#include "dca_fast.cuh"

using namespace dca;
using namespace dca::fast;

template <int MODE>
__global__ void __launch_bounds__(NT, 4) kd(int iters, long long* cyc, float* out, int act0) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemF& sm = *reinterpret_cast<SmemF*>(smem_raw);
  for (int i = threadIdx.x; i < (int)(sizeof(SmemF) / 4); i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * WH; i += blockDim.x) (&sm.w[0][0])[i] = 1e-3f * (float)(i % 97);
  for (int i = threadIdx.x; i < WPAD; i += blockDim.x) { sm.x[i] = (uint8_t)(i % 5); sm.cnt2[i] = (uint8_t)(i % 3); }
  if (threadIdx.x == 0) { sm.alpha_net = 1e-6f; sm.alpha_avg = 0.01f; sm.alpha_grad = 1e-6f; }
  Ev& ev = sm.ev[0];
  if (threadIdx.x < 8) { ev.n4b[threadIdx.x] = make_uint2(0x00010100u, 0x00000101u); ev.n2b[threadIdx.x] = make_uint2(0x00000100u, 0x00000001u); }
  if (threadIdx.x < 72) ev.cand[threadIdx.x] = (uint8_t)(threadIdx.x * 9 % 70);
  if (threadIdx.x == 0) { ev.ncand = 7; ev.type = EV_NEW; ev.cell = 24; }
  __syncthreads();
  const Lane l = make_lane();
  if (l.vw == EVW) return;
  const int rc = l.rv ? l.r : 6;
  Agent ag;
#pragma unroll
  for (int j = 0; j < NSLOT; ++j) { ag.g[j].g01 = ag.g[j].g23 = ag.g[j].g45 = make_float2(1e-4f * j, 2e-4f); ag.g[j].g6 = 1e-4f; }
  const uint2 n4b_r = ev.n4b[rc], n2b_r = ev.n2b[rc];
  TdScalars td{1e-7f, 2e-7f, -1e-7f, 1e-8f};
  int act = act0;
  float acc = 0.f;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      dense_pass<true>(sm, ag, l, act, false, td, n4b_r, n2b_r);
      act = (act + 7) % 70;
      td.ctd += sm.P[5] * 1e-20f;  // serialise consecutive passes through the plane sums
    } else {
      const VTree vt = q_phase(sm, l, ev, 7, false, n4b_r, n2b_r);
      acc += vt.val + sm.q[0];
      sm.P[it & 63] = acc * 1e-20f;
    }
  }
  const long long t1 = clock64();
  float s = acc + td.ctd;
#pragma unroll
  for (int j = 0; j < NSLOT; ++j) s += ag.g[j].g01.x + ag.g[j].g6;
  out[blockIdx.x * NT + threadIdx.x] = s;
  if (threadIdx.x % 32 == 0 && blockIdx.x == 0) cyc[threadIdx.x / 32] = t1 - t0;
}

int main() {
  long long* cyc;
  float* out;
  cudaMalloc(&cyc, 64);
  cudaMalloc(&out, 4 << 20);
  Tables tab{};
  cudaMemcpyToSymbol(c_tab, &tab, sizeof tab);
  const int iters = 2000;
  cudaFuncSetAttribute(kd<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemF));
  cudaFuncSetAttribute(kd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemF));
  for (int mode = 0; mode < 2; ++mode)
    for (int ctas = 1; ctas <= 4; ++ctas) {
      if (mode == 0) kd<0><<<148 * ctas, NT, sizeof(SmemF)>>>(iters, cyc, out, 3);
      else kd<1><<<148 * ctas, NT, sizeof(SmemF)>>>(iters, cyc, out, 3);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long c[4];
      cudaMemcpy(c, cyc, sizeof c, cudaMemcpyDeviceToHost);
      printf("%s, %d CTA/SM: %.0f cycles per pass (hardware warps 0-2: %.0f %.0f %.0f)\n", mode == 0 ? "dense_pass<true>" : "q_phase(7 cand)",
             ctas, (double)c[0] / iters, (double)c[0] / iters, (double)c[1] / iters, (double)c[2] / iters);
    }
  return 0;
}
