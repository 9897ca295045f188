// Micro-benchmark (development tool): which block ids share an SM, and on which warp slots their warps sit, for a
// grid whose CTAs have the footprint of k_run_fast (128 threads, 128 registers, 43.6 KB shared memory -> 4 per SM).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(128, 4) k(int* smid, int* wslot, int spin) {
  extern __shared__ float sm[];
  unsigned id, wid;
  asm("mov.u32 %0, %%smid;" : "=r"(id));
  asm("mov.u32 %0, %%warpid;" : "=r"(wid));
  if (threadIdx.x == 0) smid[blockIdx.x] = (int)id;
  if ((threadIdx.x & 31) == 0) wslot[blockIdx.x * 4 + (threadIdx.x >> 5)] = (int)wid;
  float v = threadIdx.x;
  for (int i = 0; i < spin; ++i) { sm[threadIdx.x] = v; __syncthreads(); v += sm[(threadIdx.x + 1) & 127]; }
  if (v == 1.234f) smid[0] = -1;
}
int main() {
  const int n = 592 * 3;
  int *d, *w;
  cudaMalloc(&d, n * 4); cudaMalloc(&w, n * 16);
  const int smem = 56 * 1024;  // 4 CTAs per SM, like the run kernel (which is limited by registers)
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<<<n, 128, smem>>>(d, w, 20000);
  cudaDeviceSynchronize();
  static int h[n], hw[n * 4];
  cudaMemcpy(h, d, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hw, w, n * 16, cudaMemcpyDeviceToHost);
  for (int s = 0; s < 3; ++s) {
    const int smid = h[s * 50];
    printf("smid %3d:", smid);
    for (int i = 0; i < n; ++i) if (h[i] == smid) printf("  b%d[w %d %d %d %d]", i, hw[4 * i], hw[4 * i + 1], hw[4 * i + 2], hw[4 * i + 3]);
    printf("\n");
  }
  for (int mode = 0; mode < 3; ++mode) {  // rotation formulas: (id/148)&3, id&3, (id/64)&3
    int distinct = 0, same = 0;
    for (int s = 0; s < 148; ++s) {
      int cnt[4] = {0, 0, 0, 0}, c = 0;
      for (int i = 0; i < 592; ++i) if (h[i] == s) { const int r = mode == 0 ? (i / 148) & 3 : mode == 1 ? i & 3 : (i / 64) & 3; cnt[r]++; ++c; }
      int mx = 0; for (int r = 0; r < 4; ++r) mx = cnt[r] > mx ? cnt[r] : mx;
      distinct += (mx == 1 && c == 4); same += mx >= 3;
    }
    printf("first wave, formula %d: SMs with 4 distinct rotations %d, with >= 3 equal %d\n", mode, distinct, same);
  }
  return 0;
}
