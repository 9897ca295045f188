// dca_ops.cuh -- the reference's two Accelerate operators on EXPLICIT arrays.
//
// accFn1 = runN bkend getAction, accFn2 = runN bkend gridStepTrain
// (src/SimRunner.hs:61-62, 211-212) are pure functions of their array arguments.
// These kernels implement exactly that contract (arrays in, arrays out, in the
// reference's own layouts: grid uint8[7][7][70], frep int64[7][7][71], weights
// float[3479] in flatten order), in both arithmetic orders, so that the Haskell
// `--backend gpu` can substitute them one for one.  They are the single-simulation
// CLI path; the batched throughput path is dca_fast.cuh.
#pragma once
#include "dca_kernels.cuh"

namespace dca {
namespace ops {

constexpr int NT = 128;

struct DotScratch {
  float R[NFEAT * 8];  // row sums [f][r], r = 7 is the zero pad
  float P[72];
  float W[4];
  float out;
  float out3[3];
  float prod[3][FREP + 1];  // REF order: the products of up to three dots, formed by all threads before the serial folds
};
// right fold p0 + (p1 + (... + (p3478 + 0))) of products already in shared memory (one thread)
__device__ __forceinline__ float ref_fold(const float* p) {
  float acc = 0.0f;
  static_assert(FREP % 7 == 0, "the fold loads seven products at a time");
#pragma unroll 2
  for (int i0 = FREP - 7; i0 >= 0; i0 -= 7) {
    float v[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) v[k] = p[i0 + k];
#pragma unroll
    for (int k = 6; k >= 0; --k) acc = __fadd_rn(v[k], acc);
  }
  return acc;
}

__device__ __forceinline__ int fidx(int r, int c, int f) { return (r * COLS + c) * NFEAT + f; }

// rowdot of the FAST order on flatten-order arrays
__device__ __forceinline__ float rowdot_flat(const long long* x, const float* w, int f, int r) {
  float e = __fmul_rn((float)x[fidx(r, 0, f)], w[fidx(r, 0, f)]);
  float o = __fmul_rn((float)x[fidx(r, 1, f)], w[fidx(r, 1, f)]);
  e = __fmaf_rn((float)x[fidx(r, 2, f)], w[fidx(r, 2, f)], e);
  o = __fmaf_rn((float)x[fidx(r, 3, f)], w[fidx(r, 3, f)], o);
  e = __fmaf_rn((float)x[fidx(r, 4, f)], w[fidx(r, 4, f)], e);
  o = __fmaf_rn((float)x[fidx(r, 5, f)], w[fidx(r, 5, f)], o);
  e = __fmaf_rn((float)x[fidx(r, 6, f)], w[fidx(r, 6, f)], e);
  return __fadd_rn(e, o);
}

// Dense dot x.w of one frep with one weight vector by a CTA of 128 threads.
//   order REF : right fold x0*w0 + (x1*w1 + (... + 0)) in flatten order (Accelerate Interpreter)
//   order FAST: the kernel's trees; g_tree selects the thread-mapped tree of x.wGradCorr
// Every thread returns the result.
__device__ __forceinline__ float cta_dense_dot(DotScratch& t, int order, bool g_tree, const long long* x,
                                               const float* w) {
  __syncthreads();
  if (order == ORDER_REF) {  // products by all threads, the (inherently sequential) additions by one
    for (int i = threadIdx.x; i < FREP; i += blockDim.x) t.prod[0][i] = __fmul_rn((float)x[i], w[i]);
    __syncthreads();
    if (threadIdx.x == 0) t.out = ref_fold(t.prod[0]);
    __syncthreads();
    return t.out;
  }
  for (int i = threadIdx.x; i < NFEAT * 8; i += blockDim.x) {
    const int f = i >> 3, r = i & 7;
    t.R[i] = r < 7 ? rowdot_flat(x, w, f, r) : 0.0f;
  }
  __syncthreads();
  const int lane = lane_id(), wid = warp_id();
  if (!g_tree) {
    for (int f = threadIdx.x; f < NFEAT; f += blockDim.x) {
      const float* a = &t.R[f * 8];
      t.P[f] = __fadd_rn(__fadd_rn(__fadd_rn(a[0], a[1]), __fadd_rn(a[2], a[3])),
                         __fadd_rn(__fadd_rn(a[4], a[5]), __fadd_rn(a[6], a[7])));
    }
    __syncthreads();
    if (wid == 0) {
      float b = __fadd_rn(__fadd_rn(t.P[lane], t.P[lane + 32]), lane < 6 ? t.P[lane + 64] : 0.0f);
      for (int s = 1; s < 32; s <<= 1) b = __fadd_rn(b, __shfl_xor_sync(0xffffffffu, b, s));
      if (lane == 0) t.out = __fadd_rn(b, t.P[NCH]);
    }
  } else {
    if (wid < 3) {  // the agent-thread mapping of the run kernel: group g = 4*wid + lane/8, row r = lane%8
      const int g = 4 * wid + (lane >> 3), r = lane & 7;
      float s = 0.0f;
      if (r < 7) {
        s = t.R[g * 8 + r];
        for (int j = 1; j < 6 && 12 * j + g <= NCH; ++j) s = __fadd_rn(s, t.R[(12 * j + g) * 8 + r]);
      }
      for (int k = 1; k < 32; k <<= 1) s = __fadd_rn(s, __shfl_xor_sync(0xffffffffu, s, k));
      if (lane == 0) t.W[wid] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) t.out = __fadd_rn(__fadd_rn(t.W[0], t.W[1]), t.W[2]);
  }
  __syncthreads();
  return t.out;
}

// q[k] = dense dot of afterstate frep k with wNet (mvMul, src/AccUtils.hs:224-228); one CTA per k
__global__ void __launch_bounds__(NT) k_dense_dots(int order, const long long* freps, int n, const float* w, float* q) {
  __shared__ DotScratch t;
  const int k = blockIdx.x;
  if (k >= n) return;
  const float v = cta_dense_dot(t, order, false, freps + (size_t)k * FREP, w);
  if (threadIdx.x == 0) q[k] = v;
}

// x . wGradCorr of one frep (vvMul, src/Agent.hs:66) in the order's tree
__global__ void __launch_bounds__(NT) k_dense_dot_g(int order, const long long* frep, const float* wg, float* out) {
  __shared__ DotScratch t;
  const float v = cta_dense_dot(t, order, true, frep, wg);
  if (threadIdx.x == 0) *out = v;
}

// argpmax (src/AccUtils.hs:203-208): right fold with `a > b ? a : b` -> the last maximum
__global__ void k_argpmax(const float* q, int n, int* idx) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int best = n - 1;
    for (int i = n - 2; i >= 0; --i)
      if (q[i] > q[best]) best = i;
    *idx = best;
  }
}

// ---- accFn1 as one stream of launches without a host round trip: the candidate count stays on the device ----
// incAfterStateFreps for the candidates k_gf_chs found (count read from *n_dev)
__global__ void k_afterstates_dev(const RepState* st, int cell, int is_end, const long long* chs, const int* n_dev,
                                  const long long* frep, long long* out) {
  const int n = *n_dev;
  const unsigned long long n4 = c_tab.n4excl[cell], n2 = c_tab.n2incl[cell];
  const int diff = is_end ? -1 : 1;
  const uint8_t want = is_end ? 1 : 0;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < (long long)n * FREP;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx / FREP), i = (int)(idx - (long long)k * FREP);
    const int c2 = i / NFEAT, f = i - c2 * NFEAT;
    const int ch = (int)chs[k];
    long long v = frep[i];  // Gridfuncs.hs:206
    if (f == ch && ((n4 >> c2) & 1ULL)) v += diff;                                                       // :212-217
    if (f == NCH && ((n2 >> c2) & 1ULL) && st->cnt2[ch * PLANE + pos_of(c2)] == want) v -= diff;      // :220-252
    out[idx] = v;
  }
}
// mvMul (src/AccUtils.hs:224-228): one CTA per candidate, NCH CTAs launched
__global__ void __launch_bounds__(NT) k_dense_dots_dev(int order, const long long* freps, const int* n_dev, const float* w, float* q) {
  __shared__ DotScratch t;
  const int k = blockIdx.x;
  if (k >= *n_dev) return;
  const float v = cta_dense_dot(t, order, false, freps + (size_t)k * FREP, w);
  if (threadIdx.x == 0) q[k] = v;
}
// argpmax + the result of getAction (src/Simulator.hs:165-172, 202-205): (Just chs[idx], afreps[idx]) or (Nothing, frep)
__global__ void k_select_dev(const float* q, const int* n_dev, const long long* chs, const long long* afreps,
                             const long long* frep, int* ch_out, long long* frep_out) {
  __shared__ int best_s;
  const int n = *n_dev;
  if (threadIdx.x == 0) {
    int best = n - 1;
    for (int i = n - 2; i >= 0; --i)
      if (q[i] > q[best]) best = i;
    best_s = best;
    *ch_out = n > 0 ? (int)chs[best] : -1;
  }
  __syncthreads();
  const long long* src = n > 0 ? afreps + (size_t)best_s * FREP : frep;
  for (int i = threadIdx.x; i < FREP; i += blockDim.x) frep_out[i] = src[i];
}

struct BackwardOut {
  float loss;
  int loss_is_nan;
  float avg_reward;
  int pad;
};

// backward (src/Agent.hs:44-81) on explicit arrays; w, wg updated in place
__device__ __forceinline__ void backward_body(DotScratch& t, int order, float alpha_net, float alpha_avg, float alpha_grad,
                                              const long long* frep, const long long* next_frep, float reward,
                                              float avg_reward, float* w, float* wg, BackwardOut* out) {
  float val, next_val, dot;
  if (order == ORDER_REF) {  // the three right folds run side by side on three threads
    __syncthreads();
    for (int i = threadIdx.x; i < FREP; i += blockDim.x) {
      const float x = (float)frep[i], xn = (float)next_frep[i];
      t.prod[0][i] = __fmul_rn(x, w[i]);
      t.prod[1][i] = __fmul_rn(xn, w[i]);
      t.prod[2][i] = __fmul_rn(x, wg[i]);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && threadIdx.x < 96) t.out3[threadIdx.x >> 5] = ref_fold(t.prod[threadIdx.x >> 5]);
    __syncthreads();
    val = t.out3[0]; next_val = t.out3[1]; dot = t.out3[2];  // :62, :63, :66
    __syncthreads();
  } else {
    val = cta_dense_dot(t, order, false, frep, w);           // :62
    next_val = cta_dense_dot(t, order, false, next_frep, w); // :63
    dot = cta_dense_dot(t, order, true, frep, wg);           // :66
  }
  const float td = __fsub_rn(__fadd_rn(__fsub_rn(reward, avg_reward), next_val), val);  // :65
  const float c = __fmul_rn(-2.0f, alpha_net);                                          // :67
  const float ctd = __fmul_rn(c, td), cavg = __fmul_rn(c, avg_reward), cdot = __fmul_rn(c, dot);
  const float sg = __fmul_rn(alpha_grad, __fsub_rn(td, dot));                           // :75
  for (int i = threadIdx.x; i < FREP; i += blockDim.x) {
    const float x = (float)frep[i], xn = (float)next_frep[i];
    float g;
    if (order == ORDER_FAST) {
      g = __fmaf_rn(-cdot, xn, __fmaf_rn(ctd, x, cavg));
      wg[i] = __fmaf_rn(sg, x, wg[i]);
    } else {
      g = __fsub_rn(__fadd_rn(__fmul_rn(ctd, x), cavg), __fmul_rn(cdot, xn));  // :68-71
      wg[i] = __fadd_rn(wg[i], __fmul_rn(sg, x));                            // :75-76
    }
    w[i] = __fsub_rn(w[i], g);  // :73
  }
  if (threadIdx.x == 0) {
    out->loss = td;
    out->loss_is_nan = (td != td) ? 1 : 0;                          // :80
    out->avg_reward = __fadd_rn(avg_reward, __fmul_rn(alpha_avg, td));  // :78
    out->pad = 0;
  }
}
__global__ void __launch_bounds__(NT) k_backward(int order, float alpha_net, float alpha_avg, float alpha_grad,
                                                 const long long* frep, const long long* next_frep, float reward,
                                                 float avg_reward, float* w, float* wg, BackwardOut* out) {
  __shared__ DotScratch t;
  backward_body(t, order, alpha_net, alpha_avg, alpha_grad, frep, next_frep, reward, avg_reward, w, wg, out);
}
// gridStepTrain's second half with the reward still on the device (reward' = fromIntegral reward, SimRunner.hs:117)
__global__ void __launch_bounds__(NT) k_backward_dev(int order, float alpha_net, float alpha_avg, float alpha_grad,
                                                     const long long* frep, const long long* next_frep,
                                                     const long long* reward_dev, float avg_reward, float* w, float* wg,
                                                     BackwardOut* out) {
  __shared__ DotScratch t;
  backward_body(t, order, alpha_net, alpha_avg, alpha_grad, frep, next_frep, (float)*reward_dev, avg_reward, w, wg, out);
}

// executeAction (src/Gridfuncs.hs:105-110) on a byte grid + boolSum (src/AccUtils.hs:166-167)
__global__ void k_grid_step(uint8_t* grid, int cell, int is_end, int ch, long long* reward) {
  __shared__ int cnt;
  if (threadIdx.x == 0) {
    cnt = 0;
    if (ch >= 0) grid[cell * NCH + ch] = is_end ? 0 : 1;
  }
  __syncthreads();
  int c = 0;
  for (int i = threadIdx.x; i < NCELL * NCH; i += blockDim.x) c += grid[i] ? 1 : 0;
  for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
  if (lane_id() == 0) atomicAdd(&cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) *reward = cnt;
}

}  // namespace ops
}  // namespace dca
