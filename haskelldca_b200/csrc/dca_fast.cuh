// dca_fast.cuh -- the throughput kernel of libdca_b200 (FAST arithmetic order).
//
// One CTA of 128 threads per replica, resident for the whole launch:
//   warps 0-2  "agent" threads (96): the value function and the TDC update.
//              Lane r (0..6; lane 7 idles) of 8-lane group g (0..11) owns row r of
//              feature plane f = 12*slot + g for slot = 0..5.  wGradCorr lives in
//              those threads' registers (6 x 7 floats), wNet in shared memory (the
//              afterstate evaluation needs it with a flexible lane mapping).
//   warp 3     "event" warp: event queue, random numbers, statistics, candidate
//              channels, stop rules (environmentStep, src/Simulator.hs:101-134).
//
// Per event (runStep, src/SimRunner.hs:60-97) -- three CTA barriers:
//   [Q]   V(s) = x.w from the plane sums (butterfly), then every candidate channel
//         on its own 8-lane group: the two planes the action touches are
//         re-evaluated and substituted into the V tree (selectAction,
//         src/Simulator.hs:185-205 with incAfterStateFreps, src/Gridfuncs.hs:186-252)
//   B2    argmax (last maximum wins, src/AccUtils.hs:203-208), TD scalars
//         (src/Agent.hs:62-78) replicated in every thread; warp 0 executes the
//         action on grid / frep / cnt2 (src/Gridfuncs.hs:105-110); the event warp
//         books the statistics, pushes the new events and pops the next one
//   B3    agent warps: dense w / wGradCorr update fused with the next event's row
//         sums (src/Agent.hs:67-76);  event warp, concurrently: candidate channels
//         of the next event, queue pre-scan, random draws, stop rules
//   B1
//
// The fp32 arithmetic is the FAST order specified in DESIGN.md ("Arithmetic
// orders"; the CPU checker restates it independently): every operation is an explicit
// round-to-nearest intrinsic, packed two lanes at a time (FFMA2 / FADD2) where
// the two lanes are independent accumulators of the specification.
#pragma once
#include "dca_kernels.cuh"

namespace dca {
namespace fast {

constexpr int NT = 128;       // threads per replica
constexpr int NAGENT = 96;    // agent threads (warps 0..2)
constexpr int EVW = 3;        // the event warp
constexpr int NSLOT = 6;      // feature planes per agent lane group
constexpr int SLOTP = 12;     // planes per slot = agent lane groups
constexpr int WH = NROW * 4;  // floats per half of the weight array

struct Ev {  // the event being processed and its candidate set
  double time;
  int type, cell, ch, hoff;
  int ncand, pad;
};

struct SmemF {
  float w[2][WH];           // wNet: columns 0-3 of row rho = f*7+r at [0][rho*4+c], columns 4-6 at [1][rho*4+c-4]
  uint8_t x[WPAD];          // frep, [f][r][8]
  uint8_t cnt2[WPAD];       // users of ch within distance 2, [ch][r][8]
  uint32_t grid[NCELL][4];
  double q_time[QCAP];
  uint32_t q_id[QCAP];
  uint16_t q_key[QCAP];
  uint8_t q_hoff[QCAP];
  long long stats[8];
  float P[72];              // plane sums of x.w (current state, current weights)
  float q[72];              // afterstate values of the candidates
  float Wg[4];              // per-warp partial sums of x.wGradCorr
  uint2 n4b[8], n2b[8];     // row byte masks of the event cell: N4 \ {cell}, N2 u {cell}
  uint2 chgb[8];            // cells whose n_elig changes under the chosen action
  unsigned long long chgmask, n4mask, n2mask;
  Ev ev;
  int status, pad0;
  uint8_t cand[72];
};
static_assert(sizeof(SmemF) % 16 == 0, "SmemF size");
static_assert(offsetof(SmemF, x) % 8 == 0 && offsetof(SmemF, cnt2) % 8 == 0 && offsetof(SmemF, grid) % 16 == 0 &&
              offsetof(SmemF, q_time) % 16 == 0 && offsetof(SmemF, q_id) % 16 == 0 && offsetof(SmemF, q_key) % 16 == 0 &&
              offsetof(SmemF, q_hoff) % 8 == 0 && offsetof(SmemF, n4b) % 8 == 0, "SmemF alignment");

// ---------------------------------------------------------------------------
// packed fp32 helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
// bytes k, k+1 of v as floats: 0x4B0000bb is 2^23 + bb exactly
template <int K>
__device__ __forceinline__ float2 cvt_pair(uint32_t v) {
  float2 m = make_float2(__uint_as_float(__byte_perm(v, 0x4B000000u, 0x7440 + K)),
                         __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7440 + K + 1)));
  return __fadd2_rn(m, f2(-8388608.0f));
}
template <int K>
__device__ __forceinline__ float cvt_one(uint32_t v) {
  return __fadd_rn(__uint_as_float(__byte_perm(v, 0x4B000000u, 0x7440 + K)), -8388608.0f);
}
struct Row7 {  // one row of the feature representation as floats
  float2 a, b, c;
  float d;
};
__device__ __forceinline__ Row7 cvt_row(uint2 xb) {
  Row7 x;
  x.a = cvt_pair<0>(xb.x);
  x.b = cvt_pair<2>(xb.x);
  x.c = cvt_pair<0>(xb.y);
  x.d = cvt_one<2>(xb.y);
  return x;
}
// rowdot of the FAST order: even / odd column chains, then e + o
__device__ __forceinline__ float rowdot2(const Row7& x, float2 w01, float2 w23, float2 w45, float w6) {
  float2 acc = __fmul2_rn(x.a, w01);
  acc = __ffma2_rn(x.b, w23, acc);
  acc = __ffma2_rn(x.c, w45, acc);
  const float e = __fmaf_rn(x.d, w6, acc.x);
  return __fadd_rn(e, acc.y);
}
__device__ __forceinline__ uint2 badd(uint2 a, uint2 b) { return make_uint2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ uint2 bsub(uint2 a, uint2 b) { return make_uint2(a.x - b.x, a.y - b.y); }

// ---------------------------------------------------------------------------
// Agent side
// ---------------------------------------------------------------------------
struct Agent {       // per-thread registers of an agent thread
  float2 g01[NSLOT], g23[NSLOT], g45[NSLOT];  // wGradCorr rows of this lane, one per slot
  float g6[NSLOT];
};

struct TdScalars {
  float ctd, cavg, ncdot, sg;
};

// The dense pass over this lane's rows.  UPDATE: w' = w - g, wg' = fma(sg, x, wg)
// (Agent.hs:67-76) with x = the pre-action features, rebuilt for the two planes
// the action touched; then the row sums of x'.w' -> plane sums P, and the lane's
// partial of x'.wg' -> warp sums Wg.  Without UPDATE only the sums are formed.
template <bool UPDATE>
__device__ __forceinline__ void dense_pass(SmemF& sm, Agent& ag, const int g8, const int r, const int act,
                                           const bool is_end, const TdScalars td, const uint2 n4b_r) {
  const int rc = r < 7 ? r : 6;
  const bool rv = r < 7;
  float rg_acc = 0.0f;
  uint2 chgb_r = make_uint2(0, 0);
  if (UPDATE) chgb_r = sm.chgb[rc];
#pragma unroll
  for (int j = 0; j < NSLOT; ++j) {
    int f = SLOTP * j + g8;
    const bool pv = (j < NSLOT - 1) || (f <= NCH);
    if (j == NSLOT - 1 && f > NCH) f = NCH;
    const bool valid = pv && rv;
    const int row = f * 7 + rc;
    const uint2 xb = *reinterpret_cast<const uint2*>(&sm.x[f * PLANE + rc * ROWPAD]);
    float4 wa = *reinterpret_cast<const float4*>(&sm.w[0][row * 4]);
    float4 wb = *reinterpret_cast<const float4*>(&sm.w[1][row * 4]);
    float2 w01 = make_float2(wa.x, wa.y), w23 = make_float2(wa.z, wa.w), w45 = make_float2(wb.x, wb.y);
    float w6 = wb.z;
    const Row7 xn = cvt_row(xb);
    if (UPDATE) {
      const bool special = act >= 0 && (f == act || f == NCH);
      Row7 xo = xn;
      if (special) {  // the pre-action features of this row
        uint2 ob;
        if (f == act) ob = is_end ? badd(xb, n4b_r) : bsub(xb, n4b_r);
        else ob = is_end ? bsub(xb, chgb_r) : badd(xb, chgb_r);
        xo = cvt_row(ob);
      }
#define DCA_UPD2(W, G, XN, XO)                                            \
  {                                                                       \
    const float2 t_ = __ffma2_rn(f2(td.ctd), XO, f2(td.cavg));            \
    const float2 g_ = __ffma2_rn(f2(td.ncdot), XN, t_);                   \
    W = __ffma2_rn(g_, f2(-1.0f), W);                                     \
    G = __ffma2_rn(f2(td.sg), XO, G);                                     \
  }
      DCA_UPD2(w01, ag.g01[j], xn.a, xo.a)
      DCA_UPD2(w23, ag.g23[j], xn.b, xo.b)
      DCA_UPD2(w45, ag.g45[j], xn.c, xo.c)
#undef DCA_UPD2
      {
        const float t_ = __fmaf_rn(td.ctd, xo.d, td.cavg);
        const float g_ = __fmaf_rn(td.ncdot, xn.d, t_);
        w6 = __fsub_rn(w6, g_);
        ag.g6[j] = __fmaf_rn(td.sg, xo.d, ag.g6[j]);
      }
      if (valid) {
        *reinterpret_cast<float4*>(&sm.w[0][row * 4]) = make_float4(w01.x, w01.y, w23.x, w23.y);
        *reinterpret_cast<float4*>(&sm.w[1][row * 4]) = make_float4(w45.x, w45.y, w6, 0.0f);
      }
    }
    const float rs = valid ? rowdot2(xn, w01, w23, w45, w6) : 0.0f;
    const float rg = rowdot2(xn, ag.g01[j], ag.g23[j], ag.g45[j], ag.g6[j]);
    const float p = tree8_lanes(rs);
    if (pv && r == 0) sm.P[f] = p;
    if (j == 0) rg_acc = rv ? rg : 0.0f;
    else if (valid) rg_acc = __fadd_rn(rg_acc, rg);
  }
  // warp sum of x.wg: xor-butterfly 1,2,4,8,16
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) rg_acc = __fadd_rn(rg_acc, __shfl_xor_sync(0xffffffffu, rg_acc, s));
  if (lane_id() == 0) sm.Wg[warp_id()] = rg_acc;
}

// V(s) = x.w from the plane sums; every lane keeps the butterfly siblings it
// received (sib[0..4]) so that any lane's leaf can be substituted later.
struct VTree {
  float val, dotg;
  float sib[5];
};
__device__ __forceinline__ VTree v_tree(const SmemF& sm) {
  const int lane = lane_id();
  VTree t;
  const float p0 = sm.P[lane], p1 = sm.P[lane + 32];
  const float p2 = lane < 6 ? sm.P[lane + 64] : 0.0f;
  float b = __fadd_rn(__fadd_rn(p0, p1), p2);
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    t.sib[i] = __shfl_xor_sync(0xffffffffu, b, 1 << i);
    b = __fadd_rn(b, t.sib[i]);
  }
  t.val = __fadd_rn(b, sm.P[NCH]);
  t.dotg = __fadd_rn(__fadd_rn(sm.Wg[0], sm.Wg[1]), sm.Wg[2]);
  return t;
}

// q[k] for every candidate: dense dot of afterstate k in the FAST tree, by
// substituting the channel plane and the n_elig plane.  All 16 lane groups of the
// CTA take one candidate each per pass.
__device__ __forceinline__ void q_eval(SmemF& sm, const VTree& vt, const int n, const bool is_end,
                                       const uint2 n4b_r, const uint2 n2b_r) {
  const int grp = threadIdx.x >> 3, r = threadIdx.x & 7;
  const int wfirst = (threadIdx.x >> 5) * 4;  // first lane group of this warp
  if (wfirst >= n) return;  // warp-uniform: this warp has no candidate at all
  const bool rv = r < 7;
  const int rc = rv ? r : 6;
  const uint32_t want = is_end ? 1u : 0u;
  const uint2 x70 = *reinterpret_cast<const uint2*>(&sm.x[NCH * PLANE + rc * ROWPAD]);
  const float4 e_a = *reinterpret_cast<const float4*>(&sm.w[0][(NCH * 7 + rc) * 4]);
  const float4 e_b = *reinterpret_cast<const float4*>(&sm.w[1][(NCH * 7 + rc) * 4]);
  for (int base = 0; base + wfirst < n; base += 16) {  // warp-uniform trip count
    const int k = base + grp;
    const bool valid = k < n;
    const int ch = sm.cand[valid ? k : n - 1];
    const int row = ch * 7 + rc;
    // channel plane: x' = x +- [cell in N4 \ {cell}]   (Gridfuncs.hs:212-217)
    uint2 xr = *reinterpret_cast<const uint2*>(&sm.x[ch * PLANE + rc * ROWPAD]);
    xr = is_end ? bsub(xr, n4b_r) : badd(xr, n4b_r);
    const float4 wa = *reinterpret_cast<const float4*>(&sm.w[0][row * 4]);
    const float4 wb = *reinterpret_cast<const float4*>(&sm.w[1][row * 4]);
    const float rowP = rv ? rowdot2(cvt_row(xr), make_float2(wa.x, wa.y), make_float2(wa.z, wa.w),
                                    make_float2(wb.x, wb.y), wb.z) : 0.0f;
    // n_elig plane: e' = e -+ [cell in N2 u {cell} and ch eligible there on grid']  (Gridfuncs.hs:220-252)
    const uint2 cr = *reinterpret_cast<const uint2*>(&sm.cnt2[ch * PLANE + rc * ROWPAD]);
    const uint2 cg = make_uint2(n2b_r.x & eq_bytes(cr.x, want), n2b_r.y & eq_bytes(cr.y, want));
    const uint2 er = is_end ? badd(x70, cg) : bsub(x70, cg);
    const float rowE = rv ? rowdot2(cvt_row(er), make_float2(e_a.x, e_a.y), make_float2(e_a.z, e_a.w),
                                    make_float2(e_b.x, e_b.y), e_b.z) : 0.0f;
    const float Pp = tree8_lanes(rowP);
    const float Ep = tree8_lanes(rowE);
    // substitute leaf (ch & 31) of the V tree
    const int lc = ch & 31, hi = ch >> 5;
    const float p0 = hi == 0 ? Pp : sm.P[lc];
    const float p1 = hi == 1 ? Pp : sm.P[lc + 32];
    const float p2 = hi == 2 ? Pp : (lc < 6 ? sm.P[lc + 64] : 0.0f);
    float b = __fadd_rn(__fadd_rn(p0, p1), p2);
#pragma unroll
    for (int i = 0; i < 5; ++i) b = __fadd_rn(b, __shfl_sync(0xffffffffu, vt.sib[i], lc));
    const float qv = __fadd_rn(b, Ep);
    if (r == 0 && valid) sm.q[k] = qv;
  }
}

// argpmax (AccUtils.hs:203-208): the largest value, the LAST index among equals;
// NaNs follow the sequential right fold exactly.  Every warp computes it for itself.
__device__ __forceinline__ int argmax_last_warp(const float* q, int n) {
  const int lane = lane_id();
  float bq = -CUDART_INF_F;
  int bk = -1;
  bool nan = false;
  for (int k = lane; k < n; k += 32) {
    const float v = q[k];
    nan |= (v != v);
    if (v >= bq) { bq = v; bk = k; }
  }
  if (__any_sync(0xffffffffu, nan)) {
    int best = n - 1;
    for (int i = n - 2; i >= 0; --i)
      if (q[i] > q[best]) best = i;
    return best;
  }
  // order-preserving map float -> uint32, then two warp reductions
  uint32_t u = __float_as_uint(bq);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  if (bq == 0.0f) u = 0x80000000u;  // -0 == +0
  if (bk < 0) u = 0u;
  const uint32_t m = __reduce_max_sync(0xffffffffu, u);
  return __reduce_max_sync(0xffffffffu, (u == m) ? bk : -1);
}

// Which cells change their n_elig feature when `ch` is (de)allocated at the event
// cell (Gridfuncs.hs:220-232): members of N2 u {cell} with cnt2 == 0 (NEW/HOFF) or
// cnt2 == 1 (END).  One warp.
__device__ __forceinline__ unsigned long long elig_change_mask(const SmemF& sm, int ch, bool is_end,
                                                               unsigned long long n2mask) {
  const int lane = lane_id();
  const uint8_t want = is_end ? 1 : 0;
  const int c0 = lane, c1 = lane + 32;
  const bool b0 = ((n2mask >> c0) & 1ULL) && sm.cnt2[ch * PLANE + c_tab.pos[c0]] == want;
  const bool b1 = (c1 < NCELL) && ((n2mask >> c1) & 1ULL) && sm.cnt2[ch * PLANE + c_tab.pos[c1]] == want;
  const unsigned lo = __ballot_sync(0xffffffffu, b0), hi = __ballot_sync(0xffffffffu, b1);
  return (unsigned long long)lo | ((unsigned long long)hi << 32);
}

// executeAction (Gridfuncs.hs:105-110) + the in-place move of frep / cnt2 to the
// chosen afterstate (`ssFrep .= nextFrep`, SimRunner.hs:90).  Warp 0.
__device__ __forceinline__ void apply_action(SmemF& sm, int cell, bool is_end, int ch, const uint2 n4b_r,
                                             const uint2 n2b_r) {
  const int lane = lane_id();
  const unsigned long long chg = elig_change_mask(sm, ch, is_end, c_tab.n2incl[cell]);
  const int r = lane & 7, job = lane >> 3;
  const uint2 chg_r = r < 7 ? spread7((uint32_t)(chg >> (7 * r)) & 0x7Fu) : make_uint2(0, 0);
  if (lane < 8) sm.chgb[lane] = chg_r;
  if (lane == 0) {
    sm.chgmask = chg;
    const uint32_t bit = 1u << (ch & 31);
    if (is_end) sm.grid[cell][ch >> 5] &= ~bit;
    else sm.grid[cell][ch >> 5] |= bit;
  }
  if (r < 7 && job < 3) {
    if (job == 0) {  // in-use feature of ch at every cell of N4 \ {cell}
      uint2* p = reinterpret_cast<uint2*>(&sm.x[ch * PLANE + r * ROWPAD]);
      *p = is_end ? bsub(*p, n4b_r) : badd(*p, n4b_r);
    } else if (job == 1) {  // n_elig feature on the changing cells
      uint2* p = reinterpret_cast<uint2*>(&sm.x[NCH * PLANE + r * ROWPAD]);
      *p = is_end ? badd(*p, chg_r) : bsub(*p, chg_r);
    } else {  // cnt2 of ch over N2 u {cell}
      uint2* p = reinterpret_cast<uint2*>(&sm.cnt2[ch * PLANE + r * ROWPAD]);
      *p = is_end ? bsub(*p, n2b_r) : badd(*p, n2b_r);
    }
  }
}

// ---------------------------------------------------------------------------
// Event side (warp 3).  All registers are warp-uniform unless noted.
// ---------------------------------------------------------------------------
struct EvRegs {
  unsigned n_end, next_id;
  unsigned long long ndraws;
  long long iter, n_events, loss_n;
  float loss_sum, min_loss;
  double mean_new, call_dur, call_dur_hoff, hoff_prob;
  int ren_from, ren_to;                 // pending reassign (EventGen.hs:62-72) as q_key values; -1 = none
  unsigned long long pt;                // pre-scan: min (time, id) over the stored queue
  unsigned pid, pslot;
  unsigned long long dword;             // per lane: random word of draw ndraws + lane
  double dnl;                           // per lane: -log(u) of that word
  // random source
  int rng_mode;
  uint32_t k0, k1;
  unsigned long long stream;
  const unsigned long long* tape_w;
  const double* tape_l;
  unsigned long long tape_len;
};

__device__ __forceinline__ unsigned long long ev_word(const EvRegs& e, unsigned long long d) {
  if (e.rng_mode == RNG_PHILOX) return philox_word(d, e.stream, e.k0, e.k1);
  if (!e.tape_len) return 0ULL;
  return e.tape_w[d >= e.tape_len ? e.tape_len - 1 : d];
}
// lane l computes draw ndraws + l (word and -log u): all the draws an event can need
__device__ __forceinline__ void ev_draws(EvRegs& e) {
  const unsigned long long d = e.ndraws + (unsigned long long)lane_id();
  if (e.rng_mode == RNG_PHILOX) {
    e.dword = philox_word(d, e.stream, e.k0, e.k1);
    e.dnl = neglog_u53(e.dword);
  } else if (e.tape_len) {
    const unsigned long long i = d >= e.tape_len ? e.tape_len - 1 : d;
    e.dword = e.tape_w[i];
    e.dnl = e.tape_l[i];
  } else {
    e.dword = 0ULL;
    e.dnl = 0.0;
  }
}

// candidate channels of the event (getAction, Simulator.hs:161-165), ascending
// (indicesOf, AccUtils.hs:233-240), and the row byte masks of the event cell
__device__ __forceinline__ void ev_candidates(SmemF& sm) {
  const int lane = lane_id();
  const int cell = sm.ev.cell;
  const bool is_end = sm.ev.type == EV_END;
  const int pos = c_tab.pos[cell];
  uint32_t m[3];
  if (is_end) {  // inuseChs, Gridfuncs.hs:86-87
    m[0] = sm.grid[cell][0]; m[1] = sm.grid[cell][1]; m[2] = sm.grid[cell][2];
  } else {       // eligibleChs, Gridfuncs.hs:69-82  <=>  cnt2 == 0
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int ch = 32 * j + lane;
      const bool el = (ch < NCH) && (sm.cnt2[ch * PLANE + pos] == 0);
      m[j] = __ballot_sync(0xffffffffu, el);
    }
  }
  int base = 0;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if ((m[j] >> lane) & 1u) sm.cand[base + __popc(m[j] & ((1u << lane) - 1u))] = (uint8_t)(32 * j + lane);
    base += __popc(m[j]);
  }
  const unsigned long long n4 = c_tab.n4excl[cell], n2 = c_tab.n2incl[cell];
  if (lane < 8) {
    sm.n4b[lane] = (lane < 7) ? spread7((uint32_t)(n4 >> (7 * lane)) & 0x7Fu) : make_uint2(0, 0);
    sm.n2b[lane] = (lane < 7) ? spread7((uint32_t)(n2 >> (7 * lane)) & 0x7Fu) : make_uint2(0, 0);
  }
  if (lane == 0) {
    sm.ev.ncand = base;
    sm.n4mask = n4;
    sm.n2mask = n2;
  }
}

// pre-scan of the stored queue: argmin over (time, id) (pop, EventGen.hs:39-50),
// applying the pending re-key of `reassign` on the way
__device__ __forceinline__ void ev_prescan(SmemF& sm, EvRegs& e) {
  const unsigned total = QNEW + e.n_end;
  const int lane = lane_id();
  unsigned long long bt = 0xFFFFFFFFFFFFFFFFULL;
  unsigned bid = 0xFFFFFFFFu, bslot = 0xFFFFFFFFu;
  if (e.ren_from >= 0) {
    const uint16_t want = (uint16_t)e.ren_from, to = (uint16_t)e.ren_to;
    for (unsigned i = QNEW + lane; i < total; i += 32)
      if (sm.q_key[i] == want) sm.q_key[i] = to;
    e.ren_from = -1;
  }
  for (unsigned i = lane; i < total; i += 32) {
    const unsigned long long t = (unsigned long long)__double_as_longlong(sm.q_time[i]);
    const unsigned id = sm.q_id[i];
    if (t < bt || (t == bt && id < bid)) { bt = t; bid = id; bslot = i; }
  }
  const unsigned hi = (unsigned)(bt >> 32), lo = (unsigned)bt;
  const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
  bool ok = (hi == mhi);
  const unsigned mlo = __reduce_min_sync(0xffffffffu, ok ? lo : 0xFFFFFFFFu);
  ok = ok && (lo == mlo);
  const unsigned mid = __reduce_min_sync(0xffffffffu, ok ? bid : 0xFFFFFFFFu);
  ok = ok && (bid == mid);
  const unsigned src = __ffs(__ballot_sync(0xffffffffu, ok)) - 1;
  e.pt = ((unsigned long long)mhi << 32) | mlo;
  e.pid = mid;
  e.pslot = __shfl_sync(0xffffffffu, bslot, src);
}

// what the event warp prepares for the event in sm.ev while the agents run the
// dense pass: candidate channels, queue pre-scan, random draws
__device__ __forceinline__ void ev_prepare(SmemF& sm, EvRegs& e) {
  ev_candidates(sm);
  const int type = sm.ev.type;
  const bool pops = !(type == EV_END && sm.ev.hoff != HOFF_NONE);
  if (pops || e.ren_from >= 0) ev_prescan(sm, e);
  if (type != EV_END) ev_draws(e);
  __syncwarp();
}

// environmentStep (Simulator.hs:101-134) for the event in the registers
// (type, cell, ev_ch, ev_hoff, time) and the action `act` (-1 = Nothing);
// leaves the next event in sm.ev.
__device__ __forceinline__ void ev_environment_step(SmemF& sm, EvRegs& e, const int type, const int cell,
                                                    const int ev_ch, const int ev_hoff, const double time,
                                                    const int act) {
  const int lane = lane_id();
  if (lane == 0) {  // Stats hooks, Stats.hs:61-81 as called from Simulator.hs:105-114
    if (type == EV_NEW) {
      sm.stats[6]++; sm.stats[0]++;
      if (act >= 0) sm.stats[1]++;
      else { sm.stats[2]++; sm.stats[7]++; }
    } else if (type == EV_HOFF) {
      sm.stats[3]++;
      if (act < 0) { sm.stats[2]++; sm.stats[7]++; }  // Simulator.hs:113
    } else {
      sm.stats[4]++;
    }
  }
  bool have_new = false, have_end = false;
  unsigned long long t_new = 0, t_end = 0;
  unsigned id_new = 0, id_end = 0;
  int end_hoff = HOFF_NONE;
  if (type == EV_NEW) {
    // generateNewEvent, EventGen.hs:75-85: always, at time + Exp(mean_new)
    const double nl0 = __shfl_sync(0xffffffffu, e.dnl, 0);
    const double tn = __dadd_rn(time, __dmul_rn(nl0, e.mean_new));
    t_new = (unsigned long long)__double_as_longlong(tn);
    id_new = e.next_id++;
    have_new = true;
    if (lane == 0) { sm.q_time[cell] = tn; sm.q_id[cell] = id_new; }
    unsigned used = 1;
    if (act >= 0) {
      const unsigned long long w1 = __shfl_sync(0xffffffffu, e.dword, 1);
      const double p = __dmul_rn((double)(w1 >> 11), 1.0 / 9007199254740992.0);  // Simulator.hs:121
      const double nl2 = __shfl_sync(0xffffffffu, e.dnl, 2);
      const double te = __dadd_rn(time, __dmul_rn(nl2, e.call_dur));  // callDurNew for both kinds, EventGen.hs:102,112
      used = 3;
      if (p < e.hoff_prob) {  // generateHoffNewEvent, EventGen.hs:94-109
        const int m = c_tab.n2cnt[cell];
        const int n_reject = 256 % m;
        int z;
        for (;;) {  // random-fu integralUniform on the low byte, with rejection
          const unsigned long long wz = used < 32 ? __shfl_sync(0xffffffffu, e.dword, used) : ev_word(e, e.ndraws + used);
          ++used;
          z = (int)(wz & 0xFFULL);
          if (z >= n_reject) break;
          if (e.rng_mode == RNG_TAPE && e.ndraws + used >= e.tape_len) break;  // exhausted tape: stop is flagged
        }
        end_hoff = c_tab.n2list[cell][z % m];
      }
      t_end = (unsigned long long)__double_as_longlong(te);
      id_end = e.next_id++;
      if (end_hoff != HOFF_NONE) e.next_id++;  // the HOFF event's id (pushed right after its END, never stored)
      have_end = true;
    }
    e.ndraws += used;
  } else if (type == EV_HOFF) {
    if (act >= 0) {  // generateHoffEndEvent, EventGen.hs:114-116
      const double nl0 = __shfl_sync(0xffffffffu, e.dnl, 0);
      const double te = __dadd_rn(time, __dmul_rn(nl0, e.call_dur_hoff));
      t_end = (unsigned long long)__double_as_longlong(te);
      id_end = e.next_id++;
      have_end = true;
      e.ndraws += 1;
    }
  } else if (act >= 0 && act != ev_ch) {  // END toCh, Simulator.hs:129-130: reassign cell act -> ev_ch
    e.ren_from = (cell << 7) | act;
    e.ren_to = (cell << 7) | ev_ch;
  }
  if (have_end) {  // push END (cell, act) [hoff] at slot QNEW + n_end
    const unsigned slot = QNEW + e.n_end;
    if (lane == 0 && slot < QCAP) {
      sm.q_time[slot] = __longlong_as_double((long long)t_end);
      sm.q_id[slot] = id_end;
      sm.q_key[slot] = (uint16_t)((cell << 7) | act);
      sm.q_hoff[slot] = (uint8_t)end_hoff;
    }
    e.n_end++;
  }
  __syncwarp();
  // ---- the next event (Simulator.hs:132-134) ----
  double nx_time;
  int nx_type, nx_cell, nx_ch = -1, nx_hoff = HOFF_NONE;
  if (type == EV_END && ev_hoff != HOFF_NONE) {
    // the HOFF that follows a hand-off END: same time, next id (EventGen.hs:107-108)
    nx_time = time; nx_type = EV_HOFF; nx_cell = ev_hoff;
  } else {
    // min over (pre-scan, pushed NEW, pushed END) by (time, id)
    unsigned long long bt = e.pt;
    unsigned bid = e.pid;
    int src = 0;
    if (have_new && (t_new < bt || (t_new == bt && id_new < bid))) { bt = t_new; bid = id_new; src = 1; }
    if (have_end && (t_end < bt || (t_end == bt && id_end < bid))) { bt = t_end; bid = id_end; src = 2; }
    nx_time = __longlong_as_double((long long)bt);
    if (src == 2) {  // the END just pushed sits in the last slot: drop it again
      nx_type = EV_END; nx_cell = cell; nx_ch = act; nx_hoff = end_hoff;
      e.n_end--;
    } else {
      const unsigned slot = src == 1 ? (unsigned)cell : e.pslot;
      if (slot < QNEW) {
        nx_type = EV_NEW; nx_cell = (int)slot;
        if (lane == 0) sm.q_time[slot] = CUDART_INF;
      } else {
        int key = sm.q_key[slot];
        if (key == e.ren_from) key = e.ren_to;  // re-keyed by this very event, not yet written
        nx_type = EV_END; nx_cell = key >> 7; nx_ch = key & 127; nx_hoff = sm.q_hoff[slot];
        const unsigned last = QNEW + e.n_end - 1;
        __syncwarp();
        if (lane == 0 && slot != last) {
          sm.q_time[slot] = sm.q_time[last];
          sm.q_id[slot] = sm.q_id[last];
          sm.q_key[slot] = sm.q_key[last];
          sm.q_hoff[slot] = sm.q_hoff[last];
        }
        e.n_end--;
      }
    }
  }
  if (lane == 0) {
    sm.ev.time = nx_time; sm.ev.type = nx_type; sm.ev.cell = nx_cell; sm.ev.ch = nx_ch; sm.ev.hoff = nx_hoff;
  }
  __syncwarp();
}

// order-independent hashes of grid and frep for the parity trace (same as
// state_hashes); event warp only
__device__ __forceinline__ void ev_hashes(const SmemF& sm, unsigned long long& hg_out, unsigned long long& hf_out) {
  unsigned long long hg = 0, hf = 0;
  for (int i = lane_id(); i < NFEAT * NCELL; i += 32) {
    const int cell = i / NFEAT, f = i - cell * NFEAT;
    hf += (unsigned long long)sm.x[f * PLANE + c_tab.pos[cell]] * mix64(0x10000ULL + (unsigned long long)i);
    if (f < NCH && ((sm.grid[cell][f >> 5] >> (f & 31)) & 1u)) hg += mix64((unsigned long long)(cell * NCH + f));
  }
  for (int off = 16; off > 0; off >>= 1) {
    hg += __shfl_xor_sync(0xffffffffu, hg, off);
    hf += __shfl_xor_sync(0xffffffffu, hf, off);
  }
  hg_out = hg;
  hf_out = hf;
}

// violatesReuseConstraint (Gridfuncs.hs:94-101) and max(frep) > 70 (SimRunner.hs:150,158-159)
__device__ __forceinline__ int ev_verify(const SmemF& sm, bool rc, bool nb) {
  const int lane = lane_id();
  bool bad_rc = false, bad_nb = false;
  if (rc) {
    for (int cell = lane; cell < NCELL; cell += 32) {
      unsigned long long nbm = c_tab.n2incl[cell] & ~(1ULL << cell);
      uint32_t o0 = 0, o1 = 0, o2 = 0;
      while (nbm) {
        const int n = __ffsll((long long)nbm) - 1;
        nbm &= nbm - 1;
        o0 |= sm.grid[n][0]; o1 |= sm.grid[n][1]; o2 |= sm.grid[n][2];
      }
      bad_rc |= ((o0 & sm.grid[cell][0]) | (o1 & sm.grid[cell][1]) | (o2 & sm.grid[cell][2])) != 0;
    }
  }
  if (nb)
    for (int i = lane; i < WPAD; i += 32) bad_nb |= sm.x[i] > 70;
  if (__any_sync(0xffffffffu, bad_rc)) return ST_REUSE_VIOLATED;
  if (__any_sync(0xffffffffu, bad_nb)) return ST_NELIG_OVERFLOW;
  return 0;
}

// ---------------------------------------------------------------------------
// staging between the global replica blob (RepState layout) and the CTA
// ---------------------------------------------------------------------------
__device__ __forceinline__ void copy16(void* dst, const void* src, int bytes) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(dst);
  for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x) d[i] = s[i];
}
__device__ __forceinline__ void copy8(void* dst, const void* src, int bytes) {
  const uint2* s = reinterpret_cast<const uint2*>(src);
  uint2* d = reinterpret_cast<uint2*>(dst);
  for (int i = threadIdx.x; i < bytes / 8; i += blockDim.x) d[i] = s[i];
}
static_assert(sizeof(((RepState*)0)->x) % 8 == 0 && sizeof(((RepState*)0)->grid) % 16 == 0 &&
              sizeof(((RepState*)0)->q_time) % 16 == 0 && sizeof(((RepState*)0)->q_id) % 16 == 0 &&
              sizeof(((RepState*)0)->q_key) % 16 == 0 && sizeof(((RepState*)0)->q_hoff) % 8 == 0 &&
              offsetof(RepState, x) % 8 == 0 && offsetof(RepState, cnt2) % 8 == 0 && offsetof(RepState, grid) % 16 == 0 &&
              offsetof(RepState, q_time) % 16 == 0 && offsetof(RepState, q_id) % 16 == 0 &&
              offsetof(RepState, q_key) % 16 == 0 && offsetof(RepState, q_hoff) % 8 == 0, "blob arrays");

__device__ __forceinline__ void stage_in(SmemF& sm, Agent& ag, const RepState* g, int g8, int r) {
  for (int row = threadIdx.x; row < NROW; row += blockDim.x) {
    const int f = row / 7, rr = row - 7 * f;
    const float4* src = reinterpret_cast<const float4*>(&g->w[f * PLANE + rr * ROWPAD]);
    *reinterpret_cast<float4*>(&sm.w[0][row * 4]) = src[0];
    float4 b = src[1];
    b.w = 0.0f;
    *reinterpret_cast<float4*>(&sm.w[1][row * 4]) = b;
  }
  copy8(sm.x, g->x, sizeof sm.x);
  copy8(sm.cnt2, g->cnt2, sizeof sm.cnt2);
  copy16(sm.grid, g->grid, sizeof sm.grid);
  copy16(sm.q_time, g->q_time, sizeof sm.q_time);
  copy16(sm.q_id, g->q_id, sizeof sm.q_id);
  copy16(sm.q_key, g->q_key, sizeof sm.q_key);
  copy8(sm.q_hoff, g->q_hoff, sizeof sm.q_hoff);
  if (threadIdx.x < 8) sm.stats[threadIdx.x] = g->stats[threadIdx.x];
  if (threadIdx.x == 0) {
    sm.ev.time = g->ev_time; sm.ev.type = g->ev_type; sm.ev.cell = g->ev_cell; sm.ev.ch = g->ev_ch;
    sm.ev.hoff = g->ev_hoff; sm.ev.ncand = 0;
    sm.status = g->status;
    sm.chgmask = 0;
  }
  if (threadIdx.x < NAGENT) {  // (warp-uniform: false for the whole event warp)
    const int rc = r < 7 ? r : 6;
#pragma unroll
    for (int j = 0; j < NSLOT; ++j) {
      int f = SLOTP * j + g8;
      if (f > NCH) f = NCH;
      const float4* src = reinterpret_cast<const float4*>(&g->wg[f * PLANE + rc * ROWPAD]);
      const float4 a = src[0], b = src[1];
      ag.g01[j] = make_float2(a.x, a.y);
      ag.g23[j] = make_float2(a.z, a.w);
      ag.g45[j] = make_float2(b.x, b.y);
      ag.g6[j] = b.z;
    }
  }
}

__device__ __forceinline__ void stage_out(const SmemF& sm, const Agent& ag, RepState* g, int g8, int r) {
  for (int row = threadIdx.x; row < NROW; row += blockDim.x) {
    const int f = row / 7, rr = row - 7 * f;
    float4* dst = reinterpret_cast<float4*>(&g->w[f * PLANE + rr * ROWPAD]);
    dst[0] = *reinterpret_cast<const float4*>(&sm.w[0][row * 4]);
    dst[1] = *reinterpret_cast<const float4*>(&sm.w[1][row * 4]);
  }
  copy8(g->x, sm.x, sizeof sm.x);
  copy8(g->cnt2, sm.cnt2, sizeof sm.cnt2);
  copy16(g->grid, sm.grid, sizeof sm.grid);
  copy16(g->q_time, sm.q_time, sizeof sm.q_time);
  copy16(g->q_id, sm.q_id, sizeof sm.q_id);
  copy16(g->q_key, sm.q_key, sizeof sm.q_key);
  copy8(g->q_hoff, sm.q_hoff, sizeof sm.q_hoff);
  if (threadIdx.x < 8) g->stats[threadIdx.x] = sm.stats[threadIdx.x];
  if (threadIdx.x < NAGENT && r < 7) {
#pragma unroll
    for (int j = 0; j < NSLOT; ++j) {
      const int f = SLOTP * j + g8;
      if (f <= NCH) {
        float4* dst = reinterpret_cast<float4*>(&g->wg[f * PLANE + r * ROWPAD]);
        dst[0] = make_float4(ag.g01[j].x, ag.g01[j].y, ag.g23[j].x, ag.g23[j].y);
        dst[1] = make_float4(ag.g45[j].x, ag.g45[j].y, ag.g6[j], 0.0f);
      }
    }
  }
}

__device__ __forceinline__ void ev_load(EvRegs& e, const RunArgs& a, const RepState* g, int rep) {
  e.n_end = g->n_end; e.next_id = g->next_id; e.ndraws = g->ndraws;
  e.iter = g->iter; e.n_events = g->n_events; e.loss_n = g->loss_n;
  e.loss_sum = g->loss_sum; e.min_loss = g->min_loss;
  e.mean_new = g->mean_new; e.call_dur = g->call_dur; e.call_dur_hoff = g->call_dur_hoff; e.hoff_prob = g->hoff_prob;
  e.ren_from = -1; e.ren_to = -1;
  e.pt = 0xFFFFFFFFFFFFFFFFULL; e.pid = 0xFFFFFFFFu; e.pslot = 0;
  e.dword = 0; e.dnl = 0.0;
  e.rng_mode = a.rng_mode; e.k0 = a.key0; e.k1 = a.key1; e.stream = g->stream;
  e.tape_w = nullptr; e.tape_l = nullptr; e.tape_len = 0;
  if (a.rng_mode == RNG_TAPE && a.tape_off) {
    const unsigned long long o = a.tape_off[rep];
    e.tape_w = a.tape_words + o;
    e.tape_l = a.tape_neglog + o;
    e.tape_len = a.tape_off[rep + 1] - o;
  }
}

// ---------------------------------------------------------------------------
// The fused run kernel (dca_run, FAST order).  grid = n_replicas CTAs of 128.
// The agent warps and the event warp run their own copies of the event loop
// (separate register allocations) and meet at the three CTA barriers.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cta_barrier() { asm volatile("bar.sync 0, 128;" ::: "memory"); }

// what every thread derives from the Q phase: the action and the TD scalars
struct Decision {
  int act;
  float tderr;
  TdScalars td;
};
__device__ __forceinline__ Decision decide(const SmemF& sm, const VTree& vt, int n, bool is_end, float& avgR,
                                           int& ncalls, float alpha_net, float alpha_avg, float alpha_grad) {
  Decision d;
  d.act = -1;
  float next_val = vt.val;  // no action: nextFrep = frep (Simulator.hs:170-171)
  if (n > 0) {
    const int k = argmax_last_warp(sm.q, n);  // selectAction, Simulator.hs:202
    d.act = sm.cand[k];
    next_val = sm.q[k];
    ncalls += is_end ? -1 : 1;  // gridStep, Simulator.hs:137-144
  }
  // backward scalars (Agent.hs:62-67,75,78)
  const float reward = (float)ncalls;  // boolSum nextGrid
  d.tderr = __fsub_rn(__fadd_rn(__fsub_rn(reward, avgR), next_val), vt.val);
  const float c = __fmul_rn(-2.0f, alpha_net);
  d.td.ctd = __fmul_rn(c, d.tderr);
  d.td.cavg = __fmul_rn(c, avgR);
  d.td.ncdot = -__fmul_rn(c, vt.dotg);
  d.td.sg = __fmul_rn(alpha_grad, __fsub_rn(d.tderr, vt.dotg));
  avgR = __fadd_rn(avgR, __fmul_rn(alpha_avg, d.tderr));
  return d;
}

template <bool TRACE>
__device__ __forceinline__ void agent_main(SmemF& sm, const RunArgs& a, RepState* g) {
  const int wid = warp_id();
  const int g8 = threadIdx.x >> 3, r = threadIdx.x & 7, rc = r < 7 ? r : 6;
  Agent ag;
  stage_in(sm, ag, g, g8, r);
  float avgR = g->avg_reward;
  int ncalls = g->ncalls;
  const float alpha_net = g->alpha_net, alpha_avg = g->alpha_avg, alpha_grad = g->alpha_grad;
  cta_barrier();
  dense_pass<false>(sm, ag, g8, r, -1, false, TdScalars{0.f, 0.f, 0.f, 0.f}, make_uint2(0, 0));
  cta_barrier();  // B1
  for (long long it = 0; it < a.max_events; ++it) {
    const int cell = sm.ev.cell, n = sm.ev.ncand;
    const bool is_end = sm.ev.type == EV_END;
    const uint2 n4b_r = sm.n4b[rc], n2b_r = sm.n2b[rc];
    // accFn1 = getAction (Simulator.hs:154-172)
    const VTree vt = v_tree(sm);
    if (n > 0) q_eval(sm, vt, n, is_end, n4b_r, n2b_r);
    cta_barrier();  // B2
    const Decision d = decide(sm, vt, n, is_end, avgR, ncalls, alpha_net, alpha_avg, alpha_grad);
    if (wid == 0 && d.act >= 0) apply_action(sm, cell, is_end, d.act, n4b_r, n2b_r);
    cta_barrier();  // B3
    // accFn2's dense part (Agent.hs:67-76) + the next event's sums
    dense_pass<true>(sm, ag, g8, r, d.act, is_end, d.td, n4b_r);
    cta_barrier();  // B1
    if (sm.status != ST_RUNNING) break;  // uniform
  }
  stage_out(sm, ag, g, g8, r);
  if (threadIdx.x == 0) {
    g->avg_reward = avgR;
    g->ncalls = ncalls;
  }
}

template <bool TRACE>
__device__ __forceinline__ void event_main(SmemF& sm, const RunArgs& a, RepState* g, int rep) {
  const int lane = lane_id();
  const int rc = (lane & 7) < 7 ? (lane & 7) : 6;
  Agent none;  // (the event warp holds no weights)
  stage_in(sm, none, g, 0, 0);
  float avgR = g->avg_reward;
  int ncalls = g->ncalls;
  const float alpha_net = g->alpha_net, alpha_avg = g->alpha_avg, alpha_grad = g->alpha_grad;
  EvRegs e;
  ev_load(e, a, g, rep);
  cta_barrier();
  ev_prepare(sm, e);
  cta_barrier();  // B1
  for (long long it = 0; it < a.max_events; ++it) {
    const int type = sm.ev.type, cell = sm.ev.cell, n = sm.ev.ncand;
    const int ev_ch = sm.ev.ch, ev_hoff = sm.ev.hoff;
    const double time = sm.ev.time;
    const bool is_end = type == EV_END;
    const uint2 n4b_r = sm.n4b[rc], n2b_r = sm.n2b[rc];
    const VTree vt = v_tree(sm);
    if (n > 0) q_eval(sm, vt, n, is_end, n4b_r, n2b_r);
    cta_barrier();  // B2
    const Decision d = decide(sm, vt, n, is_end, avgR, ncalls, alpha_net, alpha_avg, alpha_grad);
    ev_environment_step(sm, e, type, cell, ev_ch, ev_hoff, time, d.act);  // Simulator.hs:101-134
    cta_barrier();  // B3
    e.iter += 1;  // Simulator.hs:131
    // stop rules of runPeriod (SimRunner.hs:150-169), in the reference's order
    int stop = ST_RUNNING;
    if (d.tderr != d.tderr) stop = ST_NAN_LOSS;
    else {
      e.loss_sum = __fadd_rn(e.loss_sum, d.tderr);
      e.loss_n += 1;
    }
    if (stop == ST_RUNNING && (a.verify_rc || a.verify_nb)) {
      const int v = ev_verify(sm, a.verify_rc, a.verify_nb);
      if (v) stop = v;
    }
    if (stop == ST_RUNNING && e.min_loss > 0.0f && e.iter > 50 && fabsf(d.tderr) < e.min_loss) stop = ST_ZERO_LOSS;
    if (stop == ST_RUNNING && e.iter == e.n_events) stop = ST_SUCCESS;
    if (stop == ST_RUNNING && e.rng_mode == RNG_TAPE && e.ndraws + 8 > e.tape_len) stop = ST_TAPE_EXHAUSTED;
    if (TRACE && a.trace && e.iter - 1 < a.trace_cap) {
      unsigned long long hg, hf;
      ev_hashes(sm, hg, hf);
      if (lane == 0) {
        TraceRec& tr = a.trace[(size_t)rep * a.trace_cap + (e.iter - 1)];
        tr.time = time;
        tr.loss = d.tderr;
        tr.avg_reward = avgR;
        tr.reward = ncalls;
        tr.etype = (uint8_t)type;
        tr.cell = (uint8_t)cell;
        tr.ch = (int8_t)d.act;
        tr.ncand = (uint8_t)n;
        tr.end_ch = (int8_t)(is_end ? ev_ch : -1);
        for (int i = 0; i < 7; ++i) tr.pad[i] = 0;
        tr.grid_hash = hg;
        tr.frep_hash = hf;
      }
    }
    if (lane == 0) sm.status = stop;
    ev_prepare(sm, e);
    cta_barrier();  // B1
    if (sm.status != ST_RUNNING) break;  // uniform
  }
  stage_out(sm, none, g, 0, 7);
  if (lane == 0) {
    g->n_end = e.n_end; g->next_id = e.next_id; g->ndraws = e.ndraws;
    g->iter = e.iter; g->loss_n = e.loss_n; g->loss_sum = e.loss_sum;
    g->ev_time = sm.ev.time; g->ev_type = sm.ev.type; g->ev_cell = sm.ev.cell; g->ev_ch = sm.ev.ch;
    g->ev_hoff = sm.ev.hoff;
    g->status = sm.status;
  }
}

template <bool TRACE>
__global__ void __launch_bounds__(NT, 5) k_run_fast(RunArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemF& sm = *reinterpret_cast<SmemF*>(smem_raw);
  const int rep = blockIdx.x;
  if (rep >= a.n_replicas) return;
  RepState* g = a.states + rep;
  if (!g->initialized || g->status != ST_RUNNING) return;  // uniform
  if (warp_id() == EVW) event_main<TRACE>(sm, a, g, rep);
  else agent_main<TRACE>(sm, a, g);
}

// ---------------------------------------------------------------------------
// Step-wise operators in the FAST order on one replica: accFn1 / accFn2
// (SimRunner.hs:61-62).  mode 0: getAction -> out->ch.  mode 1: gridStepTrain
// with the given channel.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) k_step_fast(RepState* states, int rep, int mode, int cell, int is_end_i,
                                                  int ch_in, float alpha_net, float alpha_avg, float alpha_grad,
                                                  StepOut* out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemF& sm = *reinterpret_cast<SmemF*>(smem_raw);
  RepState* g = states + rep;
  const bool is_end = is_end_i != 0;
  const int wid = warp_id(), lane = lane_id();
  const int g8 = threadIdx.x >> 3, r = threadIdx.x & 7, rc = r < 7 ? r : 6;
  Agent ag;
  stage_in(sm, ag, g, g8, r);
  float avgR = g->avg_reward;
  int ncalls = g->ncalls;
  __syncthreads();
  const TdScalars td0{0.f, 0.f, 0.f, 0.f};
  if (wid == EVW) {
    if (lane == 0) { sm.ev.cell = cell; sm.ev.type = is_end ? EV_END : EV_NEW; }
    __syncwarp();
    ev_candidates(sm);
    __syncwarp();
    if (mode == 1 && lane == 0) {  // evaluate exactly the channel the caller executes
      sm.ev.ncand = ch_in >= 0 ? 1 : 0;
      sm.cand[0] = (uint8_t)(ch_in >= 0 ? ch_in : 0);
    }
  } else {
    dense_pass<false>(sm, ag, g8, r, -1, false, td0, make_uint2(0, 0));
  }
  __syncthreads();
  const int n = sm.ev.ncand;
  const uint2 n4b_r = sm.n4b[rc], n2b_r = sm.n2b[rc];
  const VTree vt = v_tree(sm);
  if (n > 0) q_eval(sm, vt, n, is_end, n4b_r, n2b_r);
  __syncthreads();
  int act = -1;
  float next_val = vt.val;
  if (n > 0) {
    const int k = argmax_last_warp(sm.q, n);
    act = sm.cand[k];
    next_val = sm.q[k];
  }
  if (mode == 0) {
    if (threadIdx.x == 0) out->ch = act;
    return;
  }
  if (act >= 0) ncalls += is_end ? -1 : 1;
  const float reward = (float)ncalls;
  const float tderr = __fsub_rn(__fadd_rn(__fsub_rn(reward, avgR), next_val), vt.val);
  const float c = __fmul_rn(-2.0f, alpha_net);
  TdScalars td;
  td.ctd = __fmul_rn(c, tderr);
  td.cavg = __fmul_rn(c, avgR);
  td.ncdot = -__fmul_rn(c, vt.dotg);
  td.sg = __fmul_rn(alpha_grad, __fsub_rn(tderr, vt.dotg));
  avgR = __fadd_rn(avgR, __fmul_rn(alpha_avg, tderr));
  if (wid == 0 && act >= 0) apply_action(sm, cell, is_end, act, n4b_r, n2b_r);
  __syncthreads();
  if (wid != EVW) dense_pass<true>(sm, ag, g8, r, act, is_end, td, n4b_r);
  __syncthreads();
  // write back what the step changed (the event queue is untouched)
  for (int row = threadIdx.x; row < NROW; row += blockDim.x) {
    const int f = row / 7, rr = row - 7 * f;
    float4* dst = reinterpret_cast<float4*>(&g->w[f * PLANE + rr * ROWPAD]);
    dst[0] = *reinterpret_cast<const float4*>(&sm.w[0][row * 4]);
    dst[1] = *reinterpret_cast<const float4*>(&sm.w[1][row * 4]);
  }
  copy8(g->x, sm.x, sizeof sm.x);
  copy8(g->cnt2, sm.cnt2, sizeof sm.cnt2);
  copy16(g->grid, sm.grid, sizeof sm.grid);
  if (threadIdx.x < NAGENT && r < 7) {
#pragma unroll
    for (int j = 0; j < NSLOT; ++j) {
      const int f = SLOTP * j + g8;
      if (f <= NCH) {
        float4* dst = reinterpret_cast<float4*>(&g->wg[f * PLANE + r * ROWPAD]);
        dst[0] = make_float4(ag.g01[j].x, ag.g01[j].y, ag.g23[j].x, ag.g23[j].y);
        dst[1] = make_float4(ag.g45[j].x, ag.g45[j].y, ag.g6[j], 0.0f);
      }
    }
  }
  if (threadIdx.x == 0) {
    g->avg_reward = avgR;
    g->ncalls = ncalls;
    out->ch = act;
    out->loss = tderr;
    out->loss_is_nan = (tderr != tderr) ? 1 : 0;
    out->reward = ncalls;
  }
}

}  // namespace fast
}  // namespace dca
