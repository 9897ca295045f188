// dca_kernels.cuh -- sm_100a kernels shared by the whole library: the random source,
// small helpers, mkSimState (k_init), featureRep, the Gridfuncs operators, read-back,
// and the REFERENCE-ORDER run / step kernels (k_run<ORDER_REF>, k_step<ORDER_REF>):
// the Accelerate Interpreter's arithmetic, operation for operation, used for the
// --fixed_rng single-simulation replay.  The throughput kernel is in dca_fast.cuh.
//
// Reference-order kernels: one CTA per replica; the replica's state (RepState) is staged
// in shared memory for the launch.  Warp 0 runs the event-sequential part (candidate
// channels, argmax, grid/frep update, event queue, RNG, stats); one LANE per dense right
// fold (candidates, V(s), x.wGradCorr -- all of an event in lock step on one warp, see
// ref_fold_lane); the elementwise TDC update by rows of the padded layout.  The run kernel
// keeps wNet and wGradCorr in global memory (four CTAs per SM) and runs environmentStep next
// to the folds (env_ahead, warp 3) and the rest of warp 0's tail next to the update
// (run_events_overlapped); the step kernel and the TRACE instantiation keep the plain
// sequence.  Reference: runStep, src/SimRunner.hs:60-97.
//
// All parity-critical fp32/fp64 arithmetic uses explicit round-to-nearest
// intrinsics (never contracted); the file is also compiled with -fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <cstddef>

#include "dca_state.h"

namespace dca {

__constant__ Tables c_tab;

#ifdef DCA_PHASE_CLOCK  // development only: cycles per phase of the REF-order run kernel, summed over CTAs and events
__device__ unsigned long long g_ref_phase[8];
#define DCA_RCLK(i)                                                              \
  if (threadIdx.x == 0) {                                                        \
    const unsigned now_ = (unsigned)clock64();                                   \
    atomicAdd(&g_ref_phase[i], (unsigned long long)(now_ - rclk_));              \
    rclk_ = now_;                                                                \
  }
#define DCA_RCLK_START unsigned rclk_ = (unsigned)clock64();
#else
#define DCA_RCLK(i)
#define DCA_RCLK_START
#endif

struct RunArgs {
  RepState* states;
  int n_replicas;
  long long max_events;
  int rng_mode;
  uint32_t key0, key1;              // Philox key = seed
  int verify_rc, verify_nb;
  // tape (RNG_TAPE): per replica offsets into the word / neglog arrays
  const unsigned long long* tape_words;
  const double* tape_neglog;
  const unsigned long long* tape_off;  // [n_replicas + 1]
  TraceRec* trace;                  // [n_replicas][trace_cap] or null
  int trace_cap;
  int lookahead;                    // hand-off look-ahead (extension; both orders: lookahead_folds here, k_run_fast<true, true>)
  // FAST kernel: the launch is cut into units (replica, chunk of chunk_events events), one CTA each -- see k_run_fast
  unsigned* chunk_done;             // [n_replicas]: chunks of the replica completed in this launch (zeroed before the launch)
  long long chunk_events;
  unsigned n_chunks;
};

constexpr int PROW = 72;        // floats per cell in the product arrays: 71 features + 1 pad, so that a row starts on 16 bytes
constexpr int LA_PAIRS = 504;   // 4 rounds of 126 folding threads (a candidate has at most 71 pairs)
constexpr int LA_NONE = 255;
// Scratch that lives next to RepState in shared memory (reference-order kernels).
struct Scratch {
  float q[72];     // afterstate values of the candidates
  uint2 n4b[8];    // row byte masks: N4(cell) \ {cell}
  uint2 n2b[8];    // row byte masks: N2(cell) u {cell}
  uint2 chgb[8];   // row byte masks: cells whose n_elig changes (chosen action)
  uint2 n4b_upd[8];  // n4b of the event whose update is running (the run kernel builds the next event's masks meanwhile)
  // run kernel, environmentStep ahead of the decision (env_ahead / env_post): the current event as env_ahead found it, the
  // END event it pushed without a channel, and what warp 3 -- the owner of the queue counters and the random source --
  // publishes for warp 0's stop rules
  int cur_type, cur_ch, cur_hoff, pend_slot, q_overflow;
  unsigned q_n_end;
  unsigned long long rng_ndraws;
  unsigned long long n4mask, n2mask, chgmask;
  unsigned long long hash_grid, hash_frep;
  float ctd, cavg, cdot, sg;   // TDC scalars of this event
  float val, dotg, loss;
  int act_ch, act_diff, ncand, stop;
  int ev_cell, ev_is_end, ev_type_prev, ev_ch_prev;
  double ev_time_prev;
  int flag;
  uint8_t cand[72];
  // hand-off look-ahead (extension): the (channel to free, channel to assign in the target cell) pairs of one round
  float q_plain[72];           // V(afterstate_k): the TD target stays the real next state
  float pair_q[LA_PAIRS];
  uint8_t pair_k[LA_PAIRS];    // index into cand
  uint8_t pair_h[LA_PAIRS];    // channel in the target cell, LA_NONE = the plain afterstate of the END
  int la_npairs, la_knext;
  // LAST: the products float(x_i) * wNet_i and float(x_i) * wGradCorr_i of the CURRENT state in flatten order
  // i = cell * 71 + f, formed by all threads once per event: every dense right fold of the event (V(s), x.wGradCorr, one
  // per candidate) reads them as broadcasts and only re-forms the <= 61 terms its afterstate changes.  These two arrays
  // are the STEP kernel's.  The run kernel does not allocate them: it keeps wNet and wGradCorr themselves in global
  // memory (the update reads and writes them once per event, coalesced, served by L2; a fold needs two weights per cell,
  // requested three cells ahead) and the products in the shared-memory bytes of RepState::w / RepState::wg -- 55 KB
  // instead of 86 KB per CTA, four CTAs per SM.
  alignas(16) float prod_w[PROW * NCELL];
  alignas(16) float prod_g[PROW * NCELL];
};
// where the reference-order kernels keep the weights and their products (see Scratch::prod_w)
struct RefMem {
  const float* w;   // wNet, padded plane-major layout (shared memory in the step kernel, global in the run kernel)
  RepState* gs;     // the blob whose w / wg the update reads and writes
  float* prod_w;    // products with wNet, flatten order (shared memory)
  float* prod_g;    // products with wGradCorr
};

struct Smem {
  RepState s;
  Scratch t;
};
// shared memory of the run kernel: everything up to Scratch::prod_w
constexpr size_t SMEM_RUN_BYTES = (offsetof(Smem, t) + offsetof(Scratch, prod_w) + 15) / 16 * 16;
static_assert(sizeof(((RepState*)0)->wg) >= sizeof(float) * PROW * NCELL && sizeof(((RepState*)0)->w) >= sizeof(float) * PROW * NCELL,
              "the run kernel's products live in RepState::w / RepState::wg");
static_assert(4 * (SMEM_RUN_BYTES + 1024) <= 232448, "four CTAs of the reference-order run kernel per SM");

// ---------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
// offset of cell (r, c) = r * 7 + c inside a padded plane: r * 8 + c = cell + cell / 7  (arithmetic: a constant-memory
// table indexed differently by every lane of a warp is served one address at a time)
__device__ __forceinline__ int pos_of(const int cell) { return cell + ((cell * 37) >> 8); }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// 7 mask bits -> 7 bytes of 0/1 (uint2, byte c = bit c)
__device__ __forceinline__ uint2 spread7(uint32_t bits) {
  uint2 r;
  r.x = ((bits & 0xFu) * 0x00204081u) & 0x01010101u;
  r.y = (((bits >> 4) & 0x7u) * 0x00204081u) & 0x00010101u;
  return r;
}
// per byte (values < 0x80): 0x01 where byte == k (k in {0,1}), else 0
__device__ __forceinline__ uint32_t eq_bytes(uint32_t v, uint32_t k) {
  uint32_t t = v ^ (k * 0x01010101u);
  return (~(t + 0x7F7F7F7Fu) >> 7) & 0x01010101u;
}
__device__ __forceinline__ float byte_f(uint32_t v, int i) {
  return (float)((v >> (8 * i)) & 0xFFu);
}

// tree8 over 8 consecutive lanes (xor butterfly); every lane gets the result
__device__ __forceinline__ float tree8_lanes(float v) {
  v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 2));
  v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 4));
  return v;
}
// ---------------------------------------------------------------------------
// RNG: Philox4x32-10 (counter = draw index, stream id; key = seed) and the
// operation-by-operation -log(u) of the Philox mode (DESIGN.md "Random numbers")
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long philox_word(unsigned long long d,
                                                          unsigned long long stream, uint32_t k0,
                                                          uint32_t k1) {
  uint32_t c0 = (uint32_t)d, c1 = (uint32_t)(d >> 32), c2 = (uint32_t)stream,
           c3 = (uint32_t)(stream >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return (unsigned long long)c0 | ((unsigned long long)c1 << 32);
}

__device__ __noinline__ double neglog_u53(unsigned long long word) {
  unsigned long long k = word >> 11;
  if (k == 0) return CUDART_INF;
  double u = __dmul_rn((double)k, 1.0 / 9007199254740992.0);
  unsigned long long bits = (unsigned long long)__double_as_longlong(u);
  int e = (int)((bits >> 52) & 0x7FF) - 1022;
  bits = (bits & 0x000FFFFFFFFFFFFFULL) | 0x3FE0000000000000ULL;
  double m = __longlong_as_double((long long)bits);
  if (m < 0.70710678118654752440) {
    m = __dadd_rn(m, m);
    e -= 1;
  }
  double s = __ddiv_rn(__dsub_rn(m, 1.0), __dadd_rn(m, 1.0));
  double z = __dmul_rn(s, s);
  double p = 1.0 / 21.0;
  p = __fma_rn(p, z, 1.0 / 19.0);
  p = __fma_rn(p, z, 1.0 / 17.0);
  p = __fma_rn(p, z, 1.0 / 15.0);
  p = __fma_rn(p, z, 1.0 / 13.0);
  p = __fma_rn(p, z, 1.0 / 11.0);
  p = __fma_rn(p, z, 1.0 / 9.0);
  p = __fma_rn(p, z, 1.0 / 7.0);
  p = __fma_rn(p, z, 1.0 / 5.0);
  p = __fma_rn(p, z, 1.0 / 3.0);
  p = __fma_rn(p, z, 1.0);
  double t = __dmul_rn(s, p);
  t = __dadd_rn(t, t);
  return -__fma_rn((double)e, 0.69314718055994530942, t);
}

// Random source of one replica; all lanes of warp 0 hold identical copies.
struct Rng {
  int mode;
  uint32_t k0, k1;
  unsigned long long stream;
  unsigned long long ndraws;
  const unsigned long long* words;  // tape
  const double* neglog;
  unsigned long long tape_len;

  __device__ __forceinline__ unsigned long long idx() const {
    return (mode == RNG_TAPE && ndraws >= tape_len) ? (tape_len ? tape_len - 1 : 0) : ndraws;
  }
  // getRandomWord64 (EventGen/Internal.hs:60)
  __device__ __forceinline__ unsigned long long word() {
    unsigned long long w;
    if (mode == RNG_PHILOX) w = philox_word(ndraws, stream, k0, k1);
    else w = tape_len ? words[idx()] : 0ULL;
    ++ndraws;
    return w;
  }
  // getRandomDouble = (w >> 11) / 2^53 (EventGen/Internal.hs:61)
  __device__ __forceinline__ double uniform() {
    return __dmul_rn((double)(word() >> 11), 1.0 / 9007199254740992.0);
  }
  // random-fu exponential mu = (-log u) * mu, one draw
  __device__ __forceinline__ double exponential(double mean) {
    double nl;
    if (mode == RNG_PHILOX) nl = neglog_u53(word());
    else {
      nl = tape_len ? neglog[idx()] : 0.0;
      ++ndraws;
    }
    return __dmul_rn(nl, mean);
  }
  // random-fu integralUniform on the low byte, with rejection (EventGen.hs:104)
  __device__ __forceinline__ int uniform_int(int m) {
    int n_reject = 256 % m;
    for (;;) {
      int z = (int)(word() & 0xFFULL);
      if (z >= n_reject) return z % m;
      if (mode == RNG_TAPE && ndraws >= tape_len) return z % m;  // exhausted tape: stop is flagged
    }
  }
};

// ---------------------------------------------------------------------------
// Event queue (warp 0).  Hot scalars are kept uniformly in registers.
// ---------------------------------------------------------------------------
struct QRegs {
  unsigned n_end, next_id;
  bool overflow = false;  // an END event did not fit (more than QEND calls in progress): the replica stops
};

__device__ __forceinline__ void q_push_new(RepState& s, QRegs& q, int cell, double t) {
  if (lane_id() == 0) {
    s.q_time[cell] = t;
    s.q_id[cell] = q.next_id;
  }
  q.next_id++;
}
__device__ __forceinline__ void q_push_end(RepState& s, QRegs& q, int cell, int ch, int hoff,
                                           double t) {
  unsigned slot = QNEW + q.n_end;
  if (slot >= QCAP) {  // (uniform) the table is full: nothing is written, n_end stays in bounds
    q.overflow = true;
    q.next_id++;
    return;
  }
  if (lane_id() == 0) {
    s.q_time[slot] = t;
    s.q_id[slot] = q.next_id;
    s.q_key[slot] = (uint16_t)((cell << 7) | ch);
    s.q_hoff[slot] = (uint8_t)hoff;
  }
  q.n_end++;
  q.next_id++;
}
// reassign (EventGen.hs:62-72): re-key the END event of (cell, from) to (cell, to)
__device__ __forceinline__ void q_reassign(RepState& s, const QRegs& q, int cell, int from,
                                           int to) {
  const uint16_t want = (uint16_t)((cell << 7) | from);
  for (unsigned i = QNEW + lane_id(); i < QNEW + q.n_end; i += 32)
    if (s.q_key[i] == want) s.q_key[i] = (uint16_t)((cell << 7) | to);
}
// pop (EventGen.hs:39-50): argmin over (time, id); writes the event into s.ev_*
// (`pend`: slot of an END event whose key still lacks its channel -- env_ahead -- or null: follows the entry when the
//  pop moves it, -1 when it is the event popped)
__device__ __forceinline__ void q_pop(RepState& s, QRegs& q, int* pend = nullptr) {
  const unsigned total = QNEW + q.n_end;
  const int lane = lane_id();
  unsigned long long bt = 0xFFFFFFFFFFFFFFFFULL;
  unsigned bid = 0xFFFFFFFFu, bslot = 0xFFFFFFFFu;
  for (unsigned i = lane; i < total; i += 32) {
    unsigned long long t = (unsigned long long)__double_as_longlong(s.q_time[i]);
    unsigned id = s.q_id[i];
    if (t < bt || (t == bt && id < bid)) {
      bt = t; bid = id; bslot = i;
    }
  }
  // warp argmin with 32-bit redux: time hi, time lo, id
  unsigned hi = (unsigned)(bt >> 32), lo = (unsigned)bt;
  unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
  bool ok = (hi == mhi);
  unsigned mlo = __reduce_min_sync(0xffffffffu, ok ? lo : 0xFFFFFFFFu);
  ok = ok && (lo == mlo);
  unsigned mid = __reduce_min_sync(0xffffffffu, ok ? bid : 0xFFFFFFFFu);
  ok = ok && (bid == mid);
  unsigned src = __ffs(__ballot_sync(0xffffffffu, ok)) - 1;
  unsigned slot = __shfl_sync(0xffffffffu, bslot, src);
  if (slot < QNEW) {
    if (lane == 0) {
      s.ev_time = s.q_time[slot];
      s.ev_type = EV_NEW;
      s.ev_cell = (int)slot;
      s.ev_ch = -1;
      s.ev_hoff = HOFF_NONE;
      s.q_time[slot] = CUDART_INF;
    }
  } else {
    unsigned last = QNEW + q.n_end - 1;
    if (pend) {
      if ((int)slot == *pend) *pend = -1;
      else if ((int)last == *pend) *pend = (int)slot;
    }
    if (lane == 0) {
      unsigned key = s.q_key[slot];
      s.ev_time = s.q_time[slot];
      s.ev_type = EV_END;
      s.ev_cell = (int)(key >> 7);
      s.ev_ch = (int)(key & 127u);
      s.ev_hoff = s.q_hoff[slot];
      if (slot != last) {
        s.q_time[slot] = s.q_time[last];
        s.q_id[slot] = s.q_id[last];
        s.q_key[slot] = s.q_key[last];
        s.q_hoff[slot] = s.q_hoff[last];
      }
    }
    q.n_end--;
  }
  __syncwarp();
}

// environmentStep (Simulator.hs:101-134) for the current event and action `act`
// (-1 = Nothing); afterwards s.ev_* holds the next event.  Warp 0 only.
__device__ __forceinline__ void environment_step(RepState& s, QRegs& q, Rng& rng, int act) {
  const int lane = lane_id();
  const int type = s.ev_type, cell = s.ev_cell, ev_ch = s.ev_ch, ev_hoff = s.ev_hoff;
  const double time = s.ev_time;
  __syncwarp();
  if (lane == 0) {  // Stats hooks, Stats.hs:61-81 as called from Simulator.hs:105-114
    if (type == EV_NEW) {
      s.stats[6]++; s.stats[0]++;
      if (act >= 0) s.stats[1]++;
      else { s.stats[2]++; s.stats[7]++; }
    } else if (type == EV_HOFF) {
      s.stats[3]++;
      if (act < 0) { s.stats[2]++; s.stats[7]++; }  // Simulator.hs:113
    } else {
      s.stats[4]++;
    }
  }
  if (type == EV_NEW) {
    double dt = rng.exponential(s.mean_new);  // generateNewEvent, EventGen.hs:75-85
    q_push_new(s, q, cell, __dadd_rn(time, dt));
    if (act >= 0) {
      double p = rng.uniform();  // Simulator.hs:121, drawn even when hoffProb == 0
      if (p < s.hoff_prob) {     // generateHoffNewEvent, EventGen.hs:94-109
        double d2 = rng.exponential(s.call_dur);
        int m = c_tab.n2cnt[cell];
        int to = c_tab.n2list[cell][rng.uniform_int(m)];
        q_push_end(s, q, cell, act, to, __dadd_rn(time, d2));
        q.next_id++;  // the HOFF event's id (pushed right after its END, never stored)
      } else {                   // generateEndEvent, EventGen.hs:111-112
        double d2 = rng.exponential(s.call_dur);
        q_push_end(s, q, cell, act, HOFF_NONE, __dadd_rn(time, d2));
      }
    }
  } else if (type == EV_HOFF) {
    if (act >= 0) {              // generateHoffEndEvent, EventGen.hs:114-116
      double d2 = rng.exponential(s.call_dur_hoff);
      q_push_end(s, q, cell, act, HOFF_NONE, __dadd_rn(time, d2));
    }
  } else {                       // END toCh _, Simulator.hs:129-130
    if (act >= 0 && act != ev_ch) q_reassign(s, q, cell, act, ev_ch);
  }
  __syncwarp();
  // pop the next event (Simulator.hs:132-134)
  if (type == EV_END && ev_hoff != HOFF_NONE) {
    if (lane == 0) {  // the HOFF that follows a hand-off END: same time, next id
      s.ev_type = EV_HOFF;
      s.ev_cell = ev_hoff;
      s.ev_ch = -1;
      s.ev_hoff = HOFF_NONE;
    }
    __syncwarp();
  } else {
    q_pop(s, q);
  }
}

// ---------------------------------------------------------------------------
// candidate channels (getAction, Simulator.hs:161-165); warp 0
// returns n; fills t.cand in ascending channel order (indicesOf, AccUtils.hs:233-240)
// ---------------------------------------------------------------------------
__device__ __forceinline__ int build_candidates(Smem& sm, int cell, bool is_end) {
  RepState& s = sm.s;
  const int lane = lane_id();
  const int pos = pos_of(cell);
  uint32_t m[3];
  if (is_end) {  // inuseChs, Gridfuncs.hs:86-87
    m[0] = s.grid[cell][0]; m[1] = s.grid[cell][1]; m[2] = s.grid[cell][2];
  } else {       // eligibleChs, Gridfuncs.hs:69-82  <=>  cnt2 == 0
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      int ch = 32 * j + lane;
      bool e = (ch < NCH) && (s.cnt2[ch * PLANE + pos] == 0);
      m[j] = __ballot_sync(0xffffffffu, e);
    }
  }
  int base = 0;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if ((m[j] >> lane) & 1u) sm.t.cand[base + __popc(m[j] & ((1u << lane) - 1u))] = (uint8_t)(32 * j + lane);
    base += __popc(m[j]);
  }
  // row byte masks of the event cell's neighbourhoods
  const unsigned long long n4 = c_tab.n4excl[cell], n2 = c_tab.n2incl[cell];
  if (lane < 8) {
    sm.t.n4b[lane] = (lane < 7) ? spread7((uint32_t)(n4 >> (7 * lane)) & 0x7Fu) : make_uint2(0, 0);
    sm.t.n2b[lane] = (lane < 7) ? spread7((uint32_t)(n2 >> (7 * lane)) & 0x7Fu) : make_uint2(0, 0);
  }
  if (lane == 0) {
    sm.t.n4mask = n4;
    sm.t.n2mask = n2;
  }
  __syncwarp();
  return base;
}

// argpmax (AccUtils.hs:203-208): the largest value, the LAST index among equals.
// NaNs follow the sequential right fold exactly (rare path).  Warp 0.
__device__ __forceinline__ int argmax_last(const float* q, int n) {
  const int lane = lane_id();
  float bq = -CUDART_INF_F;
  int bk = -1;
  bool nan = false;
  for (int k = lane; k < n; k += 32) {
    float v = q[k];
    nan |= (v != v);
    if (v >= bq) { bq = v; bk = k; }
  }
  if (__any_sync(0xffffffffu, nan)) {
    int best = n - 1;
    for (int i = n - 2; i >= 0; --i)
      if (q[i] > q[best]) best = i;
    return best;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    float oq = __shfl_xor_sync(0xffffffffu, bq, off);
    int ok = __shfl_xor_sync(0xffffffffu, bk, off);
    if (oq > bq || (oq == bq && ok > bk)) { bq = oq; bk = ok; }
  }
  return bk;
}

// Which cells change their n_elig feature when `ch` is (de)allocated at the event
// cell: N2 u {cell} members where ch is eligible on grid' (Gridfuncs.hs:220-232):
// NEW/HOFF: cnt2 == 0;  END (cell's own use removed): cnt2 == 1.   Warp 0.
__device__ __forceinline__ unsigned long long elig_change_mask(const RepState& s, int ch,
                                                               bool is_end,
                                                               unsigned long long n2mask) {
  const int lane = lane_id();
  const uint8_t want = is_end ? 1 : 0;
  int c0 = lane, c1 = lane + 32;
  bool b0 = ((n2mask >> c0) & 1ULL) && s.cnt2[ch * PLANE + pos_of(c0)] == want;
  bool b1 = (c1 < NCELL) && ((n2mask >> c1) & 1ULL) && s.cnt2[ch * PLANE + pos_of(c1)] == want;
  unsigned lo = __ballot_sync(0xffffffffu, b0), hi = __ballot_sync(0xffffffffu, b1);
  return (unsigned long long)lo | ((unsigned long long)hi << 32);
}

// executeAction (Gridfuncs.hs:105-110) + in-place frep/cnt2 update to the chosen
// afterstate (what `ssFrep .= nextFrep` amounts to).  Warp 0.
__device__ __forceinline__ void apply_action(Smem& sm, int cell, bool is_end, int ch,
                                             unsigned long long chg) {
  RepState& s = sm.s;
  const int lane = lane_id();
  if (lane < 8) sm.t.chgb[lane] = (lane < 7) ? spread7((uint32_t)(chg >> (7 * lane)) & 0x7Fu) : make_uint2(0, 0);
  if (lane == 0) {
    sm.t.chgmask = chg;
    uint32_t bit = 1u << (ch & 31);
    if (is_end) s.grid[cell][ch >> 5] &= ~bit;
    else s.grid[cell][ch >> 5] |= bit;
    s.ncalls += is_end ? -1 : 1;
  }
  __syncwarp();
  // three row-parallel updates on lanes 0-6 / 8-14 / 16-22
  const int r = lane & 7, job = lane >> 3;
  if (r < 7 && job < 3) {
    if (job == 0) {  // in-use feature of ch at every cell of N4 \ {cell}: +-1
      uint2* p = reinterpret_cast<uint2*>(&s.x[ch * PLANE + r * ROWPAD]);
      uint2 v = *p, d = sm.t.n4b[r];
      if (is_end) { v.x -= d.x; v.y -= d.y; } else { v.x += d.x; v.y += d.y; }
      *p = v;
    } else if (job == 1) {  // n_elig feature: -+1 on the changing cells
      uint2* p = reinterpret_cast<uint2*>(&s.x[NCH * PLANE + r * ROWPAD]);
      uint2 v = *p, d = sm.t.chgb[r];
      if (is_end) { v.x += d.x; v.y += d.y; } else { v.x -= d.x; v.y -= d.y; }
      *p = v;
    } else {  // cnt2 of ch over N2 u {cell}
      uint2* p = reinterpret_cast<uint2*>(&s.cnt2[ch * PLANE + r * ROWPAD]);
      uint2 v = *p, d = sm.t.n2b[r];
      if (is_end) { v.x -= d.x; v.y -= d.y; } else { v.x += d.x; v.y += d.y; }
      *p = v;
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------
// REF order: reference-literal arithmetic (Accelerate Interpreter)
// ---------------------------------------------------------------------------
// The dense dots of the REF order are right folds over flatten order i = cell*71 + f, i = 3478..0 (stored in rows of PROW floats), of
// float(afrep[i]) * w[i] with products and sums rounded separately (the Accelerate Interpreter).  The additions of a
// fold are inherently sequential; the products are not: they are formed once per event by all threads (ref_update_rows:
// by the update of the previous event, or alone at launch start), and a fold is then 3479 dependent FADDs fed by
// shared-memory broadcasts (ref_fold_lane; ref_fold_plain / ref_fold_two serve the hand-off look-ahead).
// x . w of the current state: p0 + (p1 + (... + (p3478 + 0)))
__device__ __noinline__ float ref_fold_plain(const float* p) {
  float acc = 0.0f;
#pragma unroll 1
  for (int cell = NCELL - 1; cell >= 0; --cell) {
    const float* pc = p + cell * PROW;
    acc = __fadd_rn(pc[NCH], acc);
#pragma unroll
    for (int f0 = NCH - 10; f0 >= 0; f0 -= 10) {
      float v[10];
#pragma unroll
      for (int k = 0; k < 10; ++k) v[k] = pc[f0 + k];
#pragma unroll
      for (int k = 9; k >= 0; --k) acc = __fadd_rn(v[k], acc);
    }
  }
  return acc;
}
// Hand-off look-ahead (extension, specified at dca_config.hoff_lookahead in include/dca_b200.h): the dense right fold of the
// state after END `ch1` at `cell` AND, if ch2 >= 0, the HOFF arrival taking `ch2` at `cell2` on that afterstate.
// Per cell the n_elig term, the term of ch1 and the term of ch2 are re-formed when the two actions change them:
//   in-use features: ch1 loses its user at cell (-1 on N4(cell) \ {cell}), ch2 gains one at cell2 (+1 on N4(cell2) \ {cell2});
//   n_elig: +1 where ch1 becomes eligible (cnt2[ch1] == 1 within N2(cell) u {cell}), then -1 where ch2 was eligible on
//   that afterstate within N2(cell2) u {cell2} (cnt2 of ch2 after the END: one less inside N2(cell) if ch2 == ch1).
__device__ __noinline__ float ref_fold_two(const Smem& sm, const RefMem& m, const int cell1, const int ch1, const int cell2, const int ch2) {
  const RepState& s = sm.s;
  const unsigned long long n4a = c_tab.n4excl[cell1], n2a = c_tab.n2incl[cell1];
  const unsigned long long n4b = ch2 >= 0 ? c_tab.n4excl[cell2] : 0ULL, n2b = ch2 >= 0 ? c_tab.n2incl[cell2] : 0ULL;
  const int c2 = ch2 >= 0 ? ch2 : ch1;  // (no second action: its masks are empty)
  float acc = 0.0f;
#pragma unroll 1
  for (int cell = NCELL - 1; cell >= 0; --cell) {
    const int pos = pos_of(cell);
    const float* p = &m.prod_w[cell * PROW];
    const bool in2a = (n2a >> cell) & 1ULL, in2b = (n2b >> cell) & 1ULL;
    int d70 = 0;
    if (in2a && s.cnt2[ch1 * PLANE + pos] == 1) d70 += 1;
    if (in2b && (int)s.cnt2[c2 * PLANE + pos] - ((c2 == ch1 && in2a) ? 1 : 0) == 0) d70 -= 1;
    float t70 = p[NCH];
    if (d70 != 0) t70 = __fmul_rn((float)((int)s.x[NCH * PLANE + pos] + d70), m.w[NCH * PLANE + pos]);
    const int d1 = (((n4a >> cell) & 1ULL) ? -1 : 0) + ((c2 == ch1 && ((n4b >> cell) & 1ULL)) ? 1 : 0);
    float t1 = p[ch1];
    if (d1 != 0) t1 = __fmul_rn((float)((int)s.x[ch1 * PLANE + pos] + d1), m.w[ch1 * PLANE + pos]);
    float t2 = p[c2];
    if (c2 != ch1 && ((n4b >> cell) & 1ULL)) t2 = __fmul_rn((float)((int)s.x[c2 * PLANE + pos] + 1), m.w[c2 * PLANE + pos]);
    acc = __fadd_rn(t70, acc);
#pragma unroll
    for (int f0 = NCH - 10; f0 >= 0; f0 -= 10) {
      float v[10];
#pragma unroll
      for (int k = 0; k < 10; ++k) v[k] = p[f0 + k];
#pragma unroll
      for (int k = 9; k >= 0; --k) {
        const int f = f0 + k;
        acc = __fadd_rn(f == ch1 ? t1 : (f == c2 ? t2 : v[k]), acc);
      }
    }
  }
  return acc;
}

// the look-ahead values of all candidates of a hand-off END: t.q[k] (selection) and t.q_plain[k] (TD target).  All threads.
__device__ __forceinline__ void lookahead_folds(Smem& sm, const RefMem& m, const int n, const int cell1, const int cell2) {
  RepState& s = sm.s;
  Scratch& t = sm.t;
  const int lane = lane_id(), wid = warp_id(), nthreads = blockDim.x;
  const int pos2 = pos_of(cell2);
  int k0 = 0;
  bool first_round = true;
  while (k0 < n) {  // uniform
    if (wid == 0) {
      // channels eligible at the target cell on the CURRENT grid; freeing k adds k itself when the departing call was
      // its only user within distance 2 of the target (the target is within distance 2 of the departure cell)
      uint32_t e0[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int h = 32 * j + lane;
        e0[j] = __ballot_sync(0xffffffffu, h < NCH && s.cnt2[h * PLANE + pos2] == 0);
      }
      int np = 0, k = k0;
      for (; k < n; ++k) {
        const int ch = t.cand[k];
        uint32_t m[3] = {e0[0], e0[1], e0[2]};
        if (s.cnt2[ch * PLANE + pos2] == 1) m[ch >> 5] |= 1u << (ch & 31);
        const int mk = __popc(m[0]) + __popc(m[1]) + __popc(m[2]);
        if (np + 1 + mk > LA_PAIRS) break;
        if (lane == 0) { t.pair_k[np] = (uint8_t)k; t.pair_h[np] = (uint8_t)LA_NONE; }
        int base = np + 1;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if ((m[j] >> lane) & 1u) {
            const int at = base + __popc(m[j] & ((1u << lane) - 1u));
            t.pair_k[at] = (uint8_t)k;
            t.pair_h[at] = (uint8_t)(32 * j + lane);
          }
          base += __popc(m[j]);
        }
        np = base;
      }
      if (lane == 0) { t.la_npairs = np; t.la_knext = k; }
    }
    __syncthreads();
    const int np = t.la_npairs;
    if (threadIdx.x < nthreads - 2) {
      for (int p = threadIdx.x; p < np; p += nthreads - 2) {
        const int h = t.pair_h[p];
        t.pair_q[p] = ref_fold_two(sm, m, cell1, t.cand[t.pair_k[p]], cell2, h == LA_NONE ? -1 : h);
      }
    } else if (first_round) {  // V(s) and x.wGradCorr of the current state, once
      const float v = ref_fold_plain(threadIdx.x == nthreads - 1 ? m.prod_w : m.prod_g);
      if (threadIdx.x == nthreads - 1) t.val = v;
      else t.dotg = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // per candidate: the plain value, and the maximum over its pairs (left fold with >, ascending h)
      int p = 0;
      while (p < np) {
        const int k = t.pair_k[p];
        const float plain = t.pair_q[p++];
        float best = plain;
        bool have = false;
        while (p < np && t.pair_k[p] == k) {
          const float v = t.pair_q[p++];
          if (!have) { best = v; have = true; }
          else if (v > best) best = v;
        }
        t.q_plain[k] = plain;
        t.q[k] = best;
      }
    }
    __syncthreads();
    k0 = t.la_knext;
    first_round = false;
  }
}

// ---- the run kernel's versions (k_run<ORDER_REF>): the same arithmetic, laid out for throughput ----
// One lane's dense right fold in flatten order: of the products `p` of the current state when ch < 0 (V(s) on prod_w,
// x.wGradCorr on prod_g), of the afterstate of candidate `ch` otherwise (incAfterStateFreps, Gridfuncs.hs:186-252, never
// materialised: per cell the n_elig term (f = 70) and the channel's own term (f = ch) are re-formed when the action changes
// them, the other 69 terms are the current state's products).  The code is the
// same for every lane, so ALL folds of an event -- up to 30 candidates, V(s) and x.wGradCorr -- run in lock step on ONE
// warp: an event then costs the issue slots of one fold (compare, select, FADD and a quarter of a 16-byte load per term
// against the dependent FADD chain, measured at ~4.6 cycles per term), and the fold warps of the CTAs that share an SM sit
// on different schedulers (the hardware rotates the warp slots of co-resident CTAs over the sub-partitions,
// tools/microbench/placement.cu).  Before: candidates on warp 0, the two plain folds on warp 3, two CTAs per SM -> two
// fold warps per scheduler, 6.5 issue cycles per term.
// A term is selected -- `f == ch ? re-formed : product` -- one batch AHEAD of its addition, so that the only dependence an
// addition waits for is the previous addition (written the plain way, every select lands two instructions in front of its
// FADD and the fold runs at 9 cycles per term instead of 5-6; profiles/r2_ref_order_kernel.md).
__device__ __forceinline__ float ref_sel(const float v, const float tch, const int ch, const int f) {
#ifdef DCA_REF_NOSEL  // timing probe only (wrong values): the fold without its selects
  return v;
#endif
  float u;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %2, %3;\n\tselp.f32 %0, %4, %1, p;\n\t}" : "=f"(u) : "f"(v), "r"(ch), "r"(f), "f"(tch));
  return u;
}
__device__ __forceinline__ void ref_add(float& acc, const float u) {
  asm volatile("add.rn.f32 %0, %1, %0;" : "+f"(acc) : "f"(u));
}
struct RefCellTerms {  // the two terms of a cell that candidate `ch` may change
  float t70, tch;
};
struct RefCellW {      // the two weights they are re-formed from: of n_elig and of the channel at this cell
  float w70, wch;
};
__device__ __forceinline__ RefCellW ref_cell_w(const float* w, const int cell, const int cc) {
  const int pos = cell + (cell * 37 >> 8);  // (cell / 7) * 8 + cell % 7 for cell < 49
  RefCellW r;
  r.w70 = w[NCH * PLANE + pos];
  r.wch = w[cc * PLANE + pos];
  return r;
}
__device__ __forceinline__ RefCellTerms ref_cell_terms(const RepState& s, const float* pc, const int cell, const int cc,
                                                       const unsigned long long n4, const unsigned long long n2,
                                                       const int diff, const uint8_t want, const RefCellW ww) {
  const int pos = cell + (cell * 37 >> 8);
  RefCellTerms r;
  // branch-free: both re-formed products are always computed, a select keeps the current state's where nothing changes
  const float r70 = __fmul_rn((float)((int)s.x[NCH * PLANE + pos] - diff), ww.w70);  // Gridfuncs.hs:220-252
  const float rch = __fmul_rn((float)((int)s.x[cc * PLANE + pos] + diff), ww.wch);   // Gridfuncs.hs:212-217
  const uint8_t c2 = s.cnt2[cc * PLANE + pos];
  const bool c70 = (((n2 >> cell) & 1ULL) != 0) & (c2 == want);
  const bool cch = (n4 >> cell) & 1ULL;
  r.t70 = c70 ? r70 : pc[NCH];
  r.tch = cch ? rch : pc[cc];
  return r;
}
// (`w`: wNet in the padded layout -- global memory in the run kernel, hence requested three cells ahead of its use)
__device__ __noinline__ float ref_fold_lane(const Smem& sm, const float* w, const float* p, const int ch, const bool is_end) {
  const RepState& s = sm.s;
  const bool cand = ch >= 0;
  const unsigned long long n4 = cand ? sm.t.n4mask : 0ULL, n2 = cand ? sm.t.n2mask : 0ULL;
  const int cc = cand ? ch : 0;
  const int diff = is_end ? -1 : 1;
  const uint8_t want = is_end ? 1 : 0;
  // A cell's row of PROW = 72 floats is read as six batches of twelve (three 16-byte loads each): features 60-71 (70 is
  // the n_elig term, handled apart; 71 the pad), 48-59, ..., 0-11.  u holds the selected terms of the batch added next.
  static_assert(PROW == 72 && NCH == 70, "six batches of twelve");
  float acc = 0.0f;
  const float* pc = p + (NCELL - 1) * PROW;
  RefCellW w1 = ref_cell_w(w, NCELL - 2, cc), w2 = ref_cell_w(w, NCELL - 3, cc), w3 = ref_cell_w(w, NCELL - 4, cc);
  RefCellTerms ct = ref_cell_terms(s, pc, NCELL - 1, cc, n4, n2, diff, want, ref_cell_w(w, NCELL - 1, cc));
  float u[12];
  {
    float v[12];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float4 q4 = *reinterpret_cast<const float4*>(pc + 60 + 4 * j);
      v[4 * j] = q4.x; v[4 * j + 1] = q4.y; v[4 * j + 2] = q4.z; v[4 * j + 3] = q4.w;
    }
#pragma unroll
    for (int k = 0; k < 10; ++k) u[k] = ref_sel(v[k], ct.tch, ch, 60 + k);
  }
#pragma unroll 1
  for (int cell = NCELL - 1; cell >= 0; --cell) {
    // the next cell's two terms (cell 0: its own again, unused), in flight while this cell's 71 terms are added
    const int celln = cell > 0 ? cell - 1 : 0;
    const float* pn = p + celln * PROW;
    const RefCellTerms cn = ref_cell_terms(s, pn, celln, cc, n4, n2, diff, want, w1);
    w1 = w2;
    w2 = w3;
    w3 = ref_cell_w(w, cell > 3 ? cell - 4 : 0, cc);  // (the next cell of iteration cell - 3: requested three iterations ahead)
    ref_add(acc, ct.t70);
#pragma unroll
    for (int b = 5; b >= 0; --b) {
      // batch b of this cell is added (b == 5: its ten terms 69..60) while the batch below -- the top batch of the next
      // cell after the last one -- is loaded and selected
      const float* q = b > 0 ? pc + 12 * (b - 1) : pn + 60;
      const int fq = b > 0 ? 12 * (b - 1) : 60;
      const float tq = b > 0 ? ct.tch : cn.tch;
      float v[12], un[12];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float4 q4 = *reinterpret_cast<const float4*>(q + 4 * j);
        v[4 * j] = q4.x; v[4 * j + 1] = q4.y; v[4 * j + 2] = q4.z; v[4 * j + 3] = q4.w;
      }
#pragma unroll
      for (int k = 11; k >= 0; --k) {
        if (b < 5 || k < 10) ref_add(acc, u[k]);
        if (b > 0 || k < 10) un[k] = ref_sel(v[k], tq, ch, fq + k);
      }
#pragma unroll
      for (int k = 0; k < 12; ++k) u[k] = un[k];
    }
    ct = cn;
    pc = pn;
  }
  return acc;
}
// the folds of one event without look-ahead: warp 0 = candidates 0..29 (lanes 0..29), x.wGradCorr (lane 30), V(s) (lane 31);
// warps 1.. = candidates 30.. (only the transient of the empty grid has that many)
__device__ __forceinline__ void ref_folds(Smem& sm, const RefMem& m, const int n, const bool is_end) {
  Scratch& t = sm.t;
  const int wid = warp_id(), lane = lane_id();
  const int k = wid == 0 ? (lane < 30 ? lane : -1) : 30 + 32 * (wid - 1) + lane;
  const bool plain = wid == 0 && lane >= 30;
  const bool is_cand = k >= 0 && k < n;
  if (!(plain || is_cand)) return;
  const float* p = (plain && lane == 30) ? m.prod_g : m.prod_w;
  const float v = ref_fold_lane(sm, m.w, p, is_cand ? (int)t.cand[k] : -1, is_end);
  if (is_cand) t.q[k] = v;
  else if (lane == 30) t.dotg = v;
  else t.val = v;
}
// The elementwise TDC update (Agent.hs:67-76, each operation rounded separately:
//   g = (ctd * x + cavg) - cdot * x';  w' = w - g;  wg' = wg + sg * x)  by rows of the padded layout --
// row (f, r) = eight bytes of frep, two 16-byte vectors of each weight array, no index arithmetic per element -- fused
// with the products of the NEXT event (frep is in its afterstate when the update runs and nothing touches frep or the
// weights until that event's folds).  UPDATE = false: the products alone (launch start).
template <bool UPDATE, int NT>
__device__ __forceinline__ void ref_update_rows(Smem& sm, const RefMem& m, const int tid) {
  RepState& s = sm.s;
  Scratch& t = sm.t;
  RepState* const gs = m.gs;
  const float ctd = t.ctd, cavg = t.cavg, cdot = t.cdot, sg = t.sg;
  const int act_ch = UPDATE ? t.act_ch : -1, diff = t.act_diff;
  constexpr int NIT = (NROW + NT - 1) / NT;   // rows per thread
#ifndef DCA_REF_UPD_CH
#define DCA_REF_UPD_CH 2
#endif
  constexpr int CH = DCA_REF_UPD_CH;          // rows whose weights are requested together
#pragma unroll
  for (int it0 = 0; it0 < NIT; it0 += CH) {
    // the weights may live in global memory: the rows of a chunk are requested before the first one is used
    float4 wa[CH], wb[CH], ga[CH], gb[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int row = tid + NT * (it0 + j);
      if (it0 + j < NIT && row < NROW) {
        const int f = row / ROWS, off = f * PLANE + (row - f * ROWS) * ROWPAD;
        wa[j] = *reinterpret_cast<const float4*>(&gs->w[off]);
        wb[j] = *reinterpret_cast<const float4*>(&gs->w[off + 4]);
        ga[j] = *reinterpret_cast<const float4*>(&gs->wg[off]);
        gb[j] = *reinterpret_cast<const float4*>(&gs->wg[off + 4]);
      }
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int row = tid + NT * (it0 + j);
      if (it0 + j >= NIT || row >= NROW) break;
      const int f = row / ROWS, r = row - f * ROWS;
      const int off = f * PLANE + r * ROWPAD;
      const uint2 xn = *reinterpret_cast<const uint2*>(&s.x[off]);
      uint2 xo = xn;
      if (UPDATE && act_ch >= 0) {
        if (f == act_ch) {  // x_old = x_new - diff on N4 \ {cell}
          const uint2 mk = t.n4b_upd[r];
          if (diff > 0) { xo.x -= mk.x; xo.y -= mk.y; } else { xo.x += mk.x; xo.y += mk.y; }
        }
        if (f == NCH) {     // n_elig: x_old = x_new + diff on the changed cells
          const uint2 mk = t.chgb[r];
          if (diff > 0) { xo.x += mk.x; xo.y += mk.y; } else { xo.x -= mk.x; xo.y -= mk.y; }
        }
      }
      float w[7] = {wa[j].x, wa[j].y, wa[j].z, wa[j].w, wb[j].x, wb[j].y, wb[j].z};
      float g[7] = {ga[j].x, ga[j].y, ga[j].z, ga[j].w, gb[j].x, gb[j].y, gb[j].z};
#pragma unroll
      for (int c = 0; c < COLS; ++c) {
        const float xnf = byte_f(c < 4 ? xn.x : xn.y, c & 3);
        if (UPDATE) {
          const float xof = byte_f(c < 4 ? xo.x : xo.y, c & 3);
          const float a1 = __fmul_rn(ctd, xof);
          const float a3 = __fmul_rn(cdot, xnf);
          const float gr = __fsub_rn(__fadd_rn(a1, cavg), a3);
          w[c] = __fsub_rn(w[c], gr);
          g[c] = __fadd_rn(g[c], __fmul_rn(sg, xof));
        }
        const int i = (r * COLS + c) * PROW + f;  // flatten order, rows of PROW floats
        m.prod_w[i] = __fmul_rn(xnf, w[c]);
        m.prod_g[i] = __fmul_rn(xnf, g[c]);
      }
      if (UPDATE) {  // (the pad column goes back as it came)
        *reinterpret_cast<float4*>(&gs->w[off]) = make_float4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<float4*>(&gs->w[off + 4]) = make_float4(w[4], w[5], w[6], wb[j].w);
        *reinterpret_cast<float4*>(&gs->wg[off]) = make_float4(g[0], g[1], g[2], g[3]);
        *reinterpret_cast<float4*>(&gs->wg[off + 4]) = make_float4(g[4], g[5], g[6], gb[j].w);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// hashes for parity traces (order-independent u64 sums); all threads
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ void state_hashes(Smem& sm) {
  RepState& s = sm.s;
  unsigned long long hg = 0, hf = 0;
  for (int i = threadIdx.x; i < NFEAT * NCELL; i += blockDim.x) {
    const int cell = i / NFEAT, f = i - cell * NFEAT;
    hf += (unsigned long long)s.x[f * PLANE + pos_of(cell)] * mix64(0x10000ULL + (unsigned long long)i);
    if (f < NCH && ((s.grid[cell][f >> 5] >> (f & 31)) & 1u)) hg += mix64((unsigned long long)(cell * NCH + f));
  }
  for (int off = 16; off > 0; off >>= 1) {
    hg += __shfl_xor_sync(0xffffffffu, hg, off);
    hf += __shfl_xor_sync(0xffffffffu, hf, off);
  }
  if (lane_id() == 0) {
    atomicAdd(&sm.t.hash_grid, hg);
    atomicAdd(&sm.t.hash_frep, hf);
  }
}

// violatesReuseConstraint (Gridfuncs.hs:94-101) and max(frep) > 70
// (SimRunner.hs:150,158-159) on the staged state; warp 0.  Returns a status or 0.
__device__ __forceinline__ int verify_state(const Smem& sm, bool rc, bool nb) {
  const RepState& s = sm.s;
  const int lane = lane_id();
  bool bad_rc = false, bad_nb = false;
  if (rc) {
    for (int cell = lane; cell < NCELL; cell += 32) {
      unsigned long long nbm = c_tab.n2incl[cell] & ~(1ULL << cell);
      uint32_t o0 = 0, o1 = 0, o2 = 0;
      while (nbm) {
        int n = __ffsll((long long)nbm) - 1;
        nbm &= nbm - 1;
        o0 |= s.grid[n][0]; o1 |= s.grid[n][1]; o2 |= s.grid[n][2];
      }
      bad_rc |= ((o0 & s.grid[cell][0]) | (o1 & s.grid[cell][1]) | (o2 & s.grid[cell][2])) != 0;
    }
  }
  if (nb) {
    for (int i = lane; i < WPAD; i += 32) bad_nb |= s.x[i] > 70;
  }
  if (__any_sync(0xffffffffu, bad_rc)) return ST_REUSE_VIOLATED;
  if (__any_sync(0xffffffffu, bad_nb)) return ST_NELIG_OVERFLOW;
  return 0;
}

// ---------------------------------------------------------------------------
// state staging
// ---------------------------------------------------------------------------
__device__ __forceinline__ void stage_in(Smem& sm, const RepState* g) {
  const uint4* src = reinterpret_cast<const uint4*>(g);
  uint4* dst = reinterpret_cast<uint4*>(&sm.s);
  for (int i = threadIdx.x; i < (int)(sizeof(RepState) / 16); i += blockDim.x) dst[i] = src[i];
}
__device__ __forceinline__ void stage_out(const Smem& sm, RepState* g) {
  const uint4* src = reinterpret_cast<const uint4*>(&sm.s);
  uint4* dst = reinterpret_cast<uint4*>(g);
  for (int i = threadIdx.x; i < (int)(sizeof(RepState) / 16); i += blockDim.x) dst[i] = src[i];
}

// the run kernel's staging: everything but wNet and wGradCorr, which stay in global memory (their shared-memory bytes hold
// the products)
static_assert(offsetof(RepState, w) == 0 && offsetof(RepState, wg) == sizeof(((RepState*)0)->w) && offsetof(RepState, x) % 16 == 0,
              "RepState starts with w and wg, a 16-byte aligned range");
__device__ __forceinline__ void stage_in_run(Smem& sm, const RepState* g) {
  const uint4* src = reinterpret_cast<const uint4*>(g);
  uint4* dst = reinterpret_cast<uint4*>(&sm.s);
  constexpr int B = (int)(offsetof(RepState, x) / 16), N = (int)(sizeof(RepState) / 16);
  for (int i = B + threadIdx.x; i < N; i += blockDim.x) dst[i] = src[i];
}
__device__ __forceinline__ void stage_out_run(const Smem& sm, RepState* g) {
  const uint4* src = reinterpret_cast<const uint4*>(&sm.s);
  uint4* dst = reinterpret_cast<uint4*>(g);
  constexpr int B = (int)(offsetof(RepState, x) / 16), N = (int)(sizeof(RepState) / 16);
  for (int i = B + threadIdx.x; i < N; i += blockDim.x) dst[i] = src[i];
}

__device__ __forceinline__ Rng make_rng(const RunArgs& a, const RepState& s, int rep) {
  Rng rng;
  rng.mode = a.rng_mode;
  rng.k0 = a.key0;
  rng.k1 = a.key1;
  rng.stream = s.stream;
  rng.ndraws = s.ndraws;
  rng.words = nullptr;
  rng.neglog = nullptr;
  rng.tape_len = 0;
  if (a.rng_mode == RNG_TAPE && a.tape_off) {
    unsigned long long o = a.tape_off[rep];
    rng.words = a.tape_words + o;
    rng.neglog = a.tape_neglog + o;
    rng.tape_len = a.tape_off[rep + 1] - o;
  }
  return rng;
}

// ---------------------------------------------------------------------------
// One event of one replica (runStep, SimRunner.hs:60-97), both orders.
// Called by every thread of the CTA.  `fixed_*`: step-wise API overrides.
// ---------------------------------------------------------------------------
template <int ORDER, bool TRACE>
__device__ __forceinline__ void run_event(Smem& sm, const RunArgs& a, QRegs& q, Rng& rng,
                                          int rep, const RefMem& m) {
  DCA_RCLK_START
  RepState& s = sm.s;
  Scratch& t = sm.t;
  const int wid = warp_id(), lane = lane_id();
  const int nthreads = blockDim.x;

  // ---- accFn1 = getAction (Simulator.hs:154-172) ----
  if (wid == 0) {
    const int cell = s.ev_cell;
    const bool is_end = s.ev_type == EV_END;
    const int n = build_candidates(sm, cell, is_end);
    if (lane == 0) {
      t.ncand = n;
      t.ev_cell = cell;
      t.ev_is_end = is_end;
      t.ev_type_prev = s.ev_type;
      t.ev_ch_prev = s.ev_ch;
      t.ev_time_prev = s.ev_time;
      t.hash_grid = 0;
      t.hash_frep = 0;
    }
    __syncwarp();
  }
  {
    __syncthreads();
    DCA_RCLK(0)
    const int n = t.ncand;
    const bool is_end = t.ev_is_end;
    // (the products of the current state are in place: formed at launch start and by every update since)
    const bool la = a.lookahead && is_end && n > 0 && s.ev_hoff != HOFF_NONE;  // uniform: the departure half of a hand-off
    if (la) {
      lookahead_folds(sm, m, n, t.ev_cell, s.ev_hoff);
    } else {
      ref_folds(sm, m, n, is_end);
      for (int k = threadIdx.x; k < n; k += nthreads) t.q_plain[k] = 0.0f;  // (unused without look-ahead)
    }
    if (threadIdx.x == 0) t.flag = la ? 1 : 0;
    __syncthreads();
    DCA_RCLK(1)
  }
  if (wid == 0) {
    const int n = t.ncand, cell = t.ev_cell;
    const bool is_end = t.ev_is_end;
    const float val = t.val, dotg = t.dotg;
    int act = -1;
    float next_val = val;  // no action: nextFrep = frep (Simulator.hs:170-171)
    if (n > 0) {
      const int k = argmax_last(t.q, n);  // selectAction, Simulator.hs:202
      act = t.cand[k];
      next_val = t.flag ? t.q_plain[k] : t.q[k];  // (look-ahead only steers the selection: nextFrep is the real afterstate)
      const unsigned long long chg = elig_change_mask(s, act, is_end, t.n2mask);
      apply_action(sm, cell, is_end, act, chg);  // gridStep, Simulator.hs:137-144
    }
    // ---- environmentStep (Simulator.hs:101-134) ----
    environment_step(s, q, rng, act);
    // ---- backward scalars (Agent.hs:62-67,75,78) ----
    const float avgR = s.avg_reward;
    const float reward = (float)s.ncalls;  // boolSum nextGrid
    const float td = __fsub_rn(__fadd_rn(__fsub_rn(reward, avgR), next_val), val);
    const float c = __fmul_rn(-2.0f, s.alpha_net);
    if (lane == 0) {
      t.ctd = __fmul_rn(c, td);
      t.cavg = __fmul_rn(c, avgR);
      t.cdot = __fmul_rn(c, dotg);
      t.sg = __fmul_rn(s.alpha_grad, __fsub_rn(td, dotg));
      t.act_ch = act;
      t.act_diff = is_end ? -1 : 1;
      t.loss = td;
      s.avg_reward = __fadd_rn(avgR, __fmul_rn(s.alpha_avg, td));
      s.iter += 1;  // Simulator.hs:131
      // stop rules of runPeriod (SimRunner.hs:150-169), in the reference's order
      int stop = ST_RUNNING;
      if (td != td) stop = ST_NAN_LOSS;
      else {
        s.loss_sum = __fadd_rn(s.loss_sum, td);
        s.loss_n += 1;
      }
      if (stop == ST_RUNNING && q.overflow) stop = ST_QUEUE_OVERFLOW;
      t.stop = stop;
    }
  }
  if (threadIdx.x < 8) t.n4b_upd[threadIdx.x] = t.n4b[threadIdx.x];
  __syncthreads();
  DCA_RCLK(2)
  // ---- accFn2's dense part: weight / gradient-correction update (and the next event's products) ----
  ref_update_rows<true, 128>(sm, m, (int)threadIdx.x);
  if (TRACE) state_hashes(sm);
  __syncthreads();
  DCA_RCLK(3)
  if (wid == 0) {
    int stop = t.stop;
    if (stop == ST_RUNNING && (a.verify_rc || a.verify_nb)) {
      int v = verify_state(sm, a.verify_rc, a.verify_nb);
      if (v) stop = v;
    }
    if (lane == 0) {
      const float loss = t.loss;
      if (stop == ST_RUNNING && s.min_loss > 0.0f && s.iter > 50 && fabsf(loss) < s.min_loss)
        stop = ST_ZERO_LOSS;
      if (stop == ST_RUNNING && s.iter == s.n_events) stop = ST_SUCCESS;
      if (stop == ST_RUNNING && a.rng_mode == RNG_TAPE && rng.ndraws + 8 > rng.tape_len)
        stop = ST_TAPE_EXHAUSTED;
      s.status = stop;
      if (TRACE && a.trace && s.iter - 1 < a.trace_cap) {
        TraceRec& tr = a.trace[(size_t)rep * a.trace_cap + (s.iter - 1)];
        tr.time = t.ev_time_prev;
        tr.loss = loss;
        tr.avg_reward = s.avg_reward;
        tr.reward = s.ncalls;
        tr.etype = (uint8_t)t.ev_type_prev;
        tr.cell = (uint8_t)t.ev_cell;
        tr.ch = (int8_t)t.act_ch;
        tr.ncand = (uint8_t)t.ncand;
        tr.end_ch = (int8_t)(t.ev_is_end ? t.ev_ch_prev : -1);
        for (int i = 0; i < 7; ++i) tr.pad[i] = 0;
        tr.grid_hash = t.hash_grid;
        tr.frep_hash = t.hash_frep;
      }
    }
  }
  __syncthreads();
  DCA_RCLK(4)
}

// ---------------------------------------------------------------------------
// The event loop of the run kernel without the parity trace: the same steps as run_event, arranged so that the
// event-sequential work of an event runs NEXT TO its folds and its dense update instead of between them.  While the
// folds run, warp 3 does the part of environmentStep that does not depend on WHICH channel is chosen (env_ahead); once
// the decision and the TD scalars are published, warps 1-3 run the update (and form the next event's products) while
// warp 0 finishes environmentStep (env_post), applies the stop rules and builds the next event's candidate list -- none
// of which reads the weights.  Three CTA barriers per event instead of five.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void ref_prepare_event(Smem& sm) {  // warp 0: candidates and masks of the current event
  RepState& s = sm.s;
  Scratch& t = sm.t;
  const int cell = s.ev_cell;
  const bool is_end = s.ev_type == EV_END;
  const int n = build_candidates(sm, cell, is_end);
  if (lane_id() == 0) {
    t.ncand = n;
    t.ev_cell = cell;
    t.ev_is_end = is_end;
    t.cur_type = s.ev_type;
    t.cur_ch = s.ev_ch;
    t.cur_hoff = s.ev_hoff;
  }
  __syncwarp();
}
// environmentStep (Simulator.hs:101-134) in two parts, as in the FAST kernel: what an event does to the queue depends on
// WHETHER a channel is chosen (`accepted` = it has a candidate), not on which.  env_ahead -- warp 3, while the folds of
// the event run -- books the statistics, draws the random numbers, pushes the new events (an END event without its
// channel: key channel CH_NONE_YET) and pops the next event into s.ev_*; env_post -- warp 0, after the decision -- writes
// the chosen channel into that END event (or into s.ev_ch when it is the event just popped) and performs the
// reassignment of an END (Simulator.hs:129-130) on the queue as the pop left it.  Pop order depends on (time, id) only,
// so the result is the one of environment_step.
constexpr int CH_NONE_YET = 127;
__device__ __forceinline__ void env_ahead(Smem& sm, QRegs& q, Rng& rng, const bool accepted) {
  RepState& s = sm.s;
  Scratch& t = sm.t;
  const int lane = lane_id();
  const int type = t.cur_type, cell = t.ev_cell, ev_hoff = t.cur_hoff;
  const double time = s.ev_time;
  int pend = -1;
  if (lane == 0) {  // Stats hooks, Stats.hs:61-81 as called from Simulator.hs:105-114
    if (type == EV_NEW) {
      s.stats[6]++; s.stats[0]++;
      if (accepted) s.stats[1]++;
      else { s.stats[2]++; s.stats[7]++; }
    } else if (type == EV_HOFF) {
      s.stats[3]++;
      if (!accepted) { s.stats[2]++; s.stats[7]++; }  // Simulator.hs:113
    } else {
      s.stats[4]++;
    }
  }
  if (type == EV_NEW) {
    const double dt = rng.exponential(s.mean_new);  // generateNewEvent, EventGen.hs:75-85
    q_push_new(s, q, cell, __dadd_rn(time, dt));
    if (accepted) {
      const double p = rng.uniform();  // Simulator.hs:121, drawn even when hoffProb == 0
      pend = QNEW + q.n_end < (unsigned)QCAP ? (int)(QNEW + q.n_end) : -1;
      if (p < s.hoff_prob) {           // generateHoffNewEvent, EventGen.hs:94-109
        const double d2 = rng.exponential(s.call_dur);
        const int mm = c_tab.n2cnt[cell];
        const int to = c_tab.n2list[cell][rng.uniform_int(mm)];
        q_push_end(s, q, cell, CH_NONE_YET, to, __dadd_rn(time, d2));
        q.next_id++;  // the HOFF event's id (pushed right after its END, never stored)
      } else {                         // generateEndEvent, EventGen.hs:111-112
        const double d2 = rng.exponential(s.call_dur);
        q_push_end(s, q, cell, CH_NONE_YET, HOFF_NONE, __dadd_rn(time, d2));
      }
    }
  } else if (type == EV_HOFF) {
    if (accepted) {                    // generateHoffEndEvent, EventGen.hs:114-116
      const double d2 = rng.exponential(s.call_dur_hoff);
      pend = QNEW + q.n_end < (unsigned)QCAP ? (int)(QNEW + q.n_end) : -1;
      q_push_end(s, q, cell, CH_NONE_YET, HOFF_NONE, __dadd_rn(time, d2));
    }
  }
  __syncwarp();
  // the next event (Simulator.hs:132-134)
  if (type == EV_END && ev_hoff != HOFF_NONE) {
    if (lane == 0) {  // the HOFF that follows a hand-off END: same time, next id
      s.ev_type = EV_HOFF;
      s.ev_cell = ev_hoff;
      s.ev_ch = -1;
      s.ev_hoff = HOFF_NONE;
    }
  } else {
    q_pop(s, q, &pend);
  }
  if (lane == 0) {
    t.pend_slot = pend;
    t.q_n_end = q.n_end;
    t.q_overflow = q.overflow ? 1 : 0;
    t.rng_ndraws = rng.ndraws;
  }
  __syncwarp();
}
__device__ __forceinline__ void env_post(Smem& sm, const int act) {
  RepState& s = sm.s;
  Scratch& t = sm.t;
  const int lane = lane_id();
  const int type = t.cur_type, cell = t.ev_cell, ev_ch = t.cur_ch;
  if (lane == 0) {
    if (t.pend_slot >= 0) s.q_key[t.pend_slot] = (uint16_t)((cell << 7) | act);
    else if (s.ev_type == EV_END && s.ev_ch == CH_NONE_YET) s.ev_ch = act;  // the call accepted just now ends before anything else happens
  }
  __syncwarp();
  if (type == EV_END && act >= 0 && act != ev_ch) {
    // END toCh, Simulator.hs:129-130: reassign cell act -> ev_ch (EventGen.hs:62-72): the call that used `act` continues
    // on ev_ch, so its END event is re-keyed -- in the queue, or in s.ev_* when it is the event just popped
    const bool popped = s.ev_type == EV_END && s.ev_cell == cell && s.ev_ch == act;  // uniform
    __syncwarp();
    if (popped) {
      if (lane == 0) s.ev_ch = ev_ch;
    } else {
      QRegs qq{t.q_n_end, 0u, false};
      q_reassign(s, qq, cell, act, ev_ch);
    }
  }
  __syncwarp();
}
__device__ __forceinline__ void run_events_overlapped(Smem& sm, const RunArgs& a, QRegs& q, Rng& rng, const RefMem& m) {
  RepState& s = sm.s;
  Scratch& t = sm.t;
  const int wid = warp_id(), lane = lane_id();
  if (wid == 0) ref_prepare_event(sm);
  DCA_RCLK_START
  for (long long it = 0;; ++it) {
    __syncthreads();  // the event's candidates, the products of its state and the status word are published
    DCA_RCLK(3)
    if (it >= a.max_events || s.status != ST_RUNNING) break;  // uniform
    const int n = t.ncand;
    const bool is_end = t.ev_is_end;
    const int cur_hoff = t.cur_hoff;
    const bool la = a.lookahead && is_end && n > 0 && cur_hoff != HOFF_NONE;  // uniform: the departure half of a hand-off
    if (wid == 3) env_ahead(sm, q, rng, n > 0);  // (warp 3 holds no fold: at most 70 candidates, 30 + 32 + 32 lanes on warps 0-2)
    if (la) lookahead_folds(sm, m, n, t.ev_cell, cur_hoff);
    else ref_folds(sm, m, n, is_end);
    __syncthreads();
    DCA_RCLK(1)
    int act = -1;
    if (wid == 0) {
      const int cell = t.ev_cell;
      const float val = t.val, dotg = t.dotg;
      float next_val = val;  // no action: nextFrep = frep (Simulator.hs:170-171)
      if (n > 0) {
        const int k = argmax_last(t.q, n);  // selectAction, Simulator.hs:202
        act = t.cand[k];
        next_val = la ? t.q_plain[k] : t.q[k];  // (look-ahead only steers the selection: nextFrep is the real afterstate)
        const unsigned long long chg = elig_change_mask(s, act, is_end, t.n2mask);
        apply_action(sm, cell, is_end, act, chg);  // gridStep, Simulator.hs:137-144
      }
      // backward scalars (Agent.hs:62-67,75,78)
      const float avgR = s.avg_reward;
      const float reward = (float)s.ncalls;  // boolSum nextGrid
      const float td = __fsub_rn(__fadd_rn(__fsub_rn(reward, avgR), next_val), val);
      const float c = __fmul_rn(-2.0f, s.alpha_net);
      if (lane < 8) t.n4b_upd[lane] = t.n4b[lane];
      if (lane == 0) {
        t.ctd = __fmul_rn(c, td);
        t.cavg = __fmul_rn(c, avgR);
        t.cdot = __fmul_rn(c, dotg);
        t.sg = __fmul_rn(s.alpha_grad, __fsub_rn(td, dotg));
        t.act_ch = act;
        t.act_diff = is_end ? -1 : 1;
        t.loss = td;
        s.avg_reward = __fadd_rn(avgR, __fmul_rn(s.alpha_avg, td));
      }
    }
    __syncthreads();  // the decision and the scalars of the update are published
    DCA_RCLK(2)
    if (wid == 0) {
      env_post(sm, act);
      int stop = ST_RUNNING;
      if (lane == 0) {
        const float td = t.loss;
        s.iter += 1;  // Simulator.hs:131
        // stop rules of runPeriod (SimRunner.hs:150-169), in the reference's order
        if (td != td) stop = ST_NAN_LOSS;
        else {
          s.loss_sum = __fadd_rn(s.loss_sum, td);
          s.loss_n += 1;
        }
        if (stop == ST_RUNNING && t.q_overflow) stop = ST_QUEUE_OVERFLOW;
      }
      stop = __shfl_sync(0xffffffffu, stop, 0);
      if (stop == ST_RUNNING && (a.verify_rc || a.verify_nb)) {
        const int v = verify_state(sm, a.verify_rc, a.verify_nb);
        if (v) stop = v;
      }
      if (lane == 0) {
        if (stop == ST_RUNNING && s.min_loss > 0.0f && s.iter > 50 && fabsf(t.loss) < s.min_loss) stop = ST_ZERO_LOSS;
        if (stop == ST_RUNNING && s.iter == s.n_events) stop = ST_SUCCESS;
        if (stop == ST_RUNNING && a.rng_mode == RNG_TAPE && t.rng_ndraws + 8 > rng.tape_len) stop = ST_TAPE_EXHAUSTED;
        s.status = stop;
      }
      stop = __shfl_sync(0xffffffffu, stop, 0);
      if (stop == ST_RUNNING && it + 1 < a.max_events) ref_prepare_event(sm);  // (does not touch what the update reads)
    }
    // accFn2's dense part: weight / gradient-correction update (and the next event's products); warp 0 joins with its
    // share of the rows once its tail is done
    ref_update_rows<true, 128>(sm, m, (int)threadIdx.x);
  }
}

// ---------------------------------------------------------------------------
// The fused run kernel: dca_run.  grid = n_replicas CTAs.
// ---------------------------------------------------------------------------
template <int ORDER, bool TRACE>
__global__ void __launch_bounds__(128, 4) k_run(RunArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int rep = blockIdx.x;
  if (rep >= a.n_replicas) return;
  RepState* g = a.states + rep;
  stage_in_run(sm, g);
  __syncthreads();
  if (!sm.s.initialized || sm.s.status != ST_RUNNING) return;  // uniform
  QRegs q{sm.s.n_end, sm.s.next_id, false};
  Rng rng = make_rng(a, sm.s, rep);
  // wNet and wGradCorr themselves stay in global memory; their shared-memory bytes hold the products (see Scratch::prod_w)
  const RefMem m{g->w, g, sm.s.w, sm.s.wg};
  ref_update_rows<false, 128>(sm, m, (int)threadIdx.x);  // the products of the staged state (the barrier in front of the first folds covers them)
  if (TRACE) {
    for (long long it = 0; it < a.max_events; ++it) {
      run_event<ORDER, TRACE>(sm, a, q, rng, rep, m);
      if (sm.s.status != ST_RUNNING) break;  // uniform after the trailing barrier
    }
  } else {
    run_events_overlapped(sm, a, q, rng, m);
  }
  if (threadIdx.x == (TRACE ? 0u : 96u)) {  // (the queue counters and the random source live on warp 0 / on warp 3: env_ahead)
    sm.s.n_end = q.n_end;
    sm.s.next_id = q.next_id;
    sm.s.ndraws = rng.ndraws;
  }
  __syncthreads();
  stage_out_run(sm, g);
}

// ---------------------------------------------------------------------------
// mkSimState (Simulator.hs:88-97) on the device: one warp per replica.
// ---------------------------------------------------------------------------
struct InitArgs {
  RepState* states;
  int n_replicas;
  int rng_mode;
  uint32_t key0, key1;
  unsigned long long first_replica;
  long long n_events;
  float min_loss;
  const double *call_dur, *call_dur_hoff, *mean_new, *hoff_prob;  // per replica
  const float *alpha_net, *alpha_avg, *alpha_grad;
  const unsigned long long* tape_words;
  const double* tape_neglog;
  const unsigned long long* tape_off;
  int only_replica;  // >= 0: initialise just this one
};

__global__ void k_init(InitArgs a) {
  const int rep = (a.only_replica >= 0) ? a.only_replica : (int)(blockIdx.x * (blockDim.x >> 5) + warp_id());
  if (rep >= a.n_replicas) return;
  if (a.only_replica >= 0 && (blockIdx.x != 0 || warp_id() != 0)) return;
  RepState& s = a.states[rep];
  const int lane = lane_id();
  // zero the whole blob (grid, agent, queue, stats)
  uint4* p = reinterpret_cast<uint4*>(&s);
  for (int i = lane; i < (int)(sizeof(RepState) / 16); i += 32) p[i] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  // frep = featureRep emptyGrid: in-use counts 0, n_elig 70 everywhere
  for (int cell = lane; cell < NCELL; cell += 32) s.x[NCH * PLANE + pos_of(cell)] = NCH;
  for (int i = lane; i < QCAP; i += 32) {
    s.q_time[i] = CUDART_INF;
    s.q_key[i] = (uint16_t)(((i < QNEW ? i : 0) << 7) | 127);
    s.q_hoff[i] = HOFF_NONE;
  }
  __syncwarp();
  RunArgs ra{};
  ra.rng_mode = a.rng_mode;
  ra.key0 = a.key0;
  ra.key1 = a.key1;
  ra.tape_words = a.tape_words;
  ra.tape_neglog = a.tape_neglog;
  ra.tape_off = a.tape_off;
  if (lane == 0) {
    s.stream = a.first_replica + (unsigned long long)rep;
    s.mean_new = a.mean_new[rep];  // 1/(callRate/60), EventGen.hs:81-82, computed on the host
    s.call_dur = a.call_dur[rep];
    s.call_dur_hoff = a.call_dur_hoff[rep];
    s.hoff_prob = a.hoff_prob[rep];
    s.alpha_net = a.alpha_net[rep];
    s.alpha_avg = a.alpha_avg[rep];
    s.alpha_grad = a.alpha_grad[rep];
    s.min_loss = a.min_loss;
    s.n_events = a.n_events;
    s.status = ST_RUNNING;
  }
  __syncwarp();
  const bool have_rng = a.rng_mode == RNG_PHILOX ||
                        (a.tape_off && a.tape_off[rep + 1] - a.tape_off[rep] >= (unsigned long long)NCELL);
  if (!have_rng) {  // tape not set yet: stays uninitialised
    if (lane == 0) { s.initialized = 0; s.status = ST_UNINITIALIZED; }
    return;
  }
  Rng rng = make_rng(ra, s, rep);
  QRegs q{0u, 0u, false};
  // mkEventGen (EventGen.hs:54-58): one NEW per cell, row-major, at 0 + Exp(mean_new)
  for (int cell = 0; cell < NCELL; ++cell) {
    double dt = rng.exponential(s.mean_new);
    q_push_new(s, q, cell, __dadd_rn(0.0, dt));
  }
  __syncwarp();
  q_pop(s, q);  // Simulator.hs:95
  if (lane == 0) {
    s.n_end = q.n_end;
    s.next_id = q.next_id;
    s.ndraws = rng.ndraws;
    s.initialized = 1;
  }
}

// ---------------------------------------------------------------------------
// featureRep from scratch (Gridfuncs.hs:153-170) + cnt2, from the grid masks of
// a replica blob.  One CTA per replica.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void feature_rep_from_grid(RepState& s) {
  for (int i = threadIdx.x; i < NCH * NCELL; i += blockDim.x) {
    const int ch = i / NCELL, cell = i - ch * NCELL;
    const int wsel = ch >> 5, bit = ch & 31;
    int in4 = 0, in2 = 0;
    unsigned long long m4 = c_tab.n4excl[cell], m2 = c_tab.n2incl[cell];
    while (m4) {
      int n = __ffsll((long long)m4) - 1;
      m4 &= m4 - 1;
      in4 += (s.grid[n][wsel] >> bit) & 1u;
    }
    while (m2) {
      int n = __ffsll((long long)m2) - 1;
      m2 &= m2 - 1;
      in2 += (s.grid[n][wsel] >> bit) & 1u;
    }
    s.x[ch * PLANE + pos_of(cell)] = (uint8_t)in4;
    s.cnt2[ch * PLANE + pos_of(cell)] = (uint8_t)in2;
  }
  __syncthreads();
  for (int cell = threadIdx.x; cell < NCELL; cell += blockDim.x) {
    int ne = 0, nc = 0;
    for (int ch = 0; ch < NCH; ++ch) ne += s.cnt2[ch * PLANE + pos_of(cell)] == 0;
    nc = __popc(s.grid[cell][0]) + __popc(s.grid[cell][1]) + __popc(s.grid[cell][2]);
    s.x[NCH * PLANE + pos_of(cell)] = (uint8_t)ne;
    atomicAdd(&s.ncalls, nc);
  }
  __syncthreads();
}

__global__ void k_feature_rep(RepState* states, int rep) {
  RepState& s = states[rep];
  if (threadIdx.x == 0) s.ncalls = 0;
  __syncthreads();
  feature_rep_from_grid(s);
}

// ---------------------------------------------------------------------------
// Step-wise operators on one replica: accFn1 / accFn2 (SimRunner.hs:61-62).
// ---------------------------------------------------------------------------
constexpr int STEP_BAD_CHANNEL = -3;  // StepOut.ch: gridStepTrain was given a channel getAction could not have returned
struct StepOut {
  int ch;
  int loss_is_nan;
  float loss;
  int pad;
  long long reward;
};

// mode 0: getAction -> out->ch.  mode 1: gridStepTrain with the given ch.
template <int ORDER>
__global__ void __launch_bounds__(128) k_step(RepState* states, int rep, int mode, int cell,
                                              int is_end_i, int ch_in, float alpha_net,
                                              float alpha_avg, float alpha_grad, StepOut* out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  RepState& s = sm.s;
  Scratch& t = sm.t;
  const bool is_end = is_end_i != 0;
  const int wid = warp_id(), lane = lane_id();
  const RefMem m{sm.s.w, &sm.s, t.prod_w, t.prod_g};  // (the step kernel keeps the whole blob and its own product arrays in shared memory)
  stage_in(sm, states + rep);
  __syncthreads();
  if (wid == 0) {
    const int n = build_candidates(sm, cell, is_end);
    bool found = mode == 0 || ch_in < 0;  // the channel must be one getAction could have returned at this cell
    for (int k = lane; k < n; k += 32) found |= t.cand[k] == ch_in;
    found = __any_sync(0xffffffffu, found);
    __syncwarp();
    if (lane == 0) {
      t.ncand = n;
      t.flag = found ? 0 : 1;
      if (mode == 1) {  // evaluate exactly the channel the caller executes
        t.ncand = ch_in >= 0 ? 1 : 0;
        t.cand[0] = (uint8_t)(ch_in >= 0 ? ch_in : 0);
      }
    }
  }
  __syncthreads();
  if (t.flag) {  // (uniform) nothing is changed
    if (threadIdx.x == 0) out->ch = STEP_BAD_CHANNEL;
    return;
  }
  const int n = t.ncand;
  {
    ref_update_rows<false, 128>(sm, m, (int)threadIdx.x);
    __syncthreads();
    ref_folds(sm, m, n, is_end);
    __syncthreads();
  }
  if (wid == 0) {
    const float val = t.val, dotg = t.dotg;
    int act = -1;
    float next_val = val;
    if (n > 0) {
      const int k = argmax_last(t.q, n);
      act = t.cand[k];
      next_val = t.q[k];
    }
    if (mode == 0) {
      if (lane == 0) out->ch = act;
      t.act_ch = -2;  // nothing to apply
    } else {
      if (act >= 0) {
        const unsigned long long chg = elig_change_mask(s, act, is_end, t.n2mask);
        apply_action(sm, cell, is_end, act, chg);
      }
      const float avgR = s.avg_reward;
      const float reward = (float)s.ncalls;
      const float td = __fsub_rn(__fadd_rn(__fsub_rn(reward, avgR), next_val), val);
      const float c = __fmul_rn(-2.0f, alpha_net);
      if (lane == 0) {
        t.ctd = __fmul_rn(c, td);
        t.cavg = __fmul_rn(c, avgR);
        t.cdot = __fmul_rn(c, dotg);
        t.sg = __fmul_rn(alpha_grad, __fsub_rn(td, dotg));
        t.act_ch = act;
        t.act_diff = is_end ? -1 : 1;
        s.avg_reward = __fadd_rn(avgR, __fmul_rn(alpha_avg, td));
        out->ch = act;
        out->loss = td;
        out->loss_is_nan = (td != td) ? 1 : 0;
        out->reward = s.ncalls;
      }
    }
  }
  if (threadIdx.x < 8) t.n4b_upd[threadIdx.x] = t.n4b[threadIdx.x];
  __syncthreads();
  if (mode == 1) {
    ref_update_rows<true, 128>(sm, m, (int)threadIdx.x);
    __syncthreads();
    stage_out(sm, states + rep);
  }
}

// ---------------------------------------------------------------------------
// Stateless Gridfuncs operators on explicit arrays (one CTA)
// ---------------------------------------------------------------------------
// grid bytes [49][70] -> masks in a scratch RepState
__global__ void k_gf_load_grid(RepState* st, const uint8_t* grid) {
  RepState& s = *st;
  for (int i = threadIdx.x; i < NCELL * 4; i += blockDim.x) (&s.grid[0][0])[i] = 0;
  for (int i = threadIdx.x; i < WPAD; i += blockDim.x) { s.x[i] = 0; s.cnt2[i] = 0; }
  if (threadIdx.x == 0) s.ncalls = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < NCELL * NCH; i += blockDim.x) {
    int cell = i / NCH, ch = i - cell * NCH;
    if (grid[i]) atomicOr(&s.grid[cell][ch >> 5], 1u << (ch & 31));
  }
  __syncthreads();
  feature_rep_from_grid(s);
}
// frep (reference flatten order, int64) out of / into the blob
__global__ void k_gf_store_frep(const RepState* st, long long* frep) {
  for (int i = threadIdx.x; i < FREP; i += blockDim.x) {
    int cell = i / NFEAT, f = i - cell * NFEAT;
    frep[i] = st->x[f * PLANE + pos_of(cell)];
  }
}
__global__ void k_gf_load_frep(RepState* st, const long long* frep) {
  for (int i = threadIdx.x; i < FREP; i += blockDim.x) {
    int cell = i / NFEAT, f = i - cell * NFEAT;
    st->x[f * PLANE + pos_of(cell)] = (uint8_t)frep[i];
  }
}
// eligible (is_end = 0) / in-use (is_end = 1) channels, ascending
__global__ void k_gf_chs(const RepState* st, int cell, int is_end, long long* chs, int* n_out) {
  __shared__ int n;
  if (threadIdx.x == 0) {
    n = 0;
    for (int ch = 0; ch < NCH; ++ch) {
      bool on = is_end ? ((st->grid[cell][ch >> 5] >> (ch & 31)) & 1u)
                       : (st->cnt2[ch * PLANE + pos_of(cell)] == 0);
      if (on) chs[n++] = ch;
    }
    *n_out = n;
  }
}
// incAfterStateFreps (Gridfuncs.hs:186-252) in the cnt2 formulation
__global__ void k_gf_inc_afterstate_freps(const RepState* st, int cell, int is_end,
                                          const long long* chs, int n, const long long* frep, long long* out) {
  const unsigned long long n4 = c_tab.n4excl[cell], n2 = c_tab.n2incl[cell];
  const int diff = is_end ? -1 : 1;
  const uint8_t want = is_end ? 1 : 0;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < (long long)n * FREP;
       idx += (long long)gridDim.x * blockDim.x) {
    int k = (int)(idx / FREP), i = (int)(idx - (long long)k * FREP);
    int c2 = i / NFEAT, f = i - c2 * NFEAT;
    int ch = (int)chs[k];
    long long v = frep[i];  // the caller's frep (it need not be featureRep grid), Gridfuncs.hs:206
    if (f == ch && ((n4 >> c2) & 1ULL)) v += diff;
    if (f == NCH && ((n2 >> c2) & 1ULL) && st->cnt2[ch * PLANE + pos_of(c2)] == want) v -= diff;
    out[idx] = v;
  }
}
__global__ void k_gf_violates(const RepState* st, int* out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  stage_in(sm, st);
  __syncthreads();
  if (warp_id() == 0) {
    int v = verify_state(sm, true, false);
    if (lane_id() == 0) *out = (v == ST_REUSE_VIOLATED) ? 1 : 0;
  }
}

// ---------------------------------------------------------------------------
// read-back helpers
// ---------------------------------------------------------------------------
struct HostView {  // unpacked replica state in the reference's layouts
  uint8_t grid[NCELL * NCH];
  long long frep[FREP];
  float w[FREP];
  float wg[FREP];
};
__global__ void k_unpack(const RepState* st, HostView* hv) {
  for (int i = threadIdx.x; i < FREP; i += blockDim.x) {
    int cell = i / NFEAT, f = i - cell * NFEAT;
    int off = f * PLANE + pos_of(cell);
    hv->frep[i] = st->x[off];
    hv->w[i] = st->w[off];
    hv->wg[i] = st->wg[off];
  }
  for (int i = threadIdx.x; i < NCELL * NCH; i += blockDim.x) {
    int cell = i / NCH, ch = i - cell * NCH;
    hv->grid[i] = (st->grid[cell][ch >> 5] >> (ch & 31)) & 1u;
  }
}
// overwrite agent weights from reference-order arrays (either may be null)
__global__ void k_pack_weights(RepState* st, const float* w, const float* wg) {
  for (int i = threadIdx.x; i < FREP; i += blockDim.x) {
    int cell = i / NFEAT, f = i - cell * NFEAT;
    int off = f * PLANE + pos_of(cell);
    if (w) st->w[off] = w[i];
    if (wg) st->wg[off] = wg[i];
  }
}

// wNet / wGradCorr of replicas [first, first + count) in the reference's flatten order (either output may be null)
__global__ void k_unpack_agents(const RepState* states, int first, int count, float* w, float* wg, float* avg) {
  const int rep = blockIdx.x;
  if (rep >= count) return;
  const RepState& s = states[first + rep];
  for (int i = threadIdx.x; i < FREP; i += blockDim.x) {
    const int cell = i / NFEAT, f = i - cell * NFEAT;
    const int off = f * PLANE + pos_of(cell);
    if (w) w[(size_t)rep * FREP + i] = s.w[off];
    if (wg) wg[(size_t)rep * FREP + i] = s.wg[off];
  }
  if (avg && threadIdx.x == 0) avg[rep] = s.avg_reward;
}

struct Summary {
  int status;
  int pad;
  long long iter;
  float avg_reward;
  float loss_sum;
  long long loss_n;
  unsigned long long ndraws;
  long long stats[8];
};
__global__ void k_summary(const RepState* states, int n, Summary* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const RepState& s = states[i];
  Summary o;
  o.status = s.initialized ? s.status : ST_UNINITIALIZED;
  o.pad = 0;
  o.iter = s.iter;
  o.avg_reward = s.avg_reward;
  o.loss_sum = s.loss_sum;
  o.loss_n = s.loss_n;
  o.ndraws = s.ndraws;
  for (int k = 0; k < 8; ++k) o.stats[k] = s.stats[k];
  out[i] = o;
}
// out[0..7] = sum of stats, out[8] = sum of iter, out[9] = replicas that failed
__global__ void k_sum_stats(const RepState* states, int n, long long* out) {
  long long acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const RepState& s = states[i];
    for (int k = 0; k < 8; ++k) acc[k] += s.stats[k];
    acc[8] += s.iter;
    acc[9] += (s.status != ST_RUNNING && s.status != ST_SUCCESS) ? 1 : 0;
  }
  for (int k = 0; k < 10; ++k) {
    long long v = acc[k];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane_id() == 0 && v) atomicAdd(reinterpret_cast<unsigned long long*>(&out[k]), (unsigned long long)v);
  }
}
__global__ void k_period_reset(RepState* states, int rep) {
  RepState& s = states[rep];
  s.loss_sum = 0.0f;
  s.loss_n = 0;
  s.stats[6] = 0;
  s.stats[7] = 0;
}

}  // namespace dca
