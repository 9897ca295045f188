// dca_fast.cuh -- the throughput kernel of libdca_b200 (FAST arithmetic order).
//
// A launch is cut into UNITS (replica, chunk of 4096 events), one CTA of 128 threads each, chained per replica:
//   warps 0-2  "agent" threads (96): the value function and the TDC update.
//              Lane r (0..6; lane 7 idles) of 8-lane group g (0..11) owns row r of
//              feature plane f = 12*slot + g for slot = 0..5.  wGradCorr (6 x 7 floats
//              per lane, never read by another thread) lives in Blackwell tensor memory
//              used as a per-thread scratchpad (tcgen05.ld / st), wNet in shared memory
//              (the afterstate evaluation needs it with a flexible lane mapping).
//              96 registers, 44.3 KB -> 5 CTAs per SM.
//   warp 3     "event" warp: event queue, random numbers, statistics, grid bits, candidate
//              lists, stop rules (environmentStep, src/Simulator.hs:101-134).  It runs ONE
//              STEP AHEAD of the agents: what an event needs from the queue depends on
//              whether it has a candidate, not on which channel is chosen.
//
// Per event (runStep, src/SimRunner.hs:60-97) the two sides meet at named barriers that
// carry exactly the data dependencies (see the barrier helpers below):
//   agents:  V(s) = x.w from the plane sums (butterfly), then every candidate channel on
//            its own 8-lane group: the two planes the action touches are re-evaluated and
//            substituted into the V tree (selectAction, src/Simulator.hs:185-205 with
//            incAfterStateFreps, src/Gridfuncs.hs:186-252)
//            -- barrier 2 (agents only): all q written --
//            argmax (last maximum wins, src/AccUtils.hs:203-208) and the TD scalars
//            (src/Agent.hs:62-78), replicated per agent thread; warp 0 publishes the decision
//            -- barrier 3 --
//            dense w / wGradCorr update (src/Agent.hs:67-76) fused with moving frep / cnt2 to
//            the chosen afterstate (the owners of the touched rows do it) and with the next
//            event's row sums
//            -- barrier 4: wait for the next event's record --
//   event:   pop of the NEXT event (ahead of the agents) -- barrier 3: the decision --
//            grid bit (src/Gridfuncs.hs:105-110), END key / reassign, row masks,
//            -- barrier 1: cnt2 is in its afterstate (agent warp 2) -- candidate list of the
//            next event, statistics, stop rules -- barrier 4: publish (no wait) --
//            push of the next event's new events
//
// Instantiations: k_run_fast<false> (the benchmarked one), <true> (parity trace / verify flags: one more barrier per event so
// that the event warp sees the afterstate) and <true, true> (hand-off look-ahead, see q_lookahead; its own instantiation so
// that the other two carry none of its code).  What bounds the kernel is the issue rate it sustains (66-68 % of the issue
// slots, five CTAs sharing four schedulers): the number of instructions per event counts, not which warp executes them
// (profiles/r2_v13_summary.md).
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
constexpr int NBLK = 22;      // queue blocks of 32 slots
constexpr int QCAPS = NBLK * 32;  // 704 >= QCAP: whole blocks can be scanned
static_assert(QCAPS >= QCAP, "queue blocks");

// Instruction-count switches of the agents' loop (late round 2; every one bit-exact, timed with tools/prof_run.py 4096
// 10000 20000 on one B200: the agents' chain is what bounds the kernel, and a static instruction less per agent
// iteration is worth ~0.08 % of throughput -- 918 -> 858 instructions, 324.9 -> 340.1 M events/s):
//   DCA_TADDR_UNIFORM  the tensor-memory address comes out of a warp reduction, i.e. lives in a uniform register: no
//                      R2UR in front of every LDTM / STTM (12 per iteration)                           324.9 -> 328.3
//   DCA_CVT_MIX=2      three of the seven byte -> float conversions of a row on the conversion unit (I2F.U8, one
//                      instruction each) instead of PRMT + FADD(2); more than three saturate that unit
//                      (=1: 329.6, =2: 332.6, =3: 331.8, =4 (all seven): 326.9)                         328.3 -> 332.6
//   DCA_PAD_REUSE      the pad column of a row is stored back as it was loaded: no register zeroed per store   -> 338.6
//   (bmad)             `is_end ? x - m : x + m` on packed bytes as one IMAD per word with the sign in a register   858 -> 826 instr.
//   DCA_SA_DELTA + DCA_STTM_X4  "is this slot's plane the chosen channel" from one difference; the wGradCorr row
//                      stored as two 4-column pieces (shorter register tuples, fewer moves)                    -> 340.1
#ifdef DCA_V12_AGENT_LOOP
#ifndef DCA_CVT_MIX
#define DCA_CVT_MIX 0
#endif
#else
#define DCA_TADDR_UNIFORM
#ifndef DCA_CVT_MIX
#define DCA_CVT_MIX 2
#endif
#define DCA_PAD_REUSE
#define DCA_SA_DELTA
#define DCA_STTM_X4
#endif

// Where wGradCorr lives (see struct Agent) and, with it, how many CTAs share an SM.
#ifndef DCA_NO_END_SLOT_INDEX
#define DCA_END_SLOT_INDEX
#endif
#if !defined(DCA_WG_REGS) && !defined(DCA_WG_TMEM)
#define DCA_WG_TMEM  // measured on B200: 294.4 vs 287.9 M events/s (4096 replicas), 299.6 vs 292.1 M (5920)
#endif
// With wGradCorr in registers: 4 CTAs per SM (128 registers/thread, no spills) beat 5 (96 registers,
// spills): 287 vs 258 M events/s on B200.  With wGradCorr in tensor memory the kernel fits 96 registers.
#ifndef DCA_MIN_CTAS
#ifdef DCA_WG_TMEM
#define DCA_MIN_CTAS 5
#else
#define DCA_MIN_CTAS 4
#endif
#endif

struct Ev {  // an event and its candidate set (double-buffered: the event warp writes
             // the next one while the agents still use the current one)
  double time;
  int type, cell, ch, hoff;
  int ncand, pad;
  uint2 n4b[8], n2b[8];  // row byte masks of the event cell: N4 \ {cell}, N2 u {cell}
  uint8_t cand[72];      // candidate channels, ascending
};

struct LaRec {      // the channels eligible at the hand-off target cell on the current grid (event warp -> agents)
  uint32_t hm[3];   // as a mask ...
  int nh;           // ... their number ...
  uint8_t hcand[72];  // ... and ascending
};
struct alignas(16) SmemF {
  float w[2][WH];           // wNet: columns 0-3 of row rho = f*7+r at [0][rho*4+c], columns 4-6 at [1][rho*4+c-4]
  uint8_t x[WPAD];          // frep, [f][r][8]
  uint8_t cnt2[WPAD];       // users of ch within distance 2, [ch][r][8]
  uint32_t grid[NCELL][4];
  double q_time[QCAPS];     // event queue slots: 0..48 NEW (one per cell), 49.. END; unused = +inf
  uint32_t q_id[QCAPS];
  uint16_t q_key[QCAP];
  uint8_t q_hoff[QCAP];
#ifdef DCA_END_SLOT_INDEX
  uint16_t end_slot[NCELL][72];  // (cell, ch) -> slot of the END event of the call on ch at cell (7 KB; without it
                                 // reassign finds the slot by a warp-wide search of q_key and 6 CTAs fit an SM:
                                 // measured 293 / 296 M events/s with 5 / 6 CTAs against 306 M with the index and 5)
#endif
  Ev ev[3];                 // the event warp runs ahead: it may fill record k + 2 while the agents still read record k
  float P[72];              // plane sums of x.w (current state, current weights)
  float q[72];              // afterstate values of the candidates
  float Wg[4];              // per-warp partial sums of x.wGradCorr
  // the decision of the current event, published by agent warp 0 for the event warp (which then needs no argmax of
  // its own): the chosen channel (-1 = Nothing), the TD error, and -- for the trace -- avgReward and the calls in progress
  struct alignas(16) Pub { int act; float tderr; float avg_reward; int ncalls; } pub;
  float alpha_net, alpha_avg, alpha_grad;  // alphaNet, alphaAvg, alphaGrad of this replica
  int status;
  uint32_t tmem_base;       // tensor-memory columns of this CTA (wGradCorr, DCA_WG_TMEM)
  float v_pub;              // (step kernel only)
  uint32_t pad_[2];
  uint2 n4rows[NCELL];      // per cell: byte r = the 7 column bits of row r of N4(cell) \ {cell} ...
  uint2 n2rows[NCELL];      // ... and of N2(cell) u {cell}  (copied from the constant tables: a shared-memory load
                            // with a computed index is ~5 x faster than a constant-memory one)
};
// what the hand-off look-ahead instantiation keeps on top (its own type: the hot instantiations are compiled against SmemF
// alone and stay byte for byte what they were)
struct alignas(16) SmemLA : SmemF {
  float qla[72];            // look-ahead values of the candidates of a hand-off END (the selection; sm.q keeps V(afterstate_k))
  LaRec la[3];              // per event record: the channels eligible at the hand-off target
  uint32_t pad_la_[2];
};
__device__ __forceinline__ SmemLA& la_of(SmemF& sm) { return static_cast<SmemLA&>(sm); }
__device__ __forceinline__ const SmemLA& la_of(const SmemF& sm) { return static_cast<const SmemLA&>(sm); }
static_assert(sizeof(SmemLA) <= (233472 / DCA_MIN_CTAS - 1024), "shared memory per CTA: DCA_MIN_CTAS of them must fit an SM");
static_assert(sizeof(SmemF) % 16 == 0 && sizeof(SmemLA) % 16 == 0, "SmemF size");
static_assert(offsetof(SmemF, x) % 8 == 0 && offsetof(SmemF, cnt2) % 8 == 0 && offsetof(SmemF, grid) % 16 == 0 &&
              offsetof(SmemF, q_time) % 16 == 0 && offsetof(SmemF, q_id) % 16 == 0 && offsetof(SmemF, q_key) % 16 == 0 &&
              offsetof(SmemF, q_hoff) % 8 == 0 && offsetof(SmemF, ev) % 8 == 0 &&
              sizeof(Ev) % 8 == 0, "SmemF alignment");

__device__ __forceinline__ void cta_barrier() { asm volatile("bar.sync 0, 128;" ::: "memory"); }
// The per-event synchronisation of the run kernel is producer / consumer, so that neither side waits for the other
// where it does not need its data:
//   barrier 2 (96):  the agents among themselves -- every candidate's value is in sm.q
//   barrier 3 (64):  the decision (sm.pub) is published: agent warp 0 arrives right after its argmax, the event warp waits
//   barrier 4 (128): the next event's record (kind, cell, masks, candidates) and the stop status are published and the
//                    plane sums of the dense pass are complete: the event warp arrives (and runs on to the next
//                    event's queue work), the agents wait
//   barrier 5 (128): TRACE instantiation only -- the afterstate is complete (the event warp hashes / verifies it)
__device__ __forceinline__ void agents_q_ready() { asm volatile("bar.sync 2, 96;" ::: "memory"); }
__device__ __forceinline__ void decision_published() { asm volatile("bar.arrive 3, 64;" ::: "memory"); }
__device__ __forceinline__ void event_wait_decision() { asm volatile("bar.sync 3, 64;" ::: "memory"); }
__device__ __forceinline__ void event_publish_next() { asm volatile("bar.arrive 4, 128;" ::: "memory"); }
__device__ __forceinline__ void agents_wait_next() { asm volatile("bar.sync 4, 128;" ::: "memory"); }
__device__ __forceinline__ void agents_afterstate_done() { asm volatile("bar.arrive 5, 128;" ::: "memory"); }
__device__ __forceinline__ void event_wait_afterstate() { asm volatile("bar.sync 5, 128;" ::: "memory"); }
// named barrier 1, 64 threads (see agent_candidates): agent warp 2 arrives once cnt2 is in its afterstate, the event warp
// waits for that before it builds the next event's candidate list
__device__ __forceinline__ void cand_signal() { asm volatile("bar.arrive 1, 64;" ::: "memory"); }
__device__ __forceinline__ void cand_wait() { asm volatile("bar.sync 1, 64;" ::: "memory"); }

#ifdef DCA_PHASE_CLOCK  // development only: cycles per role (virtual warp) and phase, summed over CTAs and events
__device__ unsigned long long g_phase[6][8];  // rows 4, 5: sub-phases of the event warp's update phase
__device__ unsigned long long g_cta[8192][2];  // per block: cycles in the run kernel, SM id
#define DCA_CLK(v) const unsigned v = (unsigned)clock64()  /* 32-bit stamps and sums: few registers */
// after a barrier: BAR.SYNC.DEFER_BLOCKING lets the warp run on to its first shared-memory access, so wait for one
#define DCA_CLKB(v)                                            \
  {                                                            \
    const int s_ = *(volatile int*)&sm.status;                 \
    if (s_ == 0x7ffffff0) asm volatile("trap;");               \
  }                                                            \
  const unsigned v = (unsigned)clock64()
#define DCA_ACC(i, a, b) ph[i] += (b) - (a)
#define DCA_SUB(i, a, b) sub[i] += (b) - (a)
#else
#define DCA_SUB(i, a, b)
#define DCA_CLK(v)
#define DCA_CLKB(v)
#define DCA_ACC(i, a, b)
#endif

// ---------------------------------------------------------------------------
// packed fp32 helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
// bytes k, k+1 of v as floats: 0x4B0000bb is 2^23 + bb exactly, so PRMT + FADD2 converts on the
// ALU / FMA pipes.  DCA_CVT_I2F: I2F.U8 with a byte selector on the XU pipe instead (fewer
// instructions, longer latency) -- measured 2-7 % slower on B200.
template <int K>
__device__ __forceinline__ float2 cvt_pair(uint32_t v) {
#ifdef DCA_CVT_I2F
  return make_float2((float)((v >> (8 * K)) & 0xFFu), (float)((v >> (8 * K + 8)) & 0xFFu));
#endif
  float2 m = make_float2(__uint_as_float(__byte_perm(v, 0x4B000000u, 0x7440 + K)),
                         __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7440 + K + 1)));
  return __fadd2_rn(m, f2(-8388608.0f));
}
template <int K>
__device__ __forceinline__ float cvt_one(uint32_t v) {
#ifdef DCA_CVT_I2F
  return (float)((v >> (8 * K)) & 0xFFu);
#endif
  return __fadd_rn(__uint_as_float(__byte_perm(v, 0x4B000000u, 0x7440 + K)), -8388608.0f);
}
struct Row7 {  // one row of the feature representation as floats
  float2 a, b, c;
  float d;
};
#ifndef DCA_CVT_MIX_EVAL  // the candidate evaluation is not where the conversion unit is busy: all seven there (345.0 -> 346.6)
#ifdef DCA_V12_AGENT_LOOP
#define DCA_CVT_MIX_EVAL 0
#else
#define DCA_CVT_MIX_EVAL 4
#endif
#endif
// MIX of the seven columns go through the conversion unit (I2F.U8 with a byte selector, one instruction each), the
// others through PRMT + FADD(2) on the ALU / FMA pipes: 0 = none, 1 = column 6, 2 = columns 4-6, 3 = 2-6, 4 = all
template <int MIX = DCA_CVT_MIX>
__device__ __forceinline__ Row7 cvt_row(uint2 xb) {
  Row7 x;
  if (MIX >= 4) x.a = make_float2((float)(xb.x & 0xFFu), (float)((xb.x >> 8) & 0xFFu));
  else x.a = cvt_pair<0>(xb.x);
  if (MIX >= 3) x.b = make_float2((float)((xb.x >> 16) & 0xFFu), (float)(xb.x >> 24));
  else x.b = cvt_pair<2>(xb.x);
  if (MIX >= 2) x.c = make_float2((float)(xb.y & 0xFFu), (float)((xb.y >> 8) & 0xFFu));
  else x.c = cvt_pair<0>(xb.y);
  if (MIX >= 1) x.d = (float)((xb.y >> 16) & 0xFFu);
  else x.d = cvt_one<2>(xb.y);
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
// xor-shuffle whose position relative to the other volatile shuffles is kept by the compiler: the steps of
// a butterfly then issue together (ptxas otherwise interleaves steps and lengthens the latency chain)
__device__ __forceinline__ float shfl_xor_ordered(const float v, const int m) {
  float r;
  asm volatile("shfl.sync.bfly.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=f"(r) : "f"(v), "r"(m));
  return r;
}
__device__ __forceinline__ void st_shared_f2(float* p, const float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" :: "r"((uint32_t)__cvta_generic_to_shared(p)), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void st_shared_f1(float* p, const float v) {
  asm volatile("st.shared.f32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(p)), "f"(v) : "memory");
}
__device__ __forceinline__ uint2 badd(uint2 a, uint2 b) { return make_uint2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ uint2 bsub(uint2 a, uint2 b) { return make_uint2(a.x - b.x, a.y - b.y); }
// a + s * b on packed bytes, s = +-1 as a register (s = -1: a - b mod 2^32, bytewise exact while no byte borrows): one
// IMAD per word where `is_end ? bsub : badd` costs a negation or a select more
__device__ __forceinline__ uint2 bmad(uint2 a, uint2 b, int s) {
#ifdef DCA_V12_AGENT_LOOP
  return s < 0 ? bsub(a, b) : badd(a, b);
#else
  return make_uint2(a.x + b.x * (uint32_t)s, a.y + b.y * (uint32_t)s);
#endif
}

// ---------------------------------------------------------------------------
// Agent side
// ---------------------------------------------------------------------------
// wGradCorr of an agent lane: 6 rows x 7 floats that no other thread ever reads.
//  * DCA_WG_REGS: 42 registers per thread, which caps the kernel at 4 CTAs per SM;
//  * DCA_WG_TMEM (default): Blackwell tensor memory used as a per-thread scratchpad -- lane l of hardware warp w
//    owns TMEM lane 32 (w % 4) + l, slot j sits in columns 8 j .. 8 j + 6 of the CTA's 64 columns
//    (tcgen05.ld / st .32x32b.x8: one row per instruction, ~12 cycles), and the kernel fits 5 CTAs per SM.
struct WgRow {
  float2 g01, g23, g45;
  float g6;
};
#ifdef DCA_WG_TMEM
struct Agent {
  uint32_t taddr;  // TMEM address (lane quarter of this warp, first column of the CTA)
};
struct WgRegs {    // a row in flight: valid after wg_wait
  uint32_t r[8];
};
__device__ __forceinline__ WgRegs wg_get(const Agent& ag, const int j) {
  WgRegs v;
#ifdef DCA_LDTM_X4
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.r[0]), "=r"(v.r[1]), "=r"(v.r[2]), "=r"(v.r[3]) : "r"(ag.taddr + 8u * (uint32_t)j));
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.r[4]), "=r"(v.r[5]), "=r"(v.r[6]), "=r"(v.r[7]) : "r"(ag.taddr + 8u * (uint32_t)j + 4u));
  return v;
#endif
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v.r[0]), "=r"(v.r[1]), "=r"(v.r[2]), "=r"(v.r[3]), "=r"(v.r[4]), "=r"(v.r[5]), "=r"(v.r[6]), "=r"(v.r[7])
               : "r"(ag.taddr + 8u * (uint32_t)j));
  return v;
}
__device__ __forceinline__ WgRow wg_wait(WgRegs& v) {  // (the registers are operands so that no use moves above the wait)
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v.r[0]), "+r"(v.r[1]), "+r"(v.r[2]), "+r"(v.r[3]), "+r"(v.r[4]), "+r"(v.r[5]), "+r"(v.r[6]), "+r"(v.r[7]));
  WgRow g;
  g.g01 = make_float2(__uint_as_float(v.r[0]), __uint_as_float(v.r[1]));
  g.g23 = make_float2(__uint_as_float(v.r[2]), __uint_as_float(v.r[3]));
  g.g45 = make_float2(__uint_as_float(v.r[4]), __uint_as_float(v.r[5]));
  g.g6 = __uint_as_float(v.r[6]);
  return g;
}
__device__ __forceinline__ void wg_put(Agent& ag, const int j, const WgRow& g, const uint32_t pad = 0u) {
#ifdef DCA_STTM_X4
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               :: "r"(ag.taddr + 8u * (uint32_t)j), "r"(__float_as_uint(g.g01.x)), "r"(__float_as_uint(g.g01.y)),
                  "r"(__float_as_uint(g.g23.x)), "r"(__float_as_uint(g.g23.y)) : "memory");
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               :: "r"(ag.taddr + 8u * (uint32_t)j + 4u), "r"(__float_as_uint(g.g45.x)), "r"(__float_as_uint(g.g45.y)),
                  "r"(__float_as_uint(g.g6)), "r"(pad) : "memory");
  return;
#endif
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :: "r"(ag.taddr + 8u * (uint32_t)j), "r"(__float_as_uint(g.g01.x)), "r"(__float_as_uint(g.g01.y)),
                  "r"(__float_as_uint(g.g23.x)), "r"(__float_as_uint(g.g23.y)), "r"(__float_as_uint(g.g45.x)),
                  "r"(__float_as_uint(g.g45.y)), "r"(__float_as_uint(g.g6)), "r"(pad) : "memory");
}
__device__ __forceinline__ void wg_commit() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
constexpr uint32_t TMEM_COLS = 64;
static_assert(8 * NSLOT <= TMEM_COLS, "six rows of eight columns per lane");
static_assert(DCA_MIN_CTAS * TMEM_COLS <= 512, "an SM has 512 tensor-memory columns: every resident CTA must get its share");
// CTA-wide: hardware warp 0 allocates, everybody learns the base
__device__ __forceinline__ void tmem_open(SmemF& sm, Agent& ag) {
  if ((threadIdx.x >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"((uint32_t)__cvta_generic_to_shared(&sm.tmem_base)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifdef DCA_TADDR_UNIFORM
  // (the result of a warp reduction lives in a uniform register: the tensor-memory instructions take their address from
  //  one, and ptxas otherwise re-converts the address before every one of them)
  ag.taddr = __reduce_max_sync(0xffffffffu, sm.tmem_base + ((threadIdx.x >> 5) * 32u << 16));
#else
  ag.taddr = sm.tmem_base + ((threadIdx.x >> 5) * 32u << 16);
#endif
}
__device__ __forceinline__ void tmem_close(SmemF& sm) {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if ((threadIdx.x >> 5) == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(sm.tmem_base), "r"(TMEM_COLS) : "memory");
}
#else
struct Agent {       // per-thread registers of an agent thread
  WgRow g[NSLOT];    // wGradCorr rows of this lane, one per slot
};
struct WgRegs {
  WgRow g;
};
__device__ __forceinline__ WgRegs wg_get(const Agent& ag, const int j) { return WgRegs{ag.g[j]}; }
__device__ __forceinline__ WgRow wg_wait(WgRegs& v) { return v.g; }
__device__ __forceinline__ void wg_put(Agent& ag, const int j, const WgRow& g, const uint32_t = 0u) { ag.g[j] = g; }
__device__ __forceinline__ void wg_commit() {}
__device__ __forceinline__ void tmem_open(SmemF&, Agent&) {}
__device__ __forceinline__ void tmem_close(SmemF&) {}
#endif

struct TdScalars {
  float ctd, cavg, ncdot, sg;
};

// Per-thread constants of an agent lane: where its rows live.
struct Lane {
  int vw;             // virtual warp (role) of this thread: 0..2 agent, 3 event
  int g8, r;          // lane group (0..11), row (0..7; 7 idles)
  int woff;           // float offset of this lane's row in a half of sm.w for slot 0 (+336 per slot)
  bool rv;            // row valid
  int pidx;           // index in sm.P of the plane sum this lane ends up with in dense_pass (-1: none)
};
// Which hardware warp plays the event role.  The hardware already spreads the warps of the four CTAs of
// an SM over the four sub-partitions: the CTA in slot k gets warp slots [4k + (k + w) % 4] (measured:
// [0 1 2 3], [5 6 7 4], [10 11 8 9], [15 12 13 14]; tools/microbench/placement.cu), so CTAs that share an
// SM must use the SAME role assignment for their event warps to land on four different schedulers
// (measured on one resident wave: 277 M events/s with equal assignments, 267-272 M with mixed ones).
// The assignment therefore only changes from wave to wave (block ids of one wave share it).
__device__ __forceinline__ int role_rotation() {
#ifdef DCA_NO_ROTATE_ROLES
  return 0;
#else
  unsigned nsm;
  asm("mov.u32 %0, %%nsmid;" : "=r"(nsm));
  return (int)((blockIdx.x / ((unsigned)DCA_MIN_CTAS * nsm)) & 3u);
#endif
}
__device__ __forceinline__ int virtual_tid() { return (int)((threadIdx.x + 32u * (unsigned)role_rotation()) & 127u); }
__device__ __forceinline__ Lane make_lane() {
  Lane l;
  const int vt = virtual_tid();
  l.vw = vt >> 5;
  l.g8 = vt >> 3;
  l.r = vt & 7;
  l.rv = l.r < 7;
  const int rc = l.rv ? l.r : 6;
  l.woff = (l.g8 * 7 + rc) * 4;
  {  // see the reduce-scatter in dense_pass: row 0 -> slot 0, 1 -> 3, 2 -> 2, 3 -> 5, 4 -> 1, 5 -> 4 (rows 6, 7 repeat 2, 3)
    const bool b0 = l.r & 1, b1 = l.r & 2, b2 = l.r & 4;
    const int slot = b1 ? (b0 ? 5 : 2) : (b2 ? (b0 ? 4 : 1) : (b0 ? 3 : 0));
    l.pidx = (l.r < 6 && l.vw < EVW && (slot < NSLOT - 1 || l.g8 != SLOTP - 1)) ? SLOTP * slot + l.g8 : -1;
  }
  return l;
}

// The dense pass over this lane's rows (one row of one plane per slot).
// UPDATE: sm.x holds the PRE-action features x.  The owners of the rows the
// action touches move them to the afterstate x' (plane `act`: +-1 on N4 \ {cell},
// Gridfuncs.hs:212-217; plane 70 and cnt2[act]: Gridfuncs.hs:220-252), then
//   g = fma(-cdot, x', fma(ctd, x, cavg));  w' = w - g;  wg' = fma(sg, x, wg)   (Agent.hs:67-76)
// and the row sums of x'.w' -> plane sums P, x'.wg' -> warp sums Wg (for the next event).
// Without UPDATE only the sums of the current state are formed.
template <bool UPDATE, bool SIGNAL = false>
__device__ __forceinline__ void dense_pass(SmemF& sm, Agent& ag, const Lane& l, const int act, const bool is_end,
                                           const TdScalars td, const uint2 n4b_r, const uint2 n2b_r) {
  const int wid = l.vw;
  float rs[NSLOT];
  float rg_acc = 0.0f;
  const Row7 dm = cvt_row(n4b_r);            // the row of N4 \ {cell} as 0.0 / 1.0
  const float sgn = is_end ? -1.0f : 1.0f;
#ifdef DCA_SA_DELTA
  const int dact = act - l.g8;
  // does this lane own a row of plane `act`?  (act - g8 in {0, 12, .., 60}; act = -1 gives -1 - g8 < 0)
  bool own_act = (unsigned)dact < (unsigned)(SLOTP * NSLOT) && ((unsigned)dact * 43u >> 9) * (unsigned)SLOTP == (unsigned)dact;
  (void)dact;
#else
  bool own_act = false;                      // does this lane own a row of plane `act`?
#endif
  (void)dm; (void)sgn;
  // The warp that owns plane 70 (n_elig) also moves cnt2[act] to the afterstate; it does so up front, where
  // the load latency hides behind the first slot: cg = the cells of N2 u {cell} where `act` is eligible on
  // grid' (cnt2 == 0, or == 1 when the call at `cell` ends).  The four groups hold the same row: same values.
  uint2 cg = make_uint2(0, 0);
  if (UPDATE && wid == 2 && act >= 0) {      // warp-uniform
    const int co = act * PLANE + (l.rv ? l.r : 6) * ROWPAD;
    const uint2 cr = *reinterpret_cast<const uint2*>(&sm.cnt2[co]);
    const uint32_t want = is_end ? 1u : 0u;
    cg = make_uint2(n2b_r.x & eq_bytes(cr.x, want), n2b_r.y & eq_bytes(cr.y, want));
    *reinterpret_cast<uint2*>(&sm.cnt2[co]) = bmad(cr, n2b_r, is_end ? -1 : 1);
  }
  if (UPDATE && SIGNAL && wid == 2) cand_signal();  // warp-uniform: cnt2 is in its afterstate
#pragma unroll
  for (int j = 0; j < NSLOT; ++j) {
    const bool last = j == NSLOT - 1;
    // slot 5 holds planes 60..70 on groups 0..10; group 11 has no plane there and shadows plane 70
    const bool pv = last ? l.g8 != SLOTP - 1 : true;
    const int f = SLOTP * j + l.g8 - (pv ? 0 : 1);
    const bool valid = pv && l.rv;
    const int wo_ = l.woff + j * SLOTP * 28 - (pv ? 0 : 28);
    const int xo_ = 2 * wo_;  // byte offset of the same row in sm.x: 56 f + 8 r = 2 (28 f + 4 r)
    const uint2 xb = *reinterpret_cast<const uint2*>(&sm.x[xo_]);
    const float4 wa = *reinterpret_cast<const float4*>(&sm.w[0][wo_]);
    const float4 wb = *reinterpret_cast<const float4*>(&sm.w[1][wo_]);
    WgRegs gin = wg_get(ag, j);
    float2 w01 = make_float2(wa.x, wa.y), w23 = make_float2(wa.z, wa.w), w45 = make_float2(wb.x, wb.y);
    float w6 = wb.z;
    const Row7 xo = cvt_row(xb);
    WgRow G = wg_wait(gin);
    Row7 xn = xo;
    if (UPDATE) {
      // x' = x + s * [cell in N4 \ {cell}], s = +-1 on plane `act` and 0 elsewhere: small integers, exact in
      // fp32, and branch-free, so that the six slot bodies form one basic block
#ifdef DCA_SA_DELTA
      // plane `act` sits in slot j of lane group g8 iff act - g8 == 12 j (group 11 shadows plane 70 in the last slot:
      // act - 11 == 60 would be channel 71)
      const float sa = dact == SLOTP * j ? sgn : 0.0f;
#else
      const float sa = f == act ? sgn : 0.0f;
      own_act = own_act || f == act;
#endif
      xn.a = __ffma2_rn(f2(sa), dm.a, xo.a);
      xn.b = __ffma2_rn(f2(sa), dm.b, xo.b);
      xn.c = __ffma2_rn(f2(sa), dm.c, xo.c);
      xn.d = __fmaf_rn(sa, dm.d, xo.d);
      if (last && wid == 2 && act >= 0) {  // warp-uniform: the owners of plane 70 (n_elig)
        if (f == NCH) {  // (the shadow lanes -- row 7, group 11 -- compute and store the same values as their originals)
          const uint2 eb = bmad(xb, cg, is_end ? 1 : -1);
          xn = cvt_row(eb);
          *reinterpret_cast<uint2*>(&sm.x[xo_]) = eb;
        }
      }
#define DCA_UPD2(W, G, XN, XO)                                            \
  {                                                                       \
    const float2 t_ = __ffma2_rn(f2(td.ctd), XO, f2(td.cavg));            \
    const float2 g_ = __ffma2_rn(f2(td.ncdot), XN, t_);                   \
    W = __ffma2_rn(g_, f2(-1.0f), W);                                     \
    G = __ffma2_rn(f2(td.sg), XO, G);                                     \
  }
      DCA_UPD2(w01, G.g01, xn.a, xo.a)
      DCA_UPD2(w23, G.g23, xn.b, xo.b)
      DCA_UPD2(w45, G.g45, xn.c, xo.c)
#undef DCA_UPD2
      {
        const float t_ = __fmaf_rn(td.ctd, xo.d, td.cavg);
        const float g_ = __fmaf_rn(td.ncdot, xn.d, t_);
        w6 = __fsub_rn(w6, g_);
        G.g6 = __fmaf_rn(td.sg, xo.d, G.g6);
#if defined(DCA_PAD_REUSE) && defined(DCA_WG_TMEM)
        wg_put(ag, j, G, gin.r[7]);  // (the pad column goes back as it came: no register has to be zeroed for it)
#else
        wg_put(ag, j, G);
#endif
      }
      // unconditional: a shadow lane (row 7 -> row 6, group 11 in the last slot -> plane 70) holds the same
      // x, w and wGradCorr as its original and stores the same values to the same address
#ifdef DCA_W_ST64
      // (pairs as they come out of the packed FMAs: an STS.128 wants four consecutive registers and costs register moves;
      //  the same number of shared-memory wavefronts; the pad column keeps its zero)
      st_shared_f2(&sm.w[0][wo_], w01);
      st_shared_f2(&sm.w[0][wo_ + 2], w23);
      st_shared_f2(&sm.w[1][wo_], w45);
      st_shared_f1(&sm.w[1][wo_ + 2], w6);
#else
      *reinterpret_cast<float4*>(&sm.w[0][wo_]) = make_float4(w01.x, w01.y, w23.x, w23.y);
#ifdef DCA_PAD_REUSE
      *reinterpret_cast<float4*>(&sm.w[1][wo_]) = make_float4(w45.x, w45.y, w6, wb.w);  // (wb.w == 0: the pad column)
#else
      *reinterpret_cast<float4*>(&sm.w[1][wo_]) = make_float4(w45.x, w45.y, w6, 0.0f);
#endif
#endif
    }
    rs[j] = valid ? rowdot2(xn, w01, w23, w45, w6) : 0.0f;
    const float rg = rowdot2(xn, G.g01, G.g23, G.g45, G.g6);
    if (j == 0) rg_acc = l.rv ? rg : 0.0f;
    else if (valid) rg_acc = __fadd_rn(rg_acc, rg);
  }
  if (UPDATE && own_act && l.rv) {  // the afterstate row of plane `act` (every slot has read its x by now)
    uint2* xp = reinterpret_cast<uint2*>(&sm.x[act * PLANE + l.r * ROWPAD]);
    const uint2 xb_act = *xp;
    *xp = bmad(xb_act, n4b_r, is_end ? -1 : 1);
  }
  // plane sums: the xor-butterfly 1,2,4 over the 8 lanes (rows) of a group, for six slots at once as a
  // reduce-scatter -- at every step a lane keeps half of its values and hands the other half to its
  // partner, so the six sums cost 3 + 2 + 1 shuffles instead of 18.  Every sum is still formed as
  // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+0)) (fp32 addition commutes), bit for bit the plane tree of the
  // FAST order.  The first three steps of the warp sum of x.wg ride along.
  {
    const bool b0 = l.r & 1, b1 = l.r & 2, b2 = l.r & 4;
    // step 1: even rows keep slots 0-2, odd rows slots 3-5
    const float k0 = b0 ? rs[3] : rs[0], k1 = b0 ? rs[4] : rs[1], k2 = b0 ? rs[5] : rs[2];
    const float s0 = b0 ? rs[0] : rs[3], s1 = b0 ? rs[1] : rs[4], s2 = b0 ? rs[2] : rs[5];
    const float x0 = shfl_xor_ordered(s0, 1), x1 = shfl_xor_ordered(s1, 1), x2 = shfl_xor_ordered(s2, 1);
    const float xg1 = shfl_xor_ordered(rg_acc, 1);
    const float t0 = __fadd_rn(k0, x0), t1 = __fadd_rn(k1, x1), t2 = __fadd_rn(k2, x2);
    rg_acc = __fadd_rn(rg_acc, xg1);
    // step 2: rows 0,1,4,5 keep the first two of the three, rows 2,3,6,7 the third
    const float y0 = shfl_xor_ordered(b1 ? t0 : t2, 2), y1 = shfl_xor_ordered(t1, 2);
    const float xg2 = shfl_xor_ordered(rg_acc, 2);
    const float u0 = __fadd_rn(b1 ? t2 : t0, y0), u1 = __fadd_rn(t1, y1);
    rg_acc = __fadd_rn(rg_acc, xg2);
    // step 3: rows 0,1 keep the first, rows 4,5 the second; rows 2,3 and 6,7 both finish the third
    const float keep = (b2 && !b1) ? u1 : u0, send = (b2 || b1) ? u0 : u1;
    const float z = shfl_xor_ordered(send, 4);
    const float xg3 = shfl_xor_ordered(rg_acc, 4);
    const float v = __fadd_rn(keep, z);
    rg_acc = __fadd_rn(rg_acc, xg3);
    if (l.pidx >= 0) sm.P[l.pidx] = v;
  }
  // warp sum of x.wg: xor-butterfly 1,2,4 (above), 8, 16
#pragma unroll
  for (int s = 8; s < 32; s <<= 1) rg_acc = __fadd_rn(rg_acc, __shfl_xor_sync(0xffffffffu, rg_acc, s));
  if (lane_id() == 0) sm.Wg[wid] = rg_acc;
  if (UPDATE) wg_commit();  // the rows written above are read again by the next pass
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
// substituting the channel plane and the n_elig plane.  The 12 agent lane groups
// take one candidate each per pass.
struct QConst {    // per-event constants of a lane in the candidate evaluation
  uint2 x70;
  float4 e_a, e_b;
};
__device__ __forceinline__ QConst q_const(const SmemF& sm, const int rc) {
  QConst c;
  c.x70 = *reinterpret_cast<const uint2*>(&sm.x[NCH * PLANE + rc * ROWPAD]);
  c.e_a = *reinterpret_cast<const float4*>(&sm.w[0][(NCH * 7 + rc) * 4]);
  c.e_b = *reinterpret_cast<const float4*>(&sm.w[1][(NCH * 7 + rc) * 4]);
  return c;
}
// Everything one candidate channel needs from shared memory, loaded up front (one round trip after the
// channel number is known): row rc of its feature plane, of its weights and of cnt2, and the three plane
// sums that share its leaf of the V tree.
struct CandLoads {
  uint2 xr, cr;
  float4 wa, wb;
  float l0, l1, l2;
};
__device__ __forceinline__ CandLoads q_loads(const SmemF& sm, const int ch, const int rc) {
  CandLoads c;
  const int row = ch * 7 + rc, lc = ch & 31;
  c.xr = *reinterpret_cast<const uint2*>(&sm.x[ch * PLANE + rc * ROWPAD]);
  c.cr = *reinterpret_cast<const uint2*>(&sm.cnt2[ch * PLANE + rc * ROWPAD]);
  c.wa = *reinterpret_cast<const float4*>(&sm.w[0][row * 4]);
  c.wb = *reinterpret_cast<const float4*>(&sm.w[1][row * 4]);
  c.l0 = sm.P[lc];
  c.l1 = sm.P[lc + 32];
  c.l2 = sm.P[lc < 6 ? lc + 64 : NCH + 1];  // P[71] == 0
  return c;
}
// row rc of the two re-evaluated planes of afterstate `ch` (branch-free; a row-7 lane shadows row 6 and is
// masked at the end)
__device__ __forceinline__ void q_rows(const CandLoads& c, const QConst& qc, const bool rv, const bool is_end,
                                       const uint2 n4b_r, const uint2 n2b_r, float& rowP, float& rowE) {
  const uint32_t want = is_end ? 1u : 0u;
  // channel plane: x' = x +- [cell in N4 \ {cell}]   (Gridfuncs.hs:212-217)
  const uint2 xr = bmad(c.xr, n4b_r, is_end ? -1 : 1);
  const float rp = rowdot2(cvt_row<DCA_CVT_MIX_EVAL>(xr), make_float2(c.wa.x, c.wa.y), make_float2(c.wa.z, c.wa.w),
                           make_float2(c.wb.x, c.wb.y), c.wb.z);
  // n_elig plane: e' = e -+ [cell in N2 u {cell} and ch eligible there on grid']  (Gridfuncs.hs:220-252)
  const uint2 cg = make_uint2(n2b_r.x & eq_bytes(c.cr.x, want), n2b_r.y & eq_bytes(c.cr.y, want));
  const uint2 er = bmad(qc.x70, cg, is_end ? 1 : -1);
  const float re = rowdot2(cvt_row<DCA_CVT_MIX_EVAL>(er), make_float2(qc.e_a.x, qc.e_a.y), make_float2(qc.e_a.z, qc.e_a.w),
                           make_float2(qc.e_b.x, qc.e_b.y), qc.e_b.z);
  rowP = rv ? rp : 0.0f;
  rowE = rv ? re : 0.0f;
}
// substitute leaf (ch & 31) of the V tree by the re-evaluated channel plane, and plane 70 by Ep
__device__ __forceinline__ float q_substitute(const CandLoads& c, const VTree& vt, const int ch, const float Pp,
                                              const float Ep) {
  const int lc = ch & 31, hi = ch >> 5;
  const float p0 = hi == 0 ? Pp : c.l0;
  const float p1 = hi == 1 ? Pp : c.l1;
  const float p2 = hi == 2 ? Pp : c.l2;
  const float s0 = __shfl_sync(0xffffffffu, vt.sib[0], lc), s1 = __shfl_sync(0xffffffffu, vt.sib[1], lc),
              s2 = __shfl_sync(0xffffffffu, vt.sib[2], lc), s3 = __shfl_sync(0xffffffffu, vt.sib[3], lc),
              s4 = __shfl_sync(0xffffffffu, vt.sib[4], lc);
  float b = __fadd_rn(__fadd_rn(p0, p1), p2);
  b = __fadd_rn(b, s0); b = __fadd_rn(b, s1); b = __fadd_rn(b, s2); b = __fadd_rn(b, s3); b = __fadd_rn(b, s4);
  return __fadd_rn(b, Ep);
}
// The agents' evaluation phase: V(s) and every q[k].  The V butterfly (5 steps) and the row
// butterflies of the first pass of candidates (3 steps each) are independent shuffle chains and
// are issued interleaved.  Candidate k of a pass sits on lane group k % 12; its channel number is
// read before the candidate count is known (a stale entry is a valid channel and is masked later).
__device__ __forceinline__ VTree q_phase(SmemF& sm, const Lane& l, const Ev& ev, const int n, const bool is_end,
                                         const uint2 n4b_r, const uint2 n2b_r) {
  const int lane = lane_id();
  const int rc = l.rv ? l.r : 6;
  const int wfirst = l.vw * 4;    // first lane group of this warp
  const bool have = wfirst < n;   // warp-uniform: this warp has a candidate in the first pass
  VTree t;
  int ch = ev.cand[l.g8];
  const float p0 = sm.P[lane], p1 = sm.P[lane + 32];
  const float p2 = lane < 6 ? sm.P[lane + 64] : 0.0f;
  const float p70 = sm.P[NCH];
  const float wg0 = sm.Wg[0], wg1 = sm.Wg[1], wg2 = sm.Wg[2];
  QConst qc;
  CandLoads c;
  float Pp = 0.0f, Ep = 0.0f;
  if (have) {
    qc = q_const(sm, rc);
    c = q_loads(sm, ch, rc);
    q_rows(c, qc, l.rv, is_end, n4b_r, n2b_r, Pp, Ep);
  }
  float b = __fadd_rn(__fadd_rn(p0, p1), p2);
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    t.sib[i] = __shfl_xor_sync(0xffffffffu, b, 1 << i);
    b = __fadd_rn(b, t.sib[i]);
    if (i < 3) {
      Pp = __fadd_rn(Pp, __shfl_xor_sync(0xffffffffu, Pp, 1 << i));
      Ep = __fadd_rn(Ep, __shfl_xor_sync(0xffffffffu, Ep, 1 << i));
    }
  }
  t.val = __fadd_rn(b, p70);
  t.dotg = __fadd_rn(__fadd_rn(wg0, wg1), wg2);
  if (have) {
    const float qv = q_substitute(c, t, ch, Pp, Ep);
    if (l.r == 0 && l.g8 < n) sm.q[l.g8] = qv;
#pragma unroll 1
    for (int base = SLOTP; base + wfirst < n; base += SLOTP) {  // further passes (warp-uniform trip count)
      const int k = base + l.g8;
      ch = ev.cand[k];
      c = q_loads(sm, ch, rc);
      q_rows(c, qc, l.rv, is_end, n4b_r, n2b_r, Pp, Ep);
#pragma unroll
      for (int s = 1; s < 8; s <<= 1) {
        Pp = __fadd_rn(Pp, __shfl_xor_sync(0xffffffffu, Pp, s));
        Ep = __fadd_rn(Ep, __shfl_xor_sync(0xffffffffu, Ep, s));
      }
      const float q2 = q_substitute(c, t, ch, Pp, Ep);
      if (l.r == 0 && k < n) sm.q[k] = q2;
    }
  }
  return t;
}

// ---------------------------------------------------------------------------
// Hand-off look-ahead in the FAST order (extension: the reference's TODO, README.md:71-72; specified at
// dca_config.hoff_lookahead in include/dca_b200.h, oracle: orc_get_action_lookahead).  On the END half of a hand-off
// (END ch (Just toCell)) the value that SELECTS the channel to free is
//   qla[k] = max over h eligible at toCell on afterstate_k of V(afterstate_k with h assigned at toCell),
// V(afterstate_k) when nothing is eligible there; a left fold with `>` over ascending h, as the oracle does it.  Every
// V is again exactly the dense dot of the materialised state in the FAST tree: up to three planes change -- the freed
// channel's (-1 on N4(cell) \ {cell}, and +1 on N4(toCell) \ {toCell} when h is that same channel), h's (+1 on
// N4(toCell) \ {toCell}) and n_elig (+1 where k becomes eligible, then -1 where h was eligible on that afterstate) -- their
// rows are re-evaluated and their plane sums substituted into the V tree, whose two changed leaves meet at the level of
// their highest differing index bit.  sm.q keeps the plain values V(afterstate_k): the TD target (Agent.hs:63) is the real
// afterstate's.  Candidate k sits on lane group k % 12 as in q_phase; all trip counts are warp-uniform.
// ---------------------------------------------------------------------------
__device__ __noinline__ void q_lookahead(SmemF& sm, const Lane& l, const Ev& ev, const LaRec& rec, const int n,
                                         const uint2 n4b_r, const uint2 n2b_r, const VTree& vt) {
  const int rc = l.rv ? l.r : 6;
  const int wfirst = l.vw * 4;
  const int to = ev.hoff;
  const int nh = rec.nh;
  const uint32_t hm0 = rec.hm[0], hm1 = rec.hm[1], hm2 = rec.hm[2];
  uint2 t4, t2;  // row rc of N4(toCell) \ {toCell} and of N2(toCell) u {toCell} as byte masks
  {
    const uint2 a = sm.n4rows[to], b = sm.n2rows[to];
    const int sh = 8 * (rc & 3);
    t4 = spread7(((rc < 4 ? a.x : a.y) >> sh) & 0x7Fu);
    t2 = spread7(((rc < 4 ? b.x : b.y) >> sh) & 0x7Fu);
  }
  const int pos_to = to + (to * 37 >> 8);
  const QConst qc = q_const(sm, rc);
  __syncwarp();  // (sm.q[k] of this group's candidates was written by its row-0 lane in q_phase)
#pragma unroll 1
  for (int base = 0; base + wfirst < n; base += SLOTP) {  // warp-uniform trip count
    const int k = base + l.g8;
    const int ch = ev.cand[k];  // (k >= n: a stale but valid channel number; masked at the store)
    const CandLoads ck = q_loads(sm, ch, rc);
    const int la = ch & 31, hia = ch >> 5;
    float sibA[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) sibA[i] = __shfl_sync(0xffffffffu, vt.sib[i], la);
    // freeing ch makes ch itself eligible at the target iff the departing call was its only user within distance 2 of it
    const bool extra_ok = sm.cnt2[ch * PLANE + pos_to] == 1;
    const uint32_t lt = (1u << la) - 1u;
    const int pos = hia == 0 ? __popc(hm0 & lt) : (hia == 1 ? __popc(hm0) + __popc(hm1 & lt) : __popc(hm0) + __popc(hm1) + __popc(hm2 & lt));
    // the END of ch at the event cell (Gridfuncs.hs:212-252 with diff = -1)
    const uint2 xk1 = bsub(ck.xr, n4b_r);
    const uint2 e1 = badd(qc.x70, make_uint2(n2b_r.x & eq_bytes(ck.cr.x, 1u), n2b_r.y & eq_bytes(ck.cr.y, 1u)));
    const uint2 crk1 = bsub(ck.cr, n2b_r);  // cnt2[ch] on that afterstate
    float best = 0.0f;
    bool have = false;
#pragma unroll 1
    for (int u = 0; u <= nh; ++u) {  // the merged ascending list: eligible channels, ch at its place (slot `pos`)
      const bool same = u == pos;
      const int h = same ? ch : rec.hcand[u < pos ? u : u - 1];
      const bool active = same ? extra_ok : true;
      const CandLoads chh = q_loads(sm, h, rc);
      const int lb = h & 31, hib = h >> 5;
      // the HOFF arrival taking h at the target on that afterstate (diff = +1)
      const uint2 xk = same ? badd(xk1, t4) : xk1;
      const uint2 xh = badd(chh.xr, t4);
      const uint2 crh = same ? crk1 : chh.cr;
      const uint2 e2 = bsub(e1, make_uint2(t2.x & eq_bytes(crh.x, 0u), t2.y & eq_bytes(crh.y, 0u)));
      float Pk = rowdot2(cvt_row<DCA_CVT_MIX_EVAL>(xk), make_float2(ck.wa.x, ck.wa.y), make_float2(ck.wa.z, ck.wa.w),
                         make_float2(ck.wb.x, ck.wb.y), ck.wb.z);
      float Ph = rowdot2(cvt_row<DCA_CVT_MIX_EVAL>(xh), make_float2(chh.wa.x, chh.wa.y), make_float2(chh.wa.z, chh.wa.w),
                         make_float2(chh.wb.x, chh.wb.y), chh.wb.z);
      float Ep = rowdot2(cvt_row<DCA_CVT_MIX_EVAL>(e2), make_float2(qc.e_a.x, qc.e_a.y), make_float2(qc.e_a.z, qc.e_a.w),
                         make_float2(qc.e_b.x, qc.e_b.y), qc.e_b.z);
      if (!l.rv) { Pk = 0.0f; Ph = 0.0f; Ep = 0.0f; }
#pragma unroll
      for (int sft = 1; sft < 8; sft <<= 1) {
        Pk = __fadd_rn(Pk, __shfl_xor_sync(0xffffffffu, Pk, sft));
        Ph = __fadd_rn(Ph, __shfl_xor_sync(0xffffffffu, Ph, sft));
        Ep = __fadd_rn(Ep, __shfl_xor_sync(0xffffffffu, Ep, sft));
      }
      // the two leaves of the V tree (leaf l = (P[l] + P[l + 32]) + P[l + 64]); h in ch's leaf but another plane: one leaf
      const bool one_leaf = same || lb == la;
      float a0 = hia == 0 ? Pk : ck.l0, a1 = hia == 1 ? Pk : ck.l1, a2 = hia == 2 ? Pk : ck.l2;
      if (!same && lb == la) {
        if (hib == 0) a0 = Ph;
        else if (hib == 1) a1 = Ph;
        else a2 = Ph;
      }
      const float b0 = hib == 0 ? Ph : chh.l0, b1 = hib == 1 ? Ph : chh.l1, b2 = hib == 2 ? Ph : chh.l2;
      float A = __fadd_rn(__fadd_rn(a0, a1), a2), B = __fadd_rn(__fadd_rn(b0, b1), b2);
      const int merge = one_leaf ? -1 : 31 - __clz(la ^ lb);  // the butterfly level at which leaf lb's block joins leaf la's path
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const float sb = __shfl_sync(0xffffffffu, vt.sib[i], lb);
        A = __fadd_rn(A, i == merge ? B : sibA[i]);
        B = __fadd_rn(B, sb);
      }
      const float v = __fadd_rn(A, Ep);
      if (active) {
        if (!have || v > best) best = v;
        have = true;
      }
    }
    if (!have) best = sm.q[k];  // nothing is eligible at the target: the hand-off will be rejected
    if (l.r == 0 && k < n) la_of(sm).qla[k] = best;
  }
}
// index of channel `ch` in the ascending candidate list (warp-wide)
__device__ __forceinline__ int cand_index(const uint8_t* cand, const int n, const int ch) {
  int c = 0;
  for (int k = lane_id(); k < n; k += 32) c += cand[k] < ch ? 1 : 0;
  return __reduce_add_sync(0xffffffffu, c);
}

// argpmax (AccUtils.hs:203-208): the largest value, the LAST index among equals;
// NaNs follow the sequential right fold exactly.  Every warp computes it for itself.
// cand_lane = cand[lane] (read before the barrier): the chosen channel rides along in the index
// reduction, and the best value is rebuilt from the reduced key, so that nothing is loaded or
// shuffled after the two reductions.  The common case (n <= 32, no NaN) is branch-free; a NaN is
// mapped to the largest key, so one warp-uniform test after the first reduction finds every
// uncommon case.  Returns the channel; qbest = q[k].  (n == 0: the result is meaningless.)
// (TAG: the look-ahead instantiation calls its own copy -- ptxas allocates an out-of-line function's registers across all of its
//  call sites, and one more caller changes the hot kernels' code)
template <int TAG = 0>
__device__ __noinline__ int argmax_uncommon(const float* q, const uint8_t* cand, int n, float& qbest) {
  const int lane = lane_id();
  bool nan = false;
  float bq = -CUDART_INF_F;
  int bk = -1;
#pragma unroll 1
  for (int k = lane; k < n; k += 32) {
    const float v = q[k];
    nan |= v != v;
    if (v >= bq) { bq = v; bk = k; }
  }
  int best;
  if (__any_sync(0xffffffffu, nan)) {  // the sequential right fold, NaNs included
    best = n - 1;
#pragma unroll 1
    for (int i = n - 2; i >= 0; --i)
      if (q[i] > q[best]) best = i;
  } else {
    uint32_t u = __float_as_uint(bq);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    if (bq == 0.0f) u = 0x80000000u;
    if (bk < 0) u = 0u;
    const uint32_t m = __reduce_max_sync(0xffffffffu, u);
    best = __reduce_max_sync(0xffffffffu, (u == m) ? bk : -1);
  }
  qbest = q[best];
  return cand[best];
}
template <int TAG = 0>
__device__ __forceinline__ int argmax_last_warp(const float* q, const uint8_t* cand, const int cand_lane, const int n,
                                                float& qbest) {
  const int lane = lane_id();
  const float bq = q[lane];  // (lanes >= n read a stale value and are masked)
  // order-preserving map float -> uint32 (shifts and xors: no predicate latency), then two warp reductions
  const uint32_t bits = __float_as_uint(__fadd_rn(bq, 0.0f));                  // -0 -> +0
  uint32_t u = bits ^ ((uint32_t)((int32_t)bits >> 31) | 0x80000000u);         // negative: ~bits, else bits | sign
  // (a NaN flags the uncommon path by itself: the FADD above returns the canonical NaN 0x7FFFFFFF for any NaN input, which
  //  this map sends to 0xFFFFFFFF, above +inf)
  if (lane >= n) u = 0u;
  const uint32_t m = __reduce_max_sync(0xffffffffu, u);
  const int key = __reduce_max_sync(0xffffffffu, (u == m) ? ((lane << 8) | cand_lane) : -1);
  if (n > 32 || m == 0xFFFFFFFFu) return argmax_uncommon<TAG>(q, cand, n, qbest);  // warp-uniform
  // the inverse map gives q[k] bit for bit, except that a best value of -0 comes back as +0: it is only
  // ever added to reward - avgR, which is never -0, so the sum is the same
  qbest = __uint_as_float(m ^ ~((uint32_t)((int32_t)m >> 31) & 0x7FFFFFFFu));
  return key & 0xFF;
}

// what every thread derives from the Q phase: the action and the TD scalars.  Everything that does not
// depend on the q values is prepared BEFORE the barrier that publishes them (decide_prepare); after it
// only argmax, three dependent flops to c * td, and the learning-rate products remain.
struct Decision {
  int act;
  float tderr;
  TdScalars td;
};
struct DecidePrep {
  int cand_lane;  // cand[lane]
  int ncalls1;    // calls in progress after this event if a channel is chosen (gridStep, Simulator.hs:137-144)
  float base;     // reward - avgR for that reward                       (Agent.hs:65)
  float c;        // -2 alphaNet                                         (Agent.hs:67)
  float cavg, ncdot;
  float alpha_avg, alpha_grad;
};
// (called before the candidate evaluation, so that its loads are in flight early; ncdot follows when
//  x.wGradCorr is known)
__device__ __forceinline__ DecidePrep decide_prepare(const SmemF& sm, const Ev& ev, const int n, const bool is_end,
                                                     const float avgR, const int ncalls) {
  DecidePrep p;
  p.cand_lane = ev.cand[lane_id()];
  p.ncalls1 = n > 0 ? ncalls + (is_end ? -1 : 1) : ncalls;
  const float reward = (float)p.ncalls1;  // boolSum nextGrid
  p.base = __fsub_rn(reward, avgR);
  p.c = __fmul_rn(-2.0f, sm.alpha_net);
  p.cavg = __fmul_rn(p.c, avgR);
  p.ncdot = 0.0f;
  p.alpha_avg = sm.alpha_avg;
  p.alpha_grad = sm.alpha_grad;
  return p;
}
__device__ __forceinline__ Decision decide_from(const Ev& ev, const VTree& vt, const DecidePrep& p, const int n,
                                                const int act, const float qbest, float& avgR, int& ncalls) {
  Decision d;
  d.act = n > 0 ? act : -1;
  const float next_val = n > 0 ? qbest : vt.val;  // no action: nextFrep = frep (Simulator.hs:170-171)
  ncalls = p.ncalls1;
  // backward scalars (Agent.hs:62-67,75,78)
  d.tderr = __fsub_rn(__fadd_rn(p.base, next_val), vt.val);
  d.td.ctd = __fmul_rn(p.c, d.tderr);
  d.td.cavg = p.cavg;
  d.td.ncdot = p.ncdot;
  d.td.sg = __fmul_rn(p.alpha_grad, __fsub_rn(d.tderr, vt.dotg));
  avgR = __fadd_rn(avgR, __fmul_rn(p.alpha_avg, d.tderr));
  return d;
}
__device__ __forceinline__ Decision decide(const SmemF& sm, const Ev& ev, const VTree& vt, const DecidePrep& p,
                                           const int n, float& avgR, int& ncalls) {
  float qbest;
  const int act = argmax_last_warp(sm.q, ev.cand, p.cand_lane, n, qbest);  // selectAction, Simulator.hs:202
  return decide_from(ev, vt, p, n, act, qbest, avgR, ncalls);
}
// the END half of a hand-off under hoff_lookahead: the look-ahead values select, the plain value of the chosen afterstate
// is the TD target
__device__ __forceinline__ Decision decide_la(const SmemF& sm, const Ev& ev, const VTree& vt, const DecidePrep& p,
                                              const int n, float& avgR, int& ncalls) {
  float qbest;
  const int act = argmax_last_warp<1>(la_of(sm).qla, ev.cand, p.cand_lane, n, qbest);
  if (n > 0) qbest = sm.q[cand_index(ev.cand, n, act)];
  return decide_from(ev, vt, p, n, act, qbest, avgR, ncalls);
}

// ---------------------------------------------------------------------------
// Event side (warp 3).  All registers are warp-uniform unless noted.
// ---------------------------------------------------------------------------
// (The event warp's chain is the most expensive place for an instruction -- one more per event costs ~0.07 % of the
//  kernel's throughput, twice what it costs on an agent warp -- so everything it counts per event is a 32-bit delta
//  of this unit, folded into the 64-bit values of the replica blob when the unit ends.)
struct EvRegs {
  unsigned n_end, next_id;
  int left_success;                     // events until iter == nEvents (clamped to int: a unit is far shorter)
  int left_50;                          // events until iter > 50 can hold (the ZeroLoss rule, SimRunner.hs:161)
  float loss_sum, min_loss;
  double mean_new, call_dur, call_dur_hoff;
  unsigned long long hoff_thresh;       // p < hoffProb  <=>  (word >> 11) < ceil(hoffProb * 2^53)   (p = (word >> 11) * 2^-53 exactly)
  unsigned stats[8];                    // Stats counters (src/Stats.hs:15-30): increments of this unit
  int dirty1;                           // a second queue block whose minimum must be re-scanned after a pop (-1 = none; rare)
  unsigned long long dbase;             // cached draws: lane l holds draw dbase + l ...
  unsigned dpos;                        // ... draws dbase .. dbase + dpos - 1 are consumed (the replica's draw count is dbase + dpos)
  unsigned long long dword;             // ... its random word (per lane)
  double dnl;                           // ... and -log(u) of that word (per lane)
  // random source
  int rng_mode;
  uint32_t k0, k1;
  unsigned long long stream;
  const unsigned long long* tape_w;
  const double* tape_l;
  unsigned long long tape_len;
};

__device__ __noinline__ unsigned long long ev_word_cold(int rng_mode, uint32_t k0, uint32_t k1,
                                                         unsigned long long stream, const unsigned long long* tape_w,
                                                         unsigned long long tape_len, unsigned long long d) {
  if (rng_mode == RNG_PHILOX) return philox_word(d, stream, k0, k1);
  if (!tape_len) return 0ULL;
  return tape_w[d >= tape_len ? tape_len - 1 : d];
}
__device__ __forceinline__ unsigned long long ev_word(const EvRegs& e, unsigned long long d) {
  return ev_word_cold(e.rng_mode, e.k0, e.k1, e.stream, e.tape_w, e.tape_len, d);
}
// refill the draw cache: lane l <- draw dbase + l (cold: one refill serves ~19 events).  Returns by
// value: taking the address of EvRegs members would force the whole struct into local memory.
struct Draw {
  unsigned long long word;
  double nl;
};
__device__ __noinline__ Draw ev_refill(int rng_mode, uint32_t k0, uint32_t k1, unsigned long long stream,
                                       const unsigned long long* tape_w, const double* tape_l,
                                       unsigned long long tape_len, unsigned long long dbase) {
  const unsigned long long d = dbase + (unsigned long long)lane_id();
  Draw r;
  if (rng_mode == RNG_PHILOX) {
    r.word = philox_word(d, stream, k0, k1);
    r.nl = neglog_u53(r.word);
  } else if (tape_len) {
    const unsigned long long i = d >= tape_len ? tape_len - 1 : d;
    r.word = tape_w[i];
    r.nl = tape_l[i];
  } else {
    r.word = 0ULL;
    r.nl = 0.0;
  }
  return r;
}
// make sure the next four draws are cached
__device__ __forceinline__ void ev_draws(EvRegs& e) {
  if (e.dpos + 4 <= 32) return;  // warp-uniform
  e.dbase += e.dpos;
  e.dpos = 0;
  const Draw r = ev_refill(e.rng_mode, e.k0, e.k1, e.stream, e.tape_w, e.tape_l, e.tape_len, e.dbase);
  e.dword = r.word;
  e.dnl = r.nl;
}
// the k-th draw from now: its word / its -log(u)   (k < 4 is always cached after ev_draws)
__device__ __forceinline__ unsigned long long ev_draw_word(const EvRegs& e, unsigned k) {
  const unsigned i = e.dpos + k;
  const unsigned long long w = __shfl_sync(0xffffffffu, e.dword, (int)(i & 31));
  return i < 32 ? w : ev_word(e, e.dbase + i);
}
__device__ __forceinline__ double ev_draw_nl(const EvRegs& e, unsigned k) {
  return __shfl_sync(0xffffffffu, e.dnl, (int)((e.dpos + k) & 31));
}

// candidate channels of event `ev` (getAction, Simulator.hs:161-165), ascending
// (indicesOf, AccUtils.hs:233-240), and the row byte masks of its cell.  Runs while
// the agents move cnt2[act] to the afterstate, so channel `act` is judged from the
// grid (which this warp owns and has already updated) instead of cnt2.
__device__ __forceinline__ void ev_candidates(const SmemF& sm, Ev& ev, const int act) {
  const int lane = lane_id();
  const int cell = ev.cell;
  const bool is_end = ev.type == EV_END;
  const int pos = pos_of(cell);
  const unsigned long long n4 = c_tab.n4excl[cell], n2 = c_tab.n2incl[cell];
  uint32_t m[3];
  if (is_end) {  // inuseChs, Gridfuncs.hs:86-87
    m[0] = sm.grid[cell][0]; m[1] = sm.grid[cell][1]; m[2] = sm.grid[cell][2];
  } else {       // eligibleChs, Gridfuncs.hs:69-82  <=>  nobody within distance 2 (incl. the cell) uses ch
    bool act_free = true;
    if (act >= 0) {
      const int c1 = lane + 32;
      const bool u0 = ((n2 >> lane) & 1ULL) && ((sm.grid[lane][act >> 5] >> (act & 31)) & 1u);
      const bool u1 = c1 < NCELL && ((n2 >> c1) & 1ULL) && ((sm.grid[c1][act >> 5] >> (act & 31)) & 1u);
      act_free = !__any_sync(0xffffffffu, u0 || u1);
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int ch = 32 * j + lane;
      bool el = false;
      if (ch < NCH) el = ch == act ? act_free : sm.cnt2[ch * PLANE + pos] == 0;
      m[j] = __ballot_sync(0xffffffffu, el);
    }
  }
  int base = 0;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if ((m[j] >> lane) & 1u) ev.cand[base + __popc(m[j] & ((1u << lane) - 1u))] = (uint8_t)(32 * j + lane);
    base += __popc(m[j]);
  }
  if (lane < 8) {
    ev.n4b[lane] = (lane < 7) ? spread7((uint32_t)(n4 >> (7 * lane)) & 0x7Fu) : make_uint2(0, 0);
    ev.n2b[lane] = (lane < 7) ? spread7((uint32_t)(n2 >> (7 * lane)) & 0x7Fu) : make_uint2(0, 0);
  }
  if (lane == 0) ev.ncand = base;
  __syncwarp();
}

// The run kernel splits that work: the event warp pops the next event while the agents still evaluate the current
// one's candidates (the pop does not depend on the chosen channel) and publishes its kind, cell and row masks
// (ev_masks); after the decision it sets the grid bit, waits at named barrier 1 until agent warp 2 has moved cnt2 to the
// afterstate, and builds the candidate list (agent_candidates) -- grid and cnt2 are in their afterstate then, so no
// channel needs a special case.
// the row byte masks of the next event's cell, from the shared tables (part of the record the event warp publishes)
__device__ __forceinline__ void ev_masks(const SmemF& sm, Ev& ev, const int cell) {
  const int lane = lane_id();
  if (lane < 8) {
    const uint2 a = sm.n4rows[cell], b = sm.n2rows[cell];
    const int sh = 8 * (lane & 3);
    const uint32_t r4 = ((lane < 4 ? a.x : a.y) >> sh) & 0x7Fu, r2 = ((lane < 4 ? b.x : b.y) >> sh) & 0x7Fu;  // (row 7: 0)
    ev.n4b[lane] = spread7(r4);
    ev.n2b[lane] = spread7(r2);
  }
}
__device__ __forceinline__ int agent_candidates(const SmemF& sm, Ev& ev, const int type, const int cell) {  // returns their number
  const int lane = lane_id();
  uint32_t m[3];
  if (type == EV_END) {     // inuseChs, Gridfuncs.hs:86-87
    const uint4 g = *reinterpret_cast<const uint4*>(&sm.grid[cell][0]);
    m[0] = g.x; m[1] = g.y; m[2] = g.z;
  } else {                  // eligibleChs, Gridfuncs.hs:69-82  <=>  nobody within distance 2 (incl. the cell) uses ch
    const int pos = cell + (cell * 37 >> 8);  // (cell / 7) * 8 + cell % 7 for cell < 49
    const uint8_t c0 = sm.cnt2[lane * PLANE + pos], c1 = sm.cnt2[(lane + 32) * PLANE + pos];
    const uint8_t c2 = sm.cnt2[(lane < 6 ? lane + 64 : 0) * PLANE + pos];
    m[0] = __ballot_sync(0xffffffffu, c0 == 0);
    m[1] = __ballot_sync(0xffffffffu, c1 == 0);
    m[2] = __ballot_sync(0xffffffffu, lane < 6 && c2 == 0);
  }
  int base = 0;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if ((m[j] >> lane) & 1u) ev.cand[base + __popc(m[j] & ((1u << lane) - 1u))] = (uint8_t)(32 * j + lane);
    base += __popc(m[j]);
  }
  if (lane == 0) ev.ncand = base;
  return base;
}

// hand-off look-ahead: the channels eligible at the hand-off target `to` (cnt2 == 0 there) on the current grid
__device__ __forceinline__ void la_candidates(const SmemF& sm, LaRec& r, const int to) {
  const int lane = lane_id();
  const int pos = to + (to * 37 >> 8);
  const uint8_t c0 = sm.cnt2[lane * PLANE + pos], c1 = sm.cnt2[(lane + 32) * PLANE + pos];
  const uint8_t c2 = sm.cnt2[(lane < 6 ? lane + 64 : 0) * PLANE + pos];
  uint32_t m[3];
  m[0] = __ballot_sync(0xffffffffu, c0 == 0);
  m[1] = __ballot_sync(0xffffffffu, c1 == 0);
  m[2] = __ballot_sync(0xffffffffu, lane < 6 && c2 == 0);
  int base = 0;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if ((m[j] >> lane) & 1u) r.hcand[base + __popc(m[j] & ((1u << lane) - 1u))] = (uint8_t)(32 * j + lane);
    base += __popc(m[j]);
  }
  if (lane == 0) { r.nh = base; r.hm[0] = m[0]; r.hm[1] = m[1]; r.hm[2] = m[2]; }
}

// ---- the event queue (EventGen, src/EventGen/Internal.hs:44-53) as a dense slot table with
// per-block minima: pop = argmin over (time, id) (EventGen.hs:39-50) costs one 22-wide reduction
// plus the re-scan of the one or two 32-slot blocks it touched.
// lexicographic warp argmin over (t, id): the lane that holds the minimum (the first one among equals).  The upper
// 32 bits of the time almost always decide (event times are spread over minutes, the upper word resolves ~1e-6 of
// one): one reduction and a ballot; the lower word and the id follow only when two lanes tie (warp-uniform).
__device__ __forceinline__ unsigned warp_argmin_lane(unsigned long long t, unsigned id) {
  const unsigned hi = (unsigned)(t >> 32), lo = (unsigned)t;
  const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
  bool ok = hi == mhi;
  unsigned eq = __ballot_sync(0xffffffffu, ok);
  if (eq & (eq - 1)) {
    const unsigned mlo = __reduce_min_sync(0xffffffffu, ok ? lo : 0xFFFFFFFFu);
    ok = ok && lo == mlo;
    eq = __ballot_sync(0xffffffffu, ok);
    if (eq & (eq - 1)) {
      const unsigned mid = __reduce_min_sync(0xffffffffu, ok ? id : 0xFFFFFFFFu);
      eq = __ballot_sync(0xffffffffu, ok && id == mid);
    }
  }
  return __ffs(eq) - 1;
}
// The minimum (time, id) of every block of 32 slots and where it is: lane b of the event warp keeps block b's in
// REGISTERS (lanes >= NBLK: +inf).  The event warp's chain is bound by shared-memory round trips -- under the agents'
// load one costs ~100 cycles -- so the per-block minima never go through shared memory.
struct QBlk {
  unsigned long long t;
  unsigned id, s;
};
__device__ __forceinline__ void q_block_rescan(const SmemF& sm, QBlk& qb, const unsigned b) {
  const unsigned i = 32 * b + lane_id();
  const unsigned long long t = (unsigned long long)__double_as_longlong(sm.q_time[i]);
  const unsigned id = sm.q_id[i];
  const unsigned w = warp_argmin_lane(t, id);
  const unsigned long long tw = __shfl_sync(0xffffffffu, t, (int)w);
  const unsigned idw = __shfl_sync(0xffffffffu, id, (int)w);
  if ((unsigned)lane_id() == b) { qb.t = tw; qb.id = idw; qb.s = 32 * b + w; }
}
// a slot of block b was just written with (t, id): lower the block's minimum if needed (lane-local)
__device__ __forceinline__ void q_block_lower(QBlk& qb, const unsigned b, const unsigned slot, const unsigned long long t,
                                              const unsigned id) {
  if ((unsigned)lane_id() == b && (t < qb.t || (t == qb.t && id < qb.id))) { qb.t = t; qb.id = id; qb.s = slot; }
}
// slot of the earliest stored event
__device__ __forceinline__ unsigned q_min_slot(const QBlk& qb) {
  const unsigned w = warp_argmin_lane(qb.t, qb.id);
  return __shfl_sync(0xffffffffu, qb.s, (int)w);
}
// everything derived from the slot table (run once per launch by the event warp)
__device__ __forceinline__ void q_build(SmemF& sm, QBlk& qb, unsigned n_end) {
  const int lane = lane_id();
  for (unsigned i = QNEW + n_end + lane; i < QCAPS; i += 32) {  // unused slots never win
    sm.q_time[i] = CUDART_INF;
    sm.q_id[i] = 0xFFFFFFFFu;
  }
#ifdef DCA_END_SLOT_INDEX
  for (unsigned i = QNEW + lane; i < QNEW + n_end; i += 32) {
    const unsigned key = sm.q_key[i];
    sm.end_slot[key >> 7][key & 127u] = (uint16_t)i;
  }
#endif
  __syncwarp();
  qb.t = 0xFFFFFFFFFFFFFFFFULL; qb.id = 0xFFFFFFFFu; qb.s = 0;
#pragma unroll 1
  for (unsigned b = 0; b < NBLK; ++b) q_block_rescan(sm, qb, b);
  __syncwarp();
}

// environmentStep (Simulator.hs:101-134) in three parts: the new events (ev_env_push), the pop of
// the next event (ev_env_pop) -- neither depends on WHICH channel was chosen, only on whether the
// call was accepted -- and the channel-dependent rest (ev_env_post: END key, reassign, grid bit).
constexpr int CH_PENDING = 127;  // channel of the END event pushed by the current event: not chosen yet
constexpr int PEND_OVERFLOW = -2;  // ev_env_push: the END event did not fit into the slot table
struct EnvNext {                 // the next event (registers, warp-uniform)
  double time;
  int type, cell, ch, hoff;
  int pend_slot;                 // slot of the END event whose key still lacks the channel; -1 = none
};

// part 1: statistics and the new events (EventGen.hs:75-116)
__device__ __forceinline__ int ev_env_push(SmemF& sm, EvRegs& e, QBlk& qb, const int type, const int cell, const double time,
                                           const bool accepted) {
  const int lane = lane_id();
  int pend_slot = -1;
  // Stats hooks, Stats.hs:61-81 as called from Simulator.hs:105-114
  if (type == EV_NEW) {
    e.stats[6]++; e.stats[0]++;
    if (accepted) e.stats[1]++;
    else { e.stats[2]++; e.stats[7]++; }
  } else if (type == EV_HOFF) {
    e.stats[3]++;
    if (!accepted) { e.stats[2]++; e.stats[7]++; }  // Simulator.hs:113
  } else {
    e.stats[4]++;
  }
  bool have_end = false, have_new = false;
  unsigned long long t_end = 0, t_new = 0;
  unsigned id_end = 0, id_new = 0;
  int end_hoff = HOFF_NONE;
  if (type == EV_NEW) {
    // generateNewEvent, EventGen.hs:75-85: always, at time + Exp(mean_new)
    t_new = (unsigned long long)__double_as_longlong(__dadd_rn(time, __dmul_rn(ev_draw_nl(e, 0), e.mean_new)));
    id_new = e.next_id++;
    have_new = true;
    unsigned used = 1;
    if (accepted) {
      const unsigned long long w1 = ev_draw_word(e, 1);  // p = stdUniform, Simulator.hs:121
      const double te = __dadd_rn(time, __dmul_rn(ev_draw_nl(e, 2), e.call_dur));  // callDurNew for both kinds, EventGen.hs:102,112
      used = 3;
      if ((w1 >> 11) < e.hoff_thresh) {  // p < hoffProb: generateHoffNewEvent, EventGen.hs:94-109
        const int m = c_tab.n2cnt[cell];
        const int n_reject = 256 % m;
        int z;
        for (;;) {  // random-fu integralUniform on the low byte, with rejection
          const unsigned long long wz = ev_draw_word(e, used);
          ++used;
          z = (int)(wz & 0xFFULL);
          if (z >= n_reject) break;
          if (e.rng_mode == RNG_TAPE && e.dbase + e.dpos + used >= e.tape_len) break;  // exhausted tape: stop is flagged
        }
        end_hoff = c_tab.n2list[cell][z % m];
      }
      t_end = (unsigned long long)__double_as_longlong(te);
      id_end = e.next_id++;
      if (end_hoff != HOFF_NONE) e.next_id++;  // the HOFF event's id (pushed right after its END, never stored)
      have_end = true;
    }
    e.dpos += used;
  } else if (type == EV_HOFF) {
    if (accepted) {  // generateHoffEndEvent, EventGen.hs:114-116
      const double te = __dadd_rn(time, __dmul_rn(ev_draw_nl(e, 0), e.call_dur_hoff));
      t_end = (unsigned long long)__double_as_longlong(te);
      id_end = e.next_id++;
      have_end = true;
      e.dpos += 1;
    }
  }
  // the queue writes: NEW at slot `cell`, END (cell, <channel pending>) [hoff] at slot QNEW + n_end -- by lane 0 (the only
  // thread that ever writes the slot table); the minima of the blocks they fall into are lowered in their lanes' registers
  const unsigned slot_end = QNEW + e.n_end;
  if (have_end) {
    if (slot_end < QCAP) {
      pend_slot = (int)slot_end;
      e.n_end++;
    } else {
      // the table is full (more than QEND calls in progress: impossible while the reuse constraint holds, reachable
      // only from a grid forced in through dca_write_state): nothing is written, n_end stays in bounds, and the
      // caller stops the replica with ST_QUEUE_OVERFLOW
      pend_slot = PEND_OVERFLOW;
      have_end = false;
    }
  }
  if (have_new) {
    if (lane == 0) {
      sm.q_time[cell] = __longlong_as_double((long long)t_new);
      sm.q_id[cell] = id_new;
    }
    q_block_lower(qb, (unsigned)cell >> 5, (unsigned)cell, t_new, id_new);
  }
  if (have_end) {
    if (lane == 0) {
      sm.q_time[slot_end] = __longlong_as_double((long long)t_end);
      sm.q_id[slot_end] = id_end;
      sm.q_key[slot_end] = (uint16_t)((cell << 7) | CH_PENDING);
      sm.q_hoff[slot_end] = (uint8_t)end_hoff;
    }
    q_block_lower(qb, slot_end >> 5, slot_end, t_end, id_end);
  }
  __syncwarp();
  return pend_slot;
}

// part 2: the next event (Simulator.hs:132-134).  The pop and the re-scan of the block it empties are ONE shared-memory
// round trip: the popped entry, its whole block and (up front, while the argmin runs) the last END entry -- which moves
// into the hole to keep the table dense -- are loaded together, the lanes whose slots change substitute the new values
// in registers, and the block's new minimum is reduced from those.
__device__ __forceinline__ EnvNext ev_env_pop(SmemF& sm, EvRegs& e, QBlk& qb, const int type, const int ev_hoff,
                                              const double time, const int pend_slot) {
  const int lane = lane_id();
  EnvNext nx;
  nx.pend_slot = pend_slot;
  nx.ch = -1;
  nx.hoff = HOFF_NONE;
  if (type == EV_END && ev_hoff != HOFF_NONE) {
    // the HOFF that follows a hand-off END: same time, next id (EventGen.hs:107-108)
    nx.time = time; nx.type = EV_HOFF; nx.cell = ev_hoff;
  } else {  // pop, EventGen.hs:39-50
    const unsigned last = QNEW + e.n_end - 1;  // (no END event stored: slot QNEW - 1, a harmless load)
    const unsigned long long lt = (unsigned long long)__double_as_longlong(sm.q_time[last]);
    const unsigned lid = sm.q_id[last], lkey = sm.q_key[last], lhoff = sm.q_hoff[last];
    const unsigned slot = q_min_slot(qb);
    const unsigned b = slot >> 5, i = 32 * b + (unsigned)lane;
    nx.time = sm.q_time[slot];
    const int key = sm.q_key[slot];    // (slot is a live one: < QNEW + n_end <= QCAP)
    const int hoff = sm.q_hoff[slot];
    unsigned long long bt = (unsigned long long)__double_as_longlong(sm.q_time[i]);
    unsigned bid = sm.q_id[i];
    if (slot < QNEW) {
      nx.type = EV_NEW; nx.cell = (int)slot;
      if (lane == 0) { sm.q_time[slot] = CUDART_INF; sm.q_id[slot] = 0xFFFFFFFFu; }
      if (i == slot) { bt = 0x7FF0000000000000ULL; bid = 0xFFFFFFFFu; }
    } else {
      nx.type = EV_END; nx.cell = key >> 7; nx.ch = key & 127; nx.hoff = hoff;
      const bool last_was_min = __shfl_sync(0xffffffffu, qb.s, (int)(last >> 5)) == last;
      if (lane == 0) {
        if (slot != last) {  // keep the table dense: the last END event moves into the hole
          sm.q_time[slot] = __longlong_as_double((long long)lt);
          sm.q_id[slot] = lid;
          sm.q_key[slot] = (uint16_t)lkey;
          sm.q_hoff[slot] = (uint8_t)lhoff;
#ifdef DCA_END_SLOT_INDEX
          if ((lkey & 127u) != CH_PENDING) sm.end_slot[lkey >> 7][lkey & 127u] = (uint16_t)slot;
#endif
        }
        sm.q_time[last] = CUDART_INF;
        sm.q_id[last] = 0xFFFFFFFFu;
      }
      if (i == slot) { bt = lt; bid = lid; }                                  // the hole takes the last entry ...
      if (i == last) { bt = 0x7FF0000000000000ULL; bid = 0xFFFFFFFFu; }      // ... whose own slot is empty now
      if ((int)slot == nx.pend_slot) nx.pend_slot = -1;         // the END just pushed is the next event: not stored
      else if ((int)last == nx.pend_slot) nx.pend_slot = (int)slot;  // ... or it moved into the hole
      e.n_end--;
      if ((last >> 5) != b && last_was_min) e.dirty1 = (int)(last >> 5);
    }
    // the new minimum of the block the pop touched
    const unsigned w = warp_argmin_lane(bt, bid);
    const unsigned long long tw = __shfl_sync(0xffffffffu, bt, (int)w);
    const unsigned idw = __shfl_sync(0xffffffffu, bid, (int)w);
    if ((unsigned)lane == b) { qb.t = tw; qb.id = idw; qb.s = 32 * b + w; }
  }
  __syncwarp();
  return nx;
}

// slot of the END event with key `key` among the n_end stored ones (reassign, EventGen.hs:62-72): warp-wide search of
// q_key, eight keys per 16-byte load.  Slots past the live range may hold stale copies of live keys: only indices
// below QNEW + n_end count; a live key is unique.  (NEW slots carry channel 127 and never match.)
__device__ __forceinline__ unsigned q_find_end(const SmemF& sm, const unsigned key, const unsigned n_end) {
  const unsigned lane = (unsigned)lane_id();
  const unsigned limit = QNEW + n_end;
  const unsigned want2 = key * 0x00010001u;
  unsigned found = 0xFFFFFFFFu;
#pragma unroll
  for (unsigned r = 0; r < 3; ++r) {
    const unsigned v = lane + 32u * r;
    if (v < (unsigned)QCAP / 8u) {
      const uint4 k = reinterpret_cast<const uint4*>(sm.q_key)[v];
      const unsigned w[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
      for (unsigned j = 0; j < 4; ++j) {
        const unsigned m = __vcmpeq2(w[j], want2);
        const unsigned i0 = 8u * v + 2u * j;
        if ((m & 0xFFFFu) && i0 < limit) found = min(found, i0);
        if ((m >> 16) && i0 + 1u < limit) found = min(found, i0 + 1u);
      }
    }
  }
  return __reduce_min_sync(0xffffffffu, found);
}

// the channel-dependent rest, once `act` (-1 = Nothing) is known; writes the next event to `out`
template <bool GRID_BIT = true>
__device__ __forceinline__ void ev_env_post(SmemF& sm, Ev& out, EnvNext nx, const int type, const int cell,
                                            const int ev_ch, const int act, const unsigned n_end) {
  const int lane = lane_id();
  if (nx.ch == CH_PENDING) nx.ch = act;  // the call accepted just now ends before anything else happens
  bool rekey = false;
  if (type == EV_END && act >= 0 && act != ev_ch) {
    // END toCh, Simulator.hs:129-130: reassign cell act -> ev_ch (EventGen.hs:62-72): the call that used
    // `act` continues on ev_ch, so its END event is re-keyed -- unless that event is the one just popped
    if (nx.type == EV_END && nx.cell == cell && nx.ch == act) nx.ch = ev_ch;
    else rekey = true;
  }
#ifdef DCA_END_SLOT_INDEX
  (void)n_end;
  unsigned rekey_slot = 0;
#else
  unsigned rekey_slot = 0;
  if (rekey) rekey_slot = q_find_end(sm, (unsigned)((cell << 7) | act), n_end);  // warp-uniform
#endif
  if (lane == 0) {
    if (GRID_BIT && act >= 0) sm.grid[cell][act >> 5] ^= 1u << (act & 31);  // executeAction, Gridfuncs.hs:105-110
    if (nx.pend_slot >= 0) {
      sm.q_key[nx.pend_slot] = (uint16_t)((cell << 7) | act);
#ifdef DCA_END_SLOT_INDEX
      sm.end_slot[cell][act] = (uint16_t)nx.pend_slot;
#endif
    }
    if (rekey) {
#ifdef DCA_END_SLOT_INDEX
      rekey_slot = sm.end_slot[cell][act];
      sm.end_slot[cell][ev_ch] = (uint16_t)rekey_slot;
#endif
      if (rekey_slot < (unsigned)QCAP) sm.q_key[rekey_slot] = (uint16_t)((cell << 7) | ev_ch);
    }
    out.time = nx.time; out.ch = nx.ch; out.hoff = nx.hoff;
    if (GRID_BIT) { out.type = nx.type; out.cell = nx.cell; }  // (the run kernel publishes kind and cell right after the pop)
  }
  __syncwarp();
}

// order-independent hashes of grid and frep for the parity trace (same as
// state_hashes); event warp only
__device__ __noinline__ void ev_hashes(const SmemF& sm, unsigned long long& hg_out, unsigned long long& hf_out) {
  unsigned long long hg = 0, hf = 0;
#pragma unroll 1
  for (int i = lane_id(); i < NFEAT * NCELL; i += 32) {
    const int cell = i / NFEAT, f = i - cell * NFEAT;
    hf += (unsigned long long)sm.x[f * PLANE + pos_of(cell)] * mix64(0x10000ULL + (unsigned long long)i);
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
__device__ __noinline__ int ev_verify(const SmemF& sm, bool rc, bool nb) {
  const int lane = lane_id();
  bool bad_rc = false, bad_nb = false;
  if (rc) {
#pragma unroll 1
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
  if (nb) {
#pragma unroll 1
    for (int i = lane; i < WPAD; i += 32) bad_nb |= sm.x[i] > 70;
  }
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
#pragma unroll 1
  for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x) d[i] = s[i];
}
__device__ __forceinline__ void copy8(void* dst, const void* src, int bytes) {
  const uint2* s = reinterpret_cast<const uint2*>(src);
  uint2* d = reinterpret_cast<uint2*>(dst);
#pragma unroll 1
  for (int i = threadIdx.x; i < bytes / 8; i += blockDim.x) d[i] = s[i];
}
static_assert(sizeof(((RepState*)0)->x) % 8 == 0 && sizeof(((RepState*)0)->grid) % 16 == 0 &&
              sizeof(((RepState*)0)->q_time) % 16 == 0 && sizeof(((RepState*)0)->q_id) % 16 == 0 &&
              sizeof(((RepState*)0)->q_key) % 16 == 0 && sizeof(((RepState*)0)->q_hoff) % 8 == 0 &&
              offsetof(RepState, x) % 8 == 0 && offsetof(RepState, cnt2) % 8 == 0 && offsetof(RepState, grid) % 16 == 0 &&
              offsetof(RepState, q_time) % 16 == 0 && offsetof(RepState, q_id) % 16 == 0 &&
              offsetof(RepState, q_key) % 16 == 0 && offsetof(RepState, q_hoff) % 8 == 0, "blob arrays");

// everything but wGradCorr (all threads of the CTA)
__device__ __forceinline__ void stage_in(SmemF& sm, const RepState* g) {
#pragma unroll 1
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
  copy16(sm.q_time, g->q_time, sizeof g->q_time);
  copy16(sm.q_id, g->q_id, sizeof g->q_id);
  copy16(sm.q_key, g->q_key, sizeof sm.q_key);
  copy8(sm.q_hoff, g->q_hoff, sizeof sm.q_hoff);
  if (threadIdx.x < 72) sm.ev[0].cand[threadIdx.x] = sm.ev[1].cand[threadIdx.x] = sm.ev[2].cand[threadIdx.x] = 0;  // (read before ncand is known)
  if (threadIdx.x < NCELL) {
    const unsigned long long n4 = c_tab.n4excl[threadIdx.x], n2 = c_tab.n2incl[threadIdx.x];
    unsigned long long r4 = 0, r2 = 0;
#pragma unroll
    for (int r = 0; r < 7; ++r) {
      r4 |= ((n4 >> (7 * r)) & 0x7FULL) << (8 * r);
      r2 |= ((n2 >> (7 * r)) & 0x7FULL) << (8 * r);
    }
    sm.n4rows[threadIdx.x] = make_uint2((uint32_t)r4, (uint32_t)(r4 >> 32));
    sm.n2rows[threadIdx.x] = make_uint2((uint32_t)r2, (uint32_t)(r2 >> 32));
  }
  if (threadIdx.x == 0) {
    sm.P[NCH + 1] = 0.0f;  // the pad leaf of the V tree
    Ev& ev = sm.ev[0];
    ev.time = g->ev_time; ev.type = g->ev_type; ev.cell = g->ev_cell; ev.ch = g->ev_ch; ev.hoff = g->ev_hoff;
    ev.ncand = 0;
    sm.status = g->status;
    sm.alpha_net = g->alpha_net; sm.alpha_avg = g->alpha_avg; sm.alpha_grad = g->alpha_grad;
  }
}
__device__ __forceinline__ void stage_out(const SmemF& sm, RepState* g) {
#pragma unroll 1
  for (int row = threadIdx.x; row < NROW; row += blockDim.x) {
    const int f = row / 7, rr = row - 7 * f;
    float4* dst = reinterpret_cast<float4*>(&g->w[f * PLANE + rr * ROWPAD]);
    dst[0] = *reinterpret_cast<const float4*>(&sm.w[0][row * 4]);
    dst[1] = *reinterpret_cast<const float4*>(&sm.w[1][row * 4]);
  }
  copy8(g->x, sm.x, sizeof sm.x);
  copy8(g->cnt2, sm.cnt2, sizeof sm.cnt2);
  copy16(g->grid, sm.grid, sizeof sm.grid);
  copy16(g->q_time, sm.q_time, sizeof g->q_time);
  copy16(g->q_id, sm.q_id, sizeof g->q_id);
  copy16(g->q_key, sm.q_key, sizeof sm.q_key);
  copy8(g->q_hoff, sm.q_hoff, sizeof sm.q_hoff);
}
// wGradCorr rows of an agent lane
__device__ __forceinline__ void load_wg(Agent& ag, const RepState* g, const Lane& l) {
  const int rc = l.rv ? l.r : 6;
#pragma unroll
  for (int j = 0; j < NSLOT; ++j) {
    int f = SLOTP * j + l.g8;
    if (f > NCH) f = NCH;
    const float4* src = reinterpret_cast<const float4*>(&g->wg[f * PLANE + rc * ROWPAD]);
    const float4 a = src[0], b = src[1];
    WgRow G;
    G.g01 = make_float2(a.x, a.y);
    G.g23 = make_float2(a.z, a.w);
    G.g45 = make_float2(b.x, b.y);
    G.g6 = b.z;
    wg_put(ag, j, G);
  }
  wg_commit();
}
__device__ __forceinline__ void store_wg(const Agent& ag, RepState* g, const Lane& l) {
#pragma unroll
  for (int j = 0; j < NSLOT; ++j) {  // (warp-uniform: the tensor-memory loads are warp-wide)
    WgRegs gin = wg_get(ag, j);
    const WgRow G = wg_wait(gin);
    const int f = SLOTP * j + l.g8;
    if (l.rv && f <= NCH) {
      float4* dst = reinterpret_cast<float4*>(&g->wg[f * PLANE + l.r * ROWPAD]);
      dst[0] = make_float4(G.g01.x, G.g01.y, G.g23.x, G.g23.y);
      dst[1] = make_float4(G.g45.x, G.g45.y, G.g6, 0.0f);
    }
  }
}

__device__ __forceinline__ void ev_load(EvRegs& e, const RunArgs& a, const RepState* g, int rep) {
  e.n_end = g->n_end; e.next_id = g->next_id;
  {
    const long long iter = g->iter, to_go = g->n_events - iter, to_50 = 50 - iter;
    e.left_success = to_go > 0x7fffffffLL ? 0x7fffffff : (to_go < 0 ? -1 : (int)to_go);  // (<= 0: the rule never fires again)
    e.left_50 = to_50 < 0 ? 0 : (int)to_50;
  }
  e.loss_sum = g->loss_sum; e.min_loss = g->min_loss;
  e.mean_new = g->mean_new; e.call_dur = g->call_dur; e.call_dur_hoff = g->call_dur_hoff;
  {  // hoffProb as an integer threshold on the 53 random bits (exact: scaling by 2^53 and ceil are)
    const double h = g->hoff_prob;
    e.hoff_thresh = !(h > 0.0) ? 0ULL : (h >= 1.0 ? (1ULL << 53) : (unsigned long long)ceil(h * 9007199254740992.0));
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) e.stats[k] = 0;
  e.dirty1 = -1;
  e.dbase = g->ndraws - 32ULL; e.dpos = 32; e.dword = 0; e.dnl = 0.0;  // (empty draw cache: the first use refills it)
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
// (separate register allocations) and meet at the CTA barriers.
// sync_end (TRACE instantiation: parity traces, verify flags): one more barrier per event so that the
// event warp sees the afterstate before it hashes / verifies it.
// ---------------------------------------------------------------------------
template <bool TRACE, bool LA>
__device__ __forceinline__ void agent_main(SmemF& sm, const RunArgs& a, RepState* g, Agent& ag, const int n_ev) {
  constexpr bool sync_end = TRACE;  // (the host launches the TRACE instantiation whenever a verify flag is set)
  const Lane l = make_lane();
  const int rc = l.rv ? l.r : 6;
  stage_in(sm, g);
  load_wg(ag, g, l);
  float avgR = g->avg_reward;
  int ncalls = g->ncalls;
  cta_barrier();
  dense_pass<false>(sm, ag, l, -1, false, TdScalars{0.f, 0.f, 0.f, 0.f}, make_uint2(0, 0), make_uint2(0, 0));
  cta_barrier();  // B1
#ifdef DCA_PHASE_CLOCK
  unsigned ph[6] = {0, 0, 0, 0, 0, 0};
#endif
  int cur = 0;  // index of the current event's record (it % 3)
#pragma unroll 1
  for (int it = 0; it < n_ev; ++it) {
    DCA_CLK(c0);
    const Ev& ev = sm.ev[cur];
    const int la_idx = LA ? cur : 0;
    (void)la_idx;
    cur = cur == 2 ? 0 : cur + 1;
    const int n = ev.ncand;
    const bool is_end = ev.type == EV_END;
    const uint2 n4b_r = ev.n4b[rc], n2b_r = ev.n2b[rc];
    // accFn1 = getAction (Simulator.hs:154-172)
    DecidePrep dp = decide_prepare(sm, ev, n, is_end, avgR, ncalls);
    const VTree vt = q_phase(sm, l, ev, n, is_end, n4b_r, n2b_r);
    bool la = false;
    if constexpr (LA) {
      la = is_end && n > 0 && ev.hoff != HOFF_NONE;  // (uniform) the departure half of a hand-off
      if (la) q_lookahead(sm, l, ev, la_of(sm).la[la_idx], n, n4b_r, n2b_r, vt);
    }
    dp.ncdot = -__fmul_rn(dp.c, vt.dotg);
    DCA_CLK(c1);
    agents_q_ready();  // B2
    DCA_CLKB(c2);
    const Decision d = (LA && la) ? decide_la(sm, ev, vt, dp, n, avgR, ncalls) : decide(sm, ev, vt, dp, n, avgR, ncalls);
    if (l.vw == 0) {  // warp-uniform: the event warp takes the decision from here
      if (lane_id() == 0) *reinterpret_cast<int4*>(&sm.pub) = make_int4(d.act, __float_as_int(d.tderr), __float_as_int(avgR), ncalls);
      decision_published();
    }
    DCA_CLK(c3);
    // accFn2 (gridStepTrain, SimRunner.hs:100-118): afterstate + dense update + the next event's sums
    dense_pass<true, true>(sm, ag, l, d.act, is_end, d.td, n4b_r, n2b_r);
    DCA_CLK(c4);
    if (sync_end) agents_afterstate_done();
    agents_wait_next();  // B1
    DCA_CLKB(c5);
    DCA_ACC(0, c0, c1); DCA_ACC(1, c1, c2); DCA_ACC(2, c2, c3); DCA_ACC(3, c3, c4); DCA_ACC(4, c4, c5); DCA_ACC(5, c0, c5);
    if (sm.status != ST_RUNNING) break;  // uniform
  }
#ifdef DCA_PHASE_CLOCK
  if (lane_id() == 0)
    for (int i = 0; i < 6; ++i) atomicAdd(&g_phase[l.vw][i], (unsigned long long)ph[i]);
#endif
  cta_barrier();
  stage_out(sm, g);
  store_wg(ag, g, l);
  if (l.vw == 0 && lane_id() == 0) {
    g->avg_reward = avgR;
    g->ncalls = ncalls;
  }
}

template <bool TRACE, bool LA>
__device__ __forceinline__ void event_main(SmemF& sm, const RunArgs& a, RepState* g, int rep, const int n_ev) {
  constexpr bool sync_end = TRACE;  // (the host launches the TRACE instantiation whenever a verify flag is set)
  const int lane = lane_id();
  stage_in(sm, g);
  float avgR = 0.0f;  // (trace only: taken from the published decision)
  int ncalls = 0;
  EvRegs e;
  ev_load(e, a, g, rep);
  const long long iter0 = TRACE ? g->iter : 0;  // (the trace is indexed by the replica's iteration count)
  (void)iter0;
  cta_barrier();
  QBlk qb;
  q_build(sm, qb, e.n_end);
  ev_candidates(sm, sm.ev[0], -1);
  if (LA) {
    for (int i = lane; i < 72; i += 32) la_of(sm).la[0].hcand[i] = la_of(sm).la[1].hcand[i] = la_of(sm).la[2].hcand[i] = 0;  // (read speculatively)
    __syncwarp();
    if (sm.ev[0].type == EV_END && sm.ev[0].hoff != HOFF_NONE) la_candidates(sm, la_of(sm).la[0], sm.ev[0].hoff);
  }
  // The event warp runs one step AHEAD of the agents: the new events of an event (ev_env_push needs only whether it
  // has a candidate) are pushed at the end of the previous event's update phase, its successor is popped while the
  // agents evaluate its candidates.  pend_slot / q_overflow carry the result of the push to the event it belongs to.
  int pend_slot = -1;
  bool q_overflow = false;
  if (n_ev > 0) {
    const Ev& ev0 = sm.ev[0];
    if (ev0.type != EV_END) ev_draws(e);
    const int pushed = ev_env_push(sm, e, qb, ev0.type, ev0.cell, ev0.time, ev0.ncand > 0);  // Simulator.hs:101-128
    q_overflow = pushed == PEND_OVERFLOW;
    pend_slot = q_overflow ? -1 : pushed;
  }
  cta_barrier();  // B1
  int it = 0;
#ifdef DCA_PHASE_CLOCK
  unsigned ph[6] = {0, 0, 0, 0, 0, 0};
  unsigned sub[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
  int cur = 0;  // index of the current event's record (it % 3)
#pragma unroll 1
  for (; it < n_ev; ++it) {
    const Ev& ev = sm.ev[cur];
    cur = cur == 2 ? 0 : cur + 1;
    Ev& nx = sm.ev[cur];
    const int type = ev.type, cell = ev.cell, n = ev.ncand;
    const int ev_ch = ev.ch, ev_hoff = ev.hoff;
    const double time = ev.time;
    const bool is_end = type == EV_END;
    DCA_CLK(c0);
    // Ahead of the agents (they may still be in the previous event's dense pass): the pop of the next event -- it
    // does not depend on WHICH channel is chosen, and this event's new events are in the queue already.  Its kind
    // and cell go to the next event record (nobody reads that one before barrier 4).
    const bool overflow_now = q_overflow;
    const EnvNext en = ev_env_pop(sm, e, qb, type, ev_hoff, time, pend_slot);   // Simulator.hs:132-134: pop
    if (lane == 0) { nx.type = en.type; nx.cell = en.cell; }
    DCA_CLK(c1);
    event_wait_decision();  // agent warp 0 has published the decision of this event
    DCA_CLKB(c2);
    struct { int act; float tderr; } d;
    {
      const int4 p = *reinterpret_cast<const int4*>(&sm.pub);
      d.act = p.x;
      d.tderr = __int_as_float(p.y);
      if (TRACE) { avgR = __int_as_float(p.z); ncalls = p.w; }
    }
    DCA_CLK(c3);
    if (lane == 0 && d.act >= 0) sm.grid[cell][d.act >> 5] ^= 1u << (d.act & 31);  // executeAction, Gridfuncs.hs:105-110
    ev_env_post<false>(sm, nx, en, type, cell, ev_ch, d.act, e.n_end);       // the channel-dependent rest
    DCA_CLK(s0);
    // the queue blocks the pop touched: their minima are needed by the next event's push / pop
    if (e.dirty1 >= 0) q_block_rescan(sm, qb, (unsigned)e.dirty1);  // (rare: the entry moved into the hole was its block's minimum)
    e.dirty1 = -1;
    DCA_CLK(s1);
    ev_masks(sm, nx, en.cell);
    DCA_CLK(s2);
    // the next event's candidate channels (getAction, Simulator.hs:161-165): the grid bit is set (above), cnt2 is in
    // its afterstate once agent warp 2 has arrived at named barrier 1
    cand_wait();
    DCA_CLKB(s3);
    const int n_next = agent_candidates(sm, nx, en.type, en.cell);
    if (LA && en.type == EV_END && en.hoff != HOFF_NONE) la_candidates(sm, la_of(sm).la[cur], en.hoff);  // (uniform)
    DCA_CLK(s4);
    // ssIter += 1 (Simulator.hs:131): the replica's iteration count is iter0 + it + 1 from here on
    // stop rules of runPeriod (SimRunner.hs:150-169), in the reference's order
    int stop = ST_RUNNING;
    if (d.tderr != d.tderr) stop = ST_NAN_LOSS;
    else e.loss_sum = __fadd_rn(e.loss_sum, d.tderr);  // (the number of losses: every event of the unit but a NaN one)
    if (stop == ST_RUNNING && overflow_now) stop = ST_QUEUE_OVERFLOW;
    if (sync_end) {
      event_wait_afterstate();  // the agents have moved frep to the afterstate
      if (stop == ST_RUNNING && (a.verify_rc || a.verify_nb)) {
        const int v = ev_verify(sm, a.verify_rc, a.verify_nb);
        if (v) stop = v;
      }
      if (TRACE && a.trace && iter0 + it < a.trace_cap) {
        unsigned long long hg, hf;
        ev_hashes(sm, hg, hf);
        if (lane == 0) {
          TraceRec& tr = a.trace[(size_t)rep * a.trace_cap + (iter0 + it)];
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
    }
    if (stop == ST_RUNNING && e.min_loss > 0.0f && it + 1 > e.left_50 && fabsf(d.tderr) < e.min_loss) stop = ST_ZERO_LOSS;
    if (stop == ST_RUNNING && it + 1 == e.left_success) stop = ST_SUCCESS;
    if (stop == ST_RUNNING && e.rng_mode == RNG_TAPE && e.dbase + e.dpos + 8 > e.tape_len) stop = ST_TAPE_EXHAUSTED;
    if (lane == 0) sm.status = stop;
    DCA_CLK(s5);
    event_publish_next();  // B1 (non-blocking): the next event's record and the status are published
    if (stop == ST_RUNNING && it + 1 < n_ev) {  // the next event is processed in this unit: its new events
      if (en.type != EV_END) ev_draws(e);
      const int pushed = ev_env_push(sm, e, qb, en.type, en.cell, en.time, n_next > 0);  // Simulator.hs:101-128
      q_overflow = pushed == PEND_OVERFLOW;
      pend_slot = q_overflow ? -1 : pushed;
    }
    DCA_CLK(c4);
    DCA_CLK(c5);
    DCA_ACC(0, c0, c1); DCA_ACC(1, c1, c2); DCA_ACC(2, c2, c3); DCA_ACC(3, c3, c4); DCA_ACC(4, c4, c5); DCA_ACC(5, c0, c5);
    DCA_SUB(0, c3, s0); DCA_SUB(1, s0, s1); DCA_SUB(2, s1, s2); DCA_SUB(3, s2, s3); DCA_SUB(4, s3, s4); DCA_SUB(5, s4, s5); DCA_SUB(6, s5, c4);
    if (stop != ST_RUNNING) { ++it; break; }
  }
  // (it = the number of events this unit processed)
#ifdef DCA_PHASE_CLOCK
  if (lane == 0) {
    for (int i = 0; i < 6; ++i) atomicAdd(&g_phase[3][i], (unsigned long long)ph[i]);
    for (int i = 0; i < 8; ++i) atomicAdd(&g_phase[4][i], (unsigned long long)sub[i]);
  }
#endif
  cta_barrier();
  stage_out(sm, g);
  if (lane == 0) {
    const Ev& ev = sm.ev[cur];  // the next event to process (popped, its new events not pushed yet)
#pragma unroll
    for (int k = 0; k < 8; ++k) g->stats[k] += (long long)e.stats[k];
    g->n_end = e.n_end; g->next_id = e.next_id; g->ndraws = e.dbase + e.dpos;
    g->iter += it;
    g->loss_n += it - (sm.status == ST_NAN_LOSS ? 1 : 0);
    g->loss_sum = e.loss_sum;
    g->ev_time = ev.time; g->ev_type = ev.type; g->ev_cell = ev.cell; g->ev_ch = ev.ch; g->ev_hoff = ev.hoff;
    g->status = sm.status;
  }
}

// The launch is cut into UNITS -- (replica, chunk of a.chunk_events events), one CTA each, numbered chunk-major: block
// u works on chunk u / n_replicas of replica u % n_replicas.  The hardware block scheduler hands the blocks out in
// order as CTAs retire, so a launch ends within one chunk of the ideal time instead of with the slowest of the last,
// partly filled wave of whole-replica CTAs (measured: 297 M events/s for one wave of 740 whole-replica CTAs against
// 320 M for eight).  Chunk c of a replica starts when chunk c - 1 has been written back (chunk_done, release /
// acquire at gpu scope: it may have run on another SM); its block has a smaller index, so it was dispatched
// earlier, and with more replicas than resident CTAs the wait is never taken in practice.
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, const unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
template <bool TRACE, bool LA = false>
__global__ void __launch_bounds__(NT, DCA_MIN_CTAS) k_run_fast(RunArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemF& sm = *reinterpret_cast<SmemF*>(smem_raw);
  const unsigned n_rep = (unsigned)a.n_replicas;
  const unsigned chunk = blockIdx.x / n_rep, rep = blockIdx.x - chunk * n_rep;
  if (chunk >= a.n_chunks) return;
  RepState* g = a.states + rep;
  if (chunk > 0) {  // the replica's previous chunk (possibly on another SM) must be written back
    if (threadIdx.x == 0)
      while (ld_acquire_gpu(&a.chunk_done[rep]) < chunk) __nanosleep(500);
    __syncthreads();
  }
  const long long left = a.max_events - (long long)chunk * a.chunk_events;
  const int n_ev = (int)min(a.chunk_events, left);
  if (n_ev > 0 && g->initialized && g->status == ST_RUNNING) {  // uniform
#ifdef DCA_PHASE_CLOCK
    const long long cta_t0 = clock64();
#endif
    Agent ag;
    tmem_open(sm, ag);
    if ((virtual_tid() >> 5) == EVW) event_main<TRACE, LA>(sm, a, g, (int)rep, n_ev);
    else agent_main<TRACE, LA>(sm, a, g, ag, n_ev);
    tmem_close(sm);
#ifdef DCA_PHASE_CLOCK
    if (threadIdx.x == 0 && blockIdx.x < 8192) {
      unsigned smid;
      asm("mov.u32 %0, %%smid;" : "=r"(smid));
      g_cta[blockIdx.x][0] = (unsigned long long)(clock64() - cta_t0);
      g_cta[blockIdx.x][1] = smid;
    }
#endif
  }
  __syncthreads();  // every thread's part of the replica is written back
  if (threadIdx.x == 0) st_release_gpu(&a.chunk_done[rep], chunk + 1u);
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
  const Lane l = make_lane();
  const int wid = l.vw, lane = lane_id();
  const int rc = l.rv ? l.r : 6;
  Agent ag;
  tmem_open(sm, ag);
  stage_in(sm, g);
  if (wid != EVW) load_wg(ag, g, l);
  float avgR = g->avg_reward;
  int ncalls = g->ncalls;
  __syncthreads();
  if (threadIdx.x == 0) { sm.alpha_net = alpha_net; sm.alpha_avg = alpha_avg; sm.alpha_grad = alpha_grad; }
  Ev& ev = sm.ev[0];
  if (wid == EVW) {
    if (lane == 0) { ev.cell = cell; ev.type = is_end ? EV_END : EV_NEW; }
    __syncwarp();
    ev_candidates(sm, ev, -1);
    if (mode == 1) {  // evaluate exactly the channel the caller executes -- it must be one getAction could have returned
      bool found = ch_in < 0;
      for (int k = lane; k < ev.ncand; k += 32) found |= ev.cand[k] == ch_in;
      found = __any_sync(0xffffffffu, found);
      __syncwarp();
      if (lane == 0) {
        sm.status = found ? 0 : -1;
        ev.ncand = ch_in >= 0 ? 1 : 0;
        ev.cand[0] = (uint8_t)(ch_in >= 0 ? ch_in : 0);
      }
    } else if (lane == 0) {
      sm.status = 0;
    }
  } else {
    dense_pass<false>(sm, ag, l, -1, false, TdScalars{0.f, 0.f, 0.f, 0.f}, make_uint2(0, 0), make_uint2(0, 0));
  }
  __syncthreads();
  if (sm.status != 0) {  // (uniform) the channel is not eligible / not in use at this cell: nothing is changed
    if (threadIdx.x == 0) out->ch = STEP_BAD_CHANNEL;
    tmem_close(sm);
    return;
  }
  const int n = ev.ncand;
  const uint2 n4b_r = ev.n4b[rc], n2b_r = ev.n2b[rc];
  const VTree vt = wid != EVW ? q_phase(sm, l, ev, n, is_end, n4b_r, n2b_r) : v_tree(sm);
  __syncthreads();
  if (mode == 0) {
    float qb;
    const int act = n > 0 ? argmax_last_warp(sm.q, ev.cand, ev.cand[lane], n, qb) : -1;
    if (threadIdx.x == 0) out->ch = act;
    tmem_close(sm);
    return;
  }
  DecidePrep dp = decide_prepare(sm, ev, n, is_end, avgR, ncalls);
  dp.ncdot = -__fmul_rn(dp.c, vt.dotg);
  const Decision d = decide(sm, ev, vt, dp, n, avgR, ncalls);
  if (wid == EVW) {
    if (lane == 0 && d.act >= 0) {  // executeAction, Gridfuncs.hs:105-110
      const uint32_t bit = 1u << (d.act & 31);
      if (is_end) sm.grid[cell][d.act >> 5] &= ~bit;
      else sm.grid[cell][d.act >> 5] |= bit;
    }
  } else {
    dense_pass<true>(sm, ag, l, d.act, is_end, d.td, n4b_r, n2b_r);
  }
  __syncthreads();
  // write back what the step changed (the event queue is untouched)
#pragma unroll 1
  for (int row = threadIdx.x; row < NROW; row += blockDim.x) {
    const int f = row / 7, rr = row - 7 * f;
    float4* dst = reinterpret_cast<float4*>(&g->w[f * PLANE + rr * ROWPAD]);
    dst[0] = *reinterpret_cast<const float4*>(&sm.w[0][row * 4]);
    dst[1] = *reinterpret_cast<const float4*>(&sm.w[1][row * 4]);
  }
  copy8(g->x, sm.x, sizeof sm.x);
  copy8(g->cnt2, sm.cnt2, sizeof sm.cnt2);
  copy16(g->grid, sm.grid, sizeof sm.grid);
  if (wid != EVW) store_wg(ag, g, l);
  if (threadIdx.x == 0) {
    g->avg_reward = avgR;
    g->ncalls = ncalls;
    out->ch = d.act;
    out->loss = d.tderr;
    out->loss_is_nan = (d.tderr != d.tderr) ? 1 : 0;
    out->reward = ncalls;
  }
  tmem_close(sm);
}

}  // namespace fast
}  // namespace dca
