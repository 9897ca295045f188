/*
 * dca_oracle.h -- CPU ORACLE for the haskelldca per-event hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check
 * in __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may
 * load it.  Nothing under haskelldca_b200/ links, imports or executes it.
 *
 * It is a plain-C restatement of the reference's algorithm (tsoernes/haskelldca,
 * Haskell + Accelerate); every function cites the reference file:line it
 * follows.  The reference cannot be built in this environment (no GHC / LLVM 6),
 * so the restatement is pinned against the reference's own fixtures instead:
 * src/ExGrids.hs (frep == featureRep grid, END clears the chosen channel) and
 * the known-answer tests of test/GridfuncsSpec.hs, test/GridneighsSpec.hs and
 * test/EventGenSpec.hs -- see tests/test_oracle_*.py.
 *
 * PARITY UNPINNED (no reference test or fixture covers them; the oracle is
 * authoritative by construction from the cited lines): Agent.backward numerics,
 * the argmax tie rule, the RNG stream (mersenne-random-pure64 0.2.2.0,
 * random-fu 0.2.7.0 -- third-party, not vendored in the reference) and the
 * fp32 summation order of Accelerate's `fold`.
 *
 * Two arithmetic "orders" are provided for every fp32 reduction:
 *   ORC_ORDER_REF  : what the reference's default backend (the Accelerate
 *                    Interpreter, src/Opt.hs:68-71) does -- dense n x 3479
 *                    afterstate GEMV, each dot a sequential RIGHT fold
 *                    x0+(x1+(...+(x_{n-1}+0))) in flatten order, products and
 *                    sums rounded separately, no FMA.
 *   ORC_ORDER_FAST : the same dense dots re-associated into fixed trees that
 *                    follow the thread layout of the sm_100a kernel (even/odd
 *                    fma chains over the 7 columns of a row -> xor-butterfly
 *                    over the rows of a feature plane -> butterfly over the
 *                    planes; see dot_fast / dot_fast_g) and the TDC update
 *                    written with explicit FMAs.  Accelerate's `fold`
 *                    requires an associative operator and leaves the order
 *                    unspecified, so this is a legal execution of the same
 *                    program; it is the order the throughput kernel uses.
 */
#ifndef DCA_ORACLE_H
#define DCA_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_ROWS 7
#define ORC_COLS 7
#define ORC_CHANNELS 70
#define ORC_NCELLS 49
#define ORC_NFEATS 71                 /* 70 in-use counts + n_elig, src/Base.hs:51 */
#define ORC_GRID_SIZE (49 * 70)       /* 3430 */
#define ORC_FREP_SIZE (49 * 71)       /* 3479 */

enum { ORC_ORDER_REF = 0, ORC_ORDER_FAST = 1 };
enum { ORC_RNG_MT19937_64 = 0, ORC_RNG_PHILOX = 1 };
enum { ORC_NEW = 0, ORC_END = 1, ORC_HOFF = 2 };          /* src/Base.hs:67-83 */

/* SimStop, src/SimRunner.hs:45-50 (+ running) */
enum {
  ORC_RUNNING = 0,        /* "Paused": more events may be run */
  ORC_SUCCESS = 1,
  ORC_ZERO_LOSS = 2,
  ORC_NAN_LOSS = 3,
  ORC_REUSE_VIOLATED = 4, /* InternalError ReuseConstraintViolated */
  ORC_NELIG_OVERFLOW = 5, /* InternalError NEligOverflow */
  ORC_HOST_ERROR = 6      /* Haskell exception analogue (empty queue, fromJust) */
};

/* Opt record, src/Opt.hs:10-25 (+ the knobs the batched build adds) */
typedef struct {
  double call_dur;        /* --call_dur, minutes, default 3.0 */
  double call_dur_hoff;   /* --call_dur_hoff, default 1.0 */
  double call_rate;       /* --call_rate, per hour, default 200 */
  double hoff_prob;       /* --hoff_prob, default 0 */
  int64_t n_events;       /* --n_events, default 10000 */
  int64_t log_iter;       /* --log_iter, default 1000 */
  float alpha_net;        /* --learning_rate, default 2.52e-6 */
  float alpha_avg;        /* --learning_rate_avg, default 0.06 */
  float alpha_grad;       /* --learning_rate_grad, default 5e-6 */
  float min_loss;         /* --min_loss, default 0 */
  int32_t verify_reuse_constraint;
  int32_t verify_nelig_bounds;
  uint64_t seed;          /* --fixed_rng => 0 */
  int32_t order;          /* ORC_ORDER_* */
  int32_t rng_mode;       /* ORC_RNG_* */
  uint64_t replica;       /* Philox stream id (ignored for MT) */
  int32_t hoff_lookahead; /* EXTENSION (the reference's only TODO, README.md:71-72; 0 = the reference's behaviour):
                             on the END half of a hand-off choose the channel to free by looking ahead at the HOFF
                             arrival that follows in the target cell -- see orc_get_action_lookahead */
  int32_t pad_;
} orc_opt;

void orc_opt_defaults(orc_opt *o);

/* One record per processed event (parity traces). 48 bytes, no padding holes. */
typedef struct {
  double time;            /* time of the processed event */
  float loss;             /* TD error (NaN if NaN) */
  float avg_reward;       /* after the update */
  int32_t reward;         /* calls in progress after the action */
  uint8_t etype;          /* ORC_NEW/END/HOFF */
  uint8_t cell;           /* r*7+c */
  int8_t ch;              /* chosen channel, -1 = Nothing */
  uint8_t ncand;          /* number of candidate channels */
  int8_t end_ch;          /* END events: the terminating channel, else -1 */
  uint8_t pad[7];
  uint64_t grid_hash;     /* after the action */
  uint64_t frep_hash;     /* after the action */
} orc_trace;

/* ---- neighbourhoods: src/Gridneighs.hs:92-128, src/Gridconsts.hs:24-71 ---- */
/* cells out as (r,c) pairs; returns the count */
int orc_periphery(int d, int r, int c, int *out_rc);
int orc_get_neighs(int d, int r, int c, int include_self, int *out_rc);
/* table-based lookup = getNeighborhoodAcc (slit of V by M[r,c]) */
int orc_get_neighborhood_acc(int d, int r, int c, int include_self, int *out_rc);
/* sizes of V (_neighs), V' (_neighsO), _neighs2 ; copies M (49x5) if non-null */
void orc_table_sizes(int *n_neighs, int *n_neighs_o, int *n_neighs2, int64_t *m_out);

/* ---- grid functions: src/Gridfuncs.hs ---- */
/* grid: uint8[7][7][70] (Bool); frep: int64[7][7][71] (Accelerate Int) */
int orc_eligible_chs(const uint8_t *grid, int r, int c, int64_t *chs_out);   /* :69-82 */
int orc_inuse_chs(const uint8_t *grid, int r, int c, int64_t *chs_out);      /* :86-87 */
int orc_violates_reuse_constraint(const uint8_t *grid);                      /* :94-101 */
void orc_execute_action(int is_end, int r, int c, int ch, uint8_t *grid);    /* :105-110 */
void orc_afterstates(const uint8_t *grid, int r, int c, int is_end,
                     const int64_t *chs, int n, uint8_t *grids_out);         /* :116-123 */
void orc_feature_rep(const uint8_t *grid, int64_t *frep);                    /* :153-170 */
void orc_feature_rep_slow(const uint8_t *grid, int64_t *frep);               /* :132-150 */
void orc_inc_afterstate_freps(int r, int c, int is_end, const int64_t *chs, int n,
                              const uint8_t *grid, const int64_t *frep,
                              int64_t *freps_out);                           /* :186-252 */
int64_t orc_bool_sum(const uint8_t *grid);                                   /* AccUtils.hs:166 */

/* ---- agent: src/Agent.hs, src/AccUtils.hs:203-228, src/Simulator.hs:137-205 ---- */
float orc_dense_dot(int order, const int64_t *frep, const float *w);         /* vvMul / mvMul row */
float orc_dense_dot_g(int order, const int64_t *frep, const float *wg);      /* vvMul x wGradCorr, Agent.hs:66 */
float orc_forward(int order, const int64_t *frep, const float *w);           /* Agent.hs:22-23 */
int orc_argpmax(const float *q, int n);                                      /* last max wins */
/* selectAction: returns idx (into chs), writes qval and the chosen afterstate frep */
int orc_select_action(int order, int r, int c, int is_end, const int64_t *chs, int n,
                      const float *w, const uint8_t *grid, const int64_t *frep,
                      float *qval_out, int64_t *afrep_out, float *qvals_out);
/* getAction (accFn1): returns ch or -1; next_frep gets afterstate (or frep) */
int orc_get_action(int order, int r, int c, int is_end, const float *w,
                   const uint8_t *grid, const int64_t *frep, int64_t *next_frep,
                   int *ncand_out);
/* Hand-off look-ahead (extension; README.md:71-72 names it as the reference's TODO, the reference defines nothing
 * more).  For an event END _ (Just toCell) at (r, c) -- the departure half of a hand-off, directly followed by the
 * HOFF arrival at toCell (EventGen.hs:94-109) -- the value of freeing the in-use channel k is not V(afterstate_k) but
 *   q[k] = max over h eligible at toCell ON afterstate_k of V(afterstate_k with h assigned at toCell),
 *   q[k] = V(afterstate_k) when no channel is eligible there (the hand-off will be rejected),
 * every V a dense dot in `order` of the materialised feature representation (incAfterStateFreps applied twice); the
 * maximum is a left fold with `>` over ascending h, the channel is argpmax over k (last maximum wins) as in selectAction.
 * Only the SELECTION looks ahead: next_frep is afterstate_k of the chosen k, so backward (Agent.hs:44-81) is unchanged. */
int orc_get_action_lookahead(int order, int r, int c, int hoff_r, int hoff_c, const float *w,
                             const uint8_t *grid, const int64_t *frep, int64_t *next_frep,
                             int *ncand_out, float *qvals_out);
/* backward: updates w, wg, avg_reward in place; returns loss; *is_nan set */
float orc_backward(int order, float alpha_n, float alpha_a, float alpha_g,
                   const int64_t *frep, const int64_t *next_frep, float reward,
                   float *w, float *wg, float *avg_reward, int *is_nan);
/* gridStepTrain (accFn2): ch<0 = Nothing; returns loss, *reward_out = boolSum */
float orc_grid_step_train(int order, float alpha_n, float alpha_a, float alpha_g,
                          int is_end, int r, int c, int ch,
                          const int64_t *frep, const int64_t *next_frep,
                          uint8_t *grid, float *w, float *wg, float *avg_reward,
                          int *is_nan, int64_t *reward_out);

/* ---- RNG: mersenne-random-pure64 / random-fu semantics; Philox4x32-10 ---- */
typedef struct orc_rng orc_rng;
orc_rng *orc_rng_create(int mode, uint64_t seed, uint64_t replica);
void orc_rng_destroy(orc_rng *g);
uint64_t orc_rng_word64(orc_rng *g);               /* getRandomWord64 */
double orc_rng_double(orc_rng *g);                 /* getRandomDouble  */
double orc_rng_exponential(orc_rng *g, double mean);
int orc_rng_uniform_int(orc_rng *g, int lo, int hi); /* integralUniform, inclusive */
uint64_t orc_rng_ndraws(const orc_rng *g);
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double orc_neglog_u53(uint64_t word);              /* the Philox-mode -log(u) spec */

/* ---- EventGen: src/EventGen.hs, src/EventGen/Internal.hs ---- */
typedef struct orc_eventgen orc_eventgen;
typedef struct {
  double time;
  int32_t etype;
  int32_t r, c;
  int32_t ch;        /* END only, else -1 */
  int32_t hoff_r;    /* END (Just cell) else -1 */
  int32_t hoff_c;
} orc_event;
orc_eventgen *orc_eg_create(const orc_opt *opt, int with_initial_events); /* _mkEventGen / mkEventGen */
void orc_eg_destroy(orc_eventgen *eg);
int orc_eg_push(orc_eventgen *eg, const orc_event *ev);         /* Internal.hs:70-78 */
int orc_eg_pop(orc_eventgen *eg, orc_event *out);               /* EventGen.hs:39-50; -1 if empty */
int orc_eg_reassign(orc_eventgen *eg, int r, int c, int from_ch, int to_ch); /* :62-72 */
int orc_eg_generate_new(orc_eventgen *eg, double time, int r, int c);        /* :75-85 */
int orc_eg_generate_hoff_new(orc_eventgen *eg, double time, int r, int c, int ch); /* :94-109 */
int orc_eg_generate_end(orc_eventgen *eg, double time, int r, int c, int ch);      /* :111-112 */
int orc_eg_generate_hoff_end(orc_eventgen *eg, double time, int r, int c, int ch); /* :114-116 */
int64_t orc_eg_id(const orc_eventgen *eg);
int orc_eg_sizes(const orc_eventgen *eg, int *heap_n, int *events_n, int *endids_n);
uint64_t orc_eg_ndraws(const orc_eventgen *eg);
orc_rng *orc_eg_rng(orc_eventgen *eg);                          /* egGen */

/* ---- simulator: src/Simulator.hs:59-134, src/SimRunner.hs:60-178 ---- */
typedef struct orc_sim orc_sim;
orc_sim *orc_sim_create(const orc_opt *opt);                    /* mkSimState */
void orc_sim_destroy(orc_sim *s);
/* run up to max_steps runStep iterations; stops early on a SimStop.
 * trace (optional) receives one record per processed event. Returns status. */
int orc_sim_run(orc_sim *s, int64_t max_steps, orc_trace *trace, int64_t trace_cap,
                int64_t *steps_done);
void orc_sim_read(const orc_sim *s, uint8_t *grid, int64_t *frep, float *w, float *wg,
                  float *avg_reward, int64_t *iter, orc_event *next_event,
                  int64_t stats8[8]);
/* sum of losses (fp32 left fold, Stats.hs:100) and count since last reset */
void orc_sim_period(orc_sim *s, float *loss_sum, int64_t *loss_n, int reset);
uint64_t orc_sim_ndraws(const orc_sim *s);
uint64_t orc_grid_hash(const uint8_t *grid);
uint64_t orc_frep_hash(const int64_t *frep);

/* Stats.hs:83-92: (cumuNew, cumuHoff, cumuTot) from the 8 counters */
void orc_stats_cumus(const int64_t stats8[8], double out3[3]);

/* ---- batched CPU baseline: n_replicas independent sims over OpenMP ----
 * per-replica parameter arrays may be NULL (use base opt). Philox stream id =
 * first_replica + i. Returns total processed events; stats summed into out. */
int64_t orc_run_replicas(const orc_opt *base, int64_t first_replica, int n_replicas,
                         int64_t n_events, int n_threads, int64_t stats_sum8[8],
                         float *avg_reward_out, double *bp_out);
int orc_max_threads(void);

/* ---- the optimised CPU competitor (dca_cpu_fast.c): same path, FAST order, delta-form q-values, bit-mask
 * grid, AVX2 over the rows of a feature plane, OpenMP over replicas.  Bit-exact against ORC_ORDER_FAST of
 * the literal oracle above (tests/test_oracle_cpu_fast.py); bench.py's cpu_baseline times it. ---- */
typedef struct cpf_sim cpf_sim;
cpf_sim *cpf_create(const orc_opt *opt);
void cpf_destroy(cpf_sim *s);
int cpf_run(cpf_sim *s, int64_t max_steps, orc_trace *trace, int64_t trace_cap, int64_t *steps_done);
void cpf_read(const cpf_sim *s, uint8_t *grid, int64_t *frep, float *w, float *wg, float *avg_reward,
              int64_t *iter, orc_event *next_event, int64_t stats8[8]);
uint64_t cpf_ndraws(const cpf_sim *s);
int64_t cpf_run_replicas(const orc_opt *base, int64_t first_replica, int n_replicas, int64_t n_events,
                         int n_threads, int64_t stats_sum8[8], float *avg_reward_out, double *bp_out);

#ifdef __cplusplus
}
#endif
#endif /* DCA_ORACLE_H */
