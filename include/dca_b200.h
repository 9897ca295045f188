/*
 * dca_b200.h -- C ABI of libdca_b200.so: the B200 (sm_100a) implementation of
 * haskelldca's per-event hot path, batched over independent simulator replicas.
 *
 * This is the drop-in boundary for the reference's `--backend` dispatch
 * (src/AccUtils.hs:76-95, src/Base.hs:127-130, src/Opt.hs:90-94).  The reference
 * has no FFI; its de-facto operator interface is the pair of compiled Accelerate
 * functions handed to runStep (src/SimRunner.hs:61-62, 211-212):
 *
 *   accFn1 = runN bkend getAction      -> dca_get_action
 *   accFn2 = runN bkend gridStepTrain  -> dca_grid_step_train
 *
 * plus the run loop around them (runPeriod, src/SimRunner.hs:132-178, with
 * environmentStep, src/Simulator.hs:101-134), which the fused kernel replaces
 * wholesale -> dca_run.  INTEGRATION.md shows the `foreign import ccall`
 * bindings a maintainer adds to the Haskell modules.
 *
 * Conventions: plain C types only; opaque handle owned by the library and
 * released by dca_destroy; every function returns 0 on success and a negative
 * DCA_ERR_* otherwise (message via dca_last_error); per-replica *domain*
 * outcomes (SimStop, src/SimRunner.hs:45-50) are data, not errors; all calls are
 * synchronous; a handle is not thread-safe, distinct handles are independent.
 * Host buffers are caller-owned, row-major, fixed width:
 *   grid  uint8_t[7*7*70]   (Accelerate `Array DIM3 Bool`, src/Base.hs:46)
 *   frep  int64_t[7*7*71]   (Accelerate `Array DIM3 Int`,  src/Base.hs:51)
 *   wNet / wGradCorr float[3479], avgReward float (Agent, src/Base.hs:113)
 * There is no CPU fallback: without a CUDA device dca_create fails.
 */
#ifndef DCA_B200_H
#define DCA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCA_ROWS 7
#define DCA_COLS 7
#define DCA_CHANNELS 70
#define DCA_GRID_SIZE 3430 /* 7*7*70 */
#define DCA_FREP_SIZE 3479 /* 7*7*71 */
#define DCA_N_STATS 8
/* The device keeps event ids (egId, src/EventGen/Internal.hs:44-53: only the tie-break between equal event times) in 32
 * bits and an event takes at most three (NEW accepted with a hand-off: next arrival, END, HOFF), so a replica's
 * n_events is limited to 2^32 / 3; iteration and draw counters are 64-bit.  dca_create rejects more (DCA_ERR_ARG). */
#define DCA_MAX_EVENTS 1400000000LL

/* error codes */
#define DCA_OK 0
#define DCA_ERR_ARG (-1)
#define DCA_ERR_CUDA (-2)
#define DCA_ERR_NO_DEVICE (-3)
#define DCA_ERR_STATE (-4)

/* fp32 reduction order (see DESIGN.md "Arithmetic orders") */
#define DCA_ORDER_REF 0  /* Accelerate-Interpreter order: dense right folds, no FMA */
#define DCA_ORDER_FAST 1 /* fixed row/plane/group tree with FMA: the throughput kernel */

/* random source */
#define DCA_RNG_TAPE 0   /* replay a host-supplied word stream (MT19937-64 for --fixed_rng) */
#define DCA_RNG_PHILOX 1 /* Philox4x32-10 keyed (seed, replica), counter = draw index */

/* event types, src/Base.hs:67-83 */
#define DCA_NEW 0
#define DCA_END 1
#define DCA_HOFF 2

/* per-replica status = SimStop, src/SimRunner.hs:45-50 */
#define DCA_RUNNING 0 /* Paused: can run more events */
#define DCA_SUCCESS 1
#define DCA_ZERO_LOSS 2
#define DCA_NAN_LOSS 3
#define DCA_REUSE_VIOLATED 4
#define DCA_NELIG_OVERFLOW 5
#define DCA_TAPE_EXHAUSTED 6 /* DCA_RNG_TAPE ran out of words */
#define DCA_UNINITIALIZED 7  /* DCA_RNG_TAPE replica without a tape yet: dca_run refuses to start (DCA_ERR_STATE) */
#define DCA_QUEUE_OVERFLOW 8 /* > 630 calls in progress: impossible while the reuse constraint holds (9 co-channel
                                cells x 70 channels); reachable only from a grid forced in with dca_write_state */

/* stats counters, src/Stats.hs:15-30 */
#define DCA_ST_ARRIVALS_NEW 0
#define DCA_ST_ACCEPTED_NEW 1
#define DCA_ST_REJECTED_NEW 2
#define DCA_ST_ARRIVALS_HOFF 3
#define DCA_ST_ENDED 4
#define DCA_ST_REJECTED_HOFF 5 /* always 0: src/Simulator.hs:113 books hand-off rejects as NEW */
#define DCA_ST_CURR_ARRIVALS_NEW 6
#define DCA_ST_CURR_REJECTED_NEW 7

typedef struct dca_handle dca_handle;

/* Mirrors the Opt record (src/Opt.hs:10-25).  Simulation parameters are
 * per replica: pass arrays of n_replicas values, or NULL to use the scalar. */
typedef struct {
  int32_t n_replicas;
  int32_t device;         /* CUDA device ordinal */
  uint64_t seed;          /* rngSeed; --fixed_rng => 0 (src/Opt.hs:79-82) */
  uint64_t first_replica; /* global index of this handle's replica 0 (Philox stream id) */
  int32_t rng_mode;       /* DCA_RNG_* */
  int32_t order;          /* DCA_ORDER_* */
  int64_t n_events;       /* nEvents: a replica stops with DCA_SUCCESS at iter == n_events */
  float min_loss;         /* minLoss */
  int32_t verify_reuse_constraint;
  int32_t verify_nelig_bounds;
  int32_t trace_capacity; /* per-replica trace records kept (0 = no tracing) */
  int32_t warps_per_replica; /* accepted for compatibility and ignored: both kernels use 128-thread CTAs */
  double call_dur, call_dur_hoff, call_rate, hoff_prob; /* callDurNew, callDurHoff, callRate, hoffProb */
  float alpha_net, alpha_avg, alpha_grad;               /* alphaNet, alphaAvg, alphaGrad */
  const double *call_dur_v, *call_dur_hoff_v, *call_rate_v, *hoff_prob_v;
  const float *alpha_net_v, *alpha_avg_v, *alpha_grad_v;
  /* EXTENSION -- hand-off look-ahead, the reference's only TODO (README.md:71-72); 0 = the reference's behaviour.
   * On the END half of a hand-off (END ch (Just toCell), src/EventGen.hs:94-109) the channel to free is chosen by
   *   q[k] = max over h eligible at toCell on afterstate_k of V(afterstate_k with h assigned at toCell)
   * (V(afterstate_k) when none is eligible) instead of V(afterstate_k); the TD update still uses the real afterstate.
   * Implemented by dca_run in both orders (the fused FAST kernel has its own instantiation for it); the step-wise
   * operators, whose signatures carry no hand-off target, ignore it. */
  int32_t hoff_lookahead;
  int32_t reserved_;
} dca_config;

/* Event, src/Base.hs:101-105 */
typedef struct {
  double time;
  int32_t etype;          /* DCA_NEW / DCA_END / DCA_HOFF */
  int32_t r, c;
  int32_t ch;             /* END: terminating channel, else -1 */
  int32_t hoff_r, hoff_c; /* END (Just cell), else -1 */
} dca_event;

/* one record per processed event (trace_capacity > 0); 48 bytes */
typedef struct {
  double time;
  float loss;
  float avg_reward;
  int32_t reward;
  uint8_t etype;
  uint8_t cell;
  int8_t ch;
  uint8_t ncand;
  int8_t end_ch;
  uint8_t pad[7];
  uint64_t grid_hash;
  uint64_t frep_hash;
} dca_trace;

/* ---- lifecycle ------------------------------------------------------------ */
/* Opt defaults, src/Opt.hs:28-76 */
int dca_config_defaults(dca_config *cfg);
/* mkSimState for every replica (src/Simulator.hs:88-97): empty grid, zero agent,
 * frep = featureRep grid, 49 initial NEW events (mkEventGen, src/EventGen.hs:54-58),
 * first event popped.  In DCA_RNG_TAPE mode initialisation is deferred until the
 * replica's tape is set. */
int dca_create(const dca_config *cfg, dca_handle **out);
int dca_destroy(dca_handle *h);
/* re-run mkSimState on every replica (same seeds, same parameters) */
int dca_reset(dca_handle *h);
/* replace the per-replica simulation parameters (host arrays of n_replicas values,
 * NULL = keep); they take effect at the next dca_reset */
int dca_set_params(dca_handle *h, const double *call_dur, const double *call_dur_hoff,
                   const double *call_rate, const double *hoff_prob, const float *alpha_net,
                   const float *alpha_avg, const float *alpha_grad);
const char *dca_last_error(const dca_handle *h);
int dca_device_count(void);

/* ---- random tape (DCA_RNG_TAPE) ------------------------------------------- */
/* Replays getRandomWord64 / getRandomDouble (src/EventGen/Internal.hs:57-63):
 * draw k returns words[k]; exponential draws use neglog[k] = -log((words[k]>>11)/2^53)
 * computed by the caller's libm so event times match the host bit for bit. */
int dca_set_rng_tape(dca_handle *h, int32_t replica, const uint64_t *words, const double *neglog,
                     size_t n);
/* convenience: generate the tape with MT19937-64 (pureMT seed) + libm log on the host */
int dca_set_rng_mt19937(dca_handle *h, int32_t replica, uint64_t seed, size_t n_draws);

/* ---- fused run: replaces runPeriod's loop of runStep ------------------------ */
/* Every replica whose status is DCA_RUNNING processes up to max_events events
 * (getAction -> environmentStep -> gridStepTrain each) on the device, stopping
 * early with its own SimStop status.  Returns when the device is done. */
int dca_run(dca_handle *h, int64_t max_events);
/* The same in two halves, so that ONE host thread (the Haskell main thread runs runSim, src/SimRunner.hs:206-243, in a
 * single StateT) can keep every GPU of a box busy with a handle per device: dca_run_async enqueues the launch and
 * returns, dca_wait blocks until it is done (and is a no-op without a launch in flight); every other call on the handle
 * collects a launch in flight first.  dca_run_many = dca_run_async on each handle, then dca_wait on each; it returns
 * the first error.  dca_run(h, n) == dca_run_async(h, n) followed by dca_wait(h). */
int dca_run_async(dca_handle *h, int64_t max_events);
int dca_wait(dca_handle *h);
int dca_run_many(dca_handle **handles, int32_t n, int64_t max_events);
/* DCA_ORDER_FAST: dca_run cuts a launch into units of this many events per replica (one CTA each, chained in
 * order), so that the device stays evenly loaded to the end of the launch; a replica's state is staged from / to
 * HBM once per unit.  Results do not depend on it. */
int64_t dca_run_unit_events(void);
/* device time of the last dca_run / dca_reset (CUDA events on the launch stream) */
int dca_elapsed_ms(const dca_handle *h, float *ms);
int dca_last_launch_count(const dca_handle *h, int64_t *n_kernels);

/* ---- step-wise operators: accFn1 / accFn2 (src/SimRunner.hs:61-62) ---------- */
/* getAction (src/Simulator.hs:154-172) on the replica's device-resident grid,
 * frep and agent.  *ch = chosen channel or -1 (Nothing); the afterstate frep is
 * kept on the device for the following dca_grid_step_train. */
int dca_get_action(dca_handle *h, int32_t replica, int32_t r, int32_t c, int32_t is_end,
                   int32_t *ch);
/* gridStepTrain (src/SimRunner.hs:100-118): execute the action on the grid,
 * reward = calls in progress, average-reward TDC update (src/Agent.hs:44-81),
 * then frep := nextFrep (src/SimRunner.hs:88-90).  ch must be a channel getAction could have
 * returned for this event (eligible at the cell, or in use there for an END) or -1 (Nothing);
 * anything else is DCA_ERR_ARG and changes nothing. */
int dca_grid_step_train(dca_handle *h, int32_t replica, float alpha_net, float alpha_avg,
                        float alpha_grad, int32_t is_end, int32_t r, int32_t c, int32_t ch,
                        float *loss, int32_t *loss_is_nan, int64_t *reward);

/* ---- the same two operators on EXPLICIT arrays (pure functions) --------------- */
/* accFn1 :: Scalar Cell -> Scalar Bool -> Agent -> Grid -> Frep -> (Scalar (Maybe Ch), Frep)
 * exactly as runStep calls it (src/SimRunner.hs:61,75): candidates from `grid`
 * (eligibleChs / inuseChs), afterstate freps from `frep` (incAfterStateFreps), q = mvMul with
 * wnet, argpmax.  *ch = -1 and next_frep = frep when there is no candidate. */
int dca_acc_get_action(int32_t device, int32_t order, int32_t r, int32_t c, int32_t is_end,
                       const float *wnet, const uint8_t *grid, const int64_t *frep, int32_t *ch,
                       int64_t *next_frep);
/* accFn2 :: alphaN -> alphaA -> alphaG -> eIsEnd -> cell -> Maybe Ch -> frep -> nextFrep -> Grid -> Agent
 *        -> (Grid, Agent, Scalar (Maybe Float), Scalar Int)          (src/SimRunner.hs:62,85-86,100-118)
 * grid, wnet, wgrad and *avg_reward are updated in place; ch = -1 is Nothing. */
int dca_acc_grid_step_train(int32_t device, int32_t order, float alpha_net, float alpha_avg,
                            float alpha_grad, int32_t is_end, int32_t r, int32_t c, int32_t ch,
                            const int64_t *frep, const int64_t *next_frep, uint8_t *grid, float *wnet,
                            float *wgrad, float *avg_reward, float *loss, int32_t *loss_is_nan,
                            int64_t *reward);
/* vvMul (src/AccUtils.hs:219-220) = forward (src/Agent.hs:22-23) of one frep with one weight
 * vector in the given order; grad_corr_tree != 0 selects the FAST tree of x . wGradCorr. */
int dca_dense_dot(int32_t device, int32_t order, int32_t grad_corr_tree, const int64_t *frep,
                  const float *w, float *out);

/* ---- state read-back / write ----------------------------------------------- */
/* any output pointer may be NULL */
int dca_read_state(dca_handle *h, int32_t replica, uint8_t *grid, int64_t *frep, float *wnet,
                   float *wgrad, float *avg_reward, int64_t *iter, dca_event *next_event);
/* overwrite grid and/or agent; frep is recomputed with featureRep (src/Gridfuncs.hs:153-170).
 * The event queue is NOT rebuilt: after a grid write it no longer matches the calls in progress
 * (an END event of a removed call would find no channel; a grid that violates the reuse constraint
 * can exceed the queue's 630 slots, which dca_run reports as DCA_QUEUE_OVERFLOW).  Meant for the
 * step-wise operators, warm-started weights and tests. */
int dca_write_state(dca_handle *h, int32_t replica, const uint8_t *grid, const float *wnet,
                    const float *wgrad, const float *avg_reward);
int dca_read_stats(dca_handle *h, int32_t replica, int64_t stats[DCA_N_STATS]);
int dca_read_status(dca_handle *h, int32_t replica, int32_t *status, int64_t *iter,
                    uint64_t *n_draws);
/* sum of losses (fp32, in event order) and their count since the last reset;
 * reset also clears the period counters like statsReportPeriod (src/Stats.hs:117-118) */
int dca_read_period(dca_handle *h, int32_t replica, float *loss_sum, int64_t *n, int32_t reset);
int dca_read_trace(dca_handle *h, int32_t replica, dca_trace *out, size_t cap, size_t *n);
/* bulk: per-replica arrays of length n_replicas (any may be NULL) */
int dca_read_summary(dca_handle *h, int32_t *status, int64_t *iter, float *avg_reward,
                     int64_t *stats /* [n_replicas][8] */);
/* bulk: the learned agents of ALL replicas in the reference's flatten order -- wnet / wgrad are
 * [n_replicas][3479] floats, avg_reward [n_replicas] (any may be NULL).  Unpacked on the device, copied
 * in chunks of 4096 replicas. */
int dca_read_agents(dca_handle *h, float *wnet, float *wgrad, float *avg_reward);
/* device-side sum over replicas: out[0..7] = stats, out[8] = events processed,
 * out[9] = replicas not RUNNING/SUCCESS.  *dev_ptr (optional) receives the device
 * address of the same int64[10] so a caller can all-reduce it across GPUs (NCCL). */
int dca_sum_stats(dca_handle *h, int64_t out[10], void **dev_ptr);
/* Stats.hs:83-92 */
void dca_stats_cumus(const int64_t stats[DCA_N_STATS], double out3[3]);

/* ---- checkpoint / resume (absent in the reference) ------------------ */
/* The complete simulator state of every replica (grid, frep, agent, event queue, RNG position,
 * statistics) as one opaque blob of dca_checkpoint_size(h) bytes.  A blob can be loaded into any
 * handle created with the same n_replicas, order, rng_mode, seed and first_replica (and, in
 * DCA_RNG_TAPE mode, the same tapes) -- the random stream lives in the handle, so only then does
 * running on continue bit for bit; anything else is rejected with DCA_ERR_STATE. */
size_t dca_checkpoint_size(const dca_handle *h);
int dca_checkpoint_save(dca_handle *h, void *buf, size_t size);
int dca_checkpoint_load(dca_handle *h, const void *buf, size_t size);

/* ---- Gridfuncs / Gridneighs operators on explicit host arrays --------------- */
/* Stateless device implementations of the reference's grid functions, for
 * callers and tests that work on arrays rather than a replica handle. */
int dca_get_neighs(int32_t d, int32_t r, int32_t c, int32_t include_self, int32_t *out_rc,
                   int32_t *n);                                  /* Gridneighs.hs:92-98 */
int dca_gf_eligible_chs(int32_t device, const uint8_t *grid, int32_t r, int32_t c, int64_t *chs,
                        int32_t *n);                             /* Gridfuncs.hs:81-82 */
int dca_gf_inuse_chs(int32_t device, const uint8_t *grid, int32_t r, int32_t c, int64_t *chs,
                     int32_t *n);                                /* Gridfuncs.hs:86-87 */
int dca_gf_feature_rep(int32_t device, const uint8_t *grid, int64_t *frep); /* :153-170 */
int dca_gf_inc_afterstate_freps(int32_t device, int32_t r, int32_t c, int32_t is_end,
                                const int64_t *chs, int32_t n, const uint8_t *grid,
                                const int64_t *frep, int64_t *freps_out);   /* :186-252 */
int dca_gf_violates_reuse_constraint(int32_t device, const uint8_t *grid, int32_t *violates);
                                                                            /* :94-101 */

/* ---- multi-GPU (one handle per device, one process or many) ------------------ */
/* Sum the int64[10] of dca_sum_stats over several handles of THIS process with
 * ncclAllReduce (ncclCommInitAll).  With one process per GPU, all-reduce the
 * dev_ptr of dca_sum_stats with the caller's own communicator instead. */
int dca_allreduce_stats(dca_handle **handles, int32_t n, int64_t out[10]);

#ifdef __cplusplus
}
#endif
#endif /* DCA_B200_H */
