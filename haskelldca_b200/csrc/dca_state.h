// dca_state.h -- per-replica device state layout shared by host code and kernels.
//
// One replica = one simulator instance of the reference (SimState,
// src/Simulator.hs:59-73): grid, frep, agent, event generator, stats.  The blob
// is what a CTA stages into shared memory for the whole run.
//
// Layout choices (see DESIGN.md "Data layout"):
//  * feature-plane-major, rows padded 7 -> 8: element (f, r, c) lives at
//    f*56 + r*8 + c.  A row is one 32-byte (fp32) / 8-byte (u8) aligned vector
//    load, and the FAST reduction tree (row -> plane -> group) follows it.
//  * frep as u8 (in-use counts <= 42, n_elig <= 70), grid as 70-bit masks,
//    cnt2[ch][cell] = users of ch within distance 2 of cell incl. itself, so that
//    eligibility (Gridfuncs.hs:69-76) is cnt2 == 0 and the nested neighbour scan of
//    incAfterStateFreps (Gridfuncs.hs:220-232) is cnt2 == 0 (NEW) / cnt2 == 1 (END).
//  * the event queue (EventGen, src/EventGen/Internal.hs:44-53) as a dense slot
//    table: slots 0..48 hold the one pending NEW event of each cell, slots 49..
//    hold one END event per call in progress (at most 9 cells can share a
//    channel under the reuse constraint => at most 630 calls), popped by argmin
//    over (time, id).  A HOFF event is never stored: it always directly follows
//    its END event (same time, next id; EventGen.hs:107-108).
#pragma once
#include <stdint.h>

namespace dca {

constexpr int ROWS = 7, COLS = 7, NCELL = 49, NCH = 70, NFEAT = 71;
constexpr int ROWPAD = 8;              // 7 columns + 1 pad
constexpr int PLANE = ROWS * ROWPAD;   // 56
constexpr int WPAD = NFEAT * PLANE;    // 3976 padded elements
constexpr int NROW = NFEAT * ROWS;     // 497 rows
constexpr int FREP = NCELL * NFEAT;    // 3479
constexpr int QNEW = NCELL;            // NEW slots
constexpr int QEND = 630;              // max simultaneous calls
constexpr int QCAP = 680;              // QNEW + QEND, padded to a multiple of 8
constexpr int HOFF_NONE = 255;

enum { EV_NEW = 0, EV_END = 1, EV_HOFF = 2 };
enum { ST_RUNNING = 0, ST_SUCCESS, ST_ZERO_LOSS, ST_NAN_LOSS, ST_REUSE_VIOLATED,
       ST_NELIG_OVERFLOW, ST_TAPE_EXHAUSTED,
       ST_UNINITIALIZED,    // DCA_RNG_TAPE replica whose tape has not been set (or is shorter than the 49 initial draws)
       ST_QUEUE_OVERFLOW }; // more than QEND calls in progress: only reachable from a state that violates the reuse constraint
enum { ORDER_REF = 0, ORDER_FAST = 1 };
enum { RNG_TAPE = 0, RNG_PHILOX = 1 };

struct alignas(16) RepState {
  float w[WPAD];            // wNet
  float wg[WPAD];           // wGradCorr
  uint8_t x[WPAD];          // frep
  uint8_t cnt2[WPAD];       // planes 0..69 used
  uint32_t grid[NCELL][4];  // words 0..2 = channels 0..69, word 3 unused
  double q_time[QCAP];
  uint32_t q_id[QCAP];
  uint16_t q_key[QCAP];     // cell << 7 | ch   (NEW slots: cell << 7 | 127)
  uint8_t q_hoff[QCAP];     // END: hand-off target cell or HOFF_NONE
  // ---- scalars ----
  double ev_time;           // current event (ssEvent)
  double mean_new;          // 1 / (callRate / 60), EventGen.hs:81-82
  double call_dur, call_dur_hoff, hoff_prob;
  unsigned long long ndraws;
  unsigned long long stream;    // Philox stream id = first_replica + replica
  long long iter;               // ssIter
  long long n_events;
  long long loss_n;
  long long stats[8];
  int ev_type, ev_cell, ev_ch, ev_hoff;
  unsigned n_end, next_id;      // egId
  int status, ncalls;
  float avg_reward, loss_sum;
  float alpha_net, alpha_avg, alpha_grad, min_loss;
  int initialized, pad0;
  unsigned long long pad1;
};
static_assert(sizeof(RepState) % 16 == 0, "RepState must be a multiple of 16 bytes");

struct TraceRec {  // == dca_trace
  double time;
  float loss, avg_reward;
  int32_t reward;
  uint8_t etype, cell;
  int8_t ch;
  uint8_t ncand;
  int8_t end_ch;
  uint8_t pad[7];
  uint64_t grid_hash, frep_hash;
};
static_assert(sizeof(TraceRec) == 48, "TraceRec layout");

// host-built neighbourhood tables (uploaded once per device)
struct Tables {
  uint64_t n2incl[NCELL];   // bit (r*7+c): cells within distance 2, incl. self
  uint64_t n4excl[NCELL];   // cells within distance 4, excl. self
  uint8_t n2list[NCELL][20];  // getNeighs 2 cell False in reference order (<= 18)
  uint8_t n2cnt[NCELL];
  uint8_t pos[NCELL];       // (cell/7)*8 + cell%7
};

}  // namespace dca
