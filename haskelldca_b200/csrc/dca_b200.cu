// dca_b200.cu -- C-ABI implementation (include/dca_b200.h) over the sm_100a kernels.
//
// Pure CUDA runtime: no torch types, no CPU compute path.  Host code here only
// builds the neighbourhood tables (Gridneighs.hs:92-128), marshals buffers and
// launches kernels.
#include "../../include/dca_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <random>
#include <string>
#include <vector>

#include "dca_kernels.cuh"
#include "dca_fast.cuh"
#include "dca_ops.cuh"

using namespace dca;

static_assert(sizeof(dca_trace) == sizeof(TraceRec), "trace record layouts must agree");

namespace {

thread_local std::string g_last_error;

#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      char buf_[512];                                                                     \
      snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
               __FILE__, __LINE__);                                                       \
      return fail(h, DCA_ERR_CUDA, buf_);                                                 \
    }                                                                                     \
  } while (0)

// ---- neighbourhood tables (Gridneighs.hs:92-128) --------------------------
// Ring d around (r, c) in axial coordinates, walked in the reference's order:
// six sides, each d cells long, starting at (r, c-d) and turning clockwise in the
// order ps6, ps5, ps4, ps3, ps2, ps1 of Gridneighs.hs:111-121; out-of-grid cells
// are dropped (isValid, :124-128).
struct Side { int sr, sc, dr, dc; };  // start = (r + sr*d, c + sc*d), step (dr, dc)
const Side kSides[6] = {
    {0, -1, -1, +1},   // ps6
    {+1, -1, -1, 0},   // ps5
    {+1, 0, 0, -1},    // ps4
    {0, +1, +1, -1},   // ps3
    {-1, +1, +1, 0},   // ps2
    {-1, 0, 0, +1},    // ps1
};

int ring(int d, int r, int c, int* out_rc) {
  int n = 0;
  for (const Side& s : kSides) {
    int rr = r + s.sr * d, cc = c + s.sc * d;
    for (int k = 0; k < d; ++k, rr += s.dr, cc += s.dc)
      if (rr >= 0 && rr < ROWS && cc >= 0 && cc < COLS) {
        out_rc[2 * n] = rr;
        out_rc[2 * n + 1] = cc;
        ++n;
      }
  }
  return n;
}

int neighs(int d, int r, int c, bool include_self, int* out_rc) {  // getNeighs, :92-98
  int n = 0;
  if (include_self) {
    out_rc[0] = r;
    out_rc[1] = c;
    n = 1;
  }
  for (int dd = 1; dd <= d; ++dd) n += ring(dd, r, c, out_rc + 2 * n);
  return n;
}

void build_tables(Tables& t) {
  memset(&t, 0, sizeof t);
  int buf[2 * 64];
  for (int cell = 0; cell < NCELL; ++cell) {
    int r = cell / COLS, c = cell % COLS;
    t.pos[cell] = (uint8_t)(r * ROWPAD + c);
    int n2 = neighs(2, r, c, false, buf);
    t.n2cnt[cell] = (uint8_t)n2;
    uint64_t m2 = 1ULL << cell;
    for (int i = 0; i < n2; ++i) {
      int o = buf[2 * i] * COLS + buf[2 * i + 1];
      t.n2list[cell][i] = (uint8_t)o;
      m2 |= 1ULL << o;
    }
    t.n2incl[cell] = m2;
    int n4 = neighs(4, r, c, false, buf);
    uint64_t m4 = 0;
    for (int i = 0; i < n4; ++i) m4 |= 1ULL << (buf[2 * i] * COLS + buf[2 * i + 1]);
    t.n4excl[cell] = m4;
  }
}

std::mutex g_tab_mutex;
bool g_tab_uploaded[64] = {false};

}  // namespace

namespace {
struct OpIO {  // host-visible part; the same layout on the device and in pinned host memory
  // inputs, ordered so that every operator's inputs are one contiguous range starting at `grid`
  uint8_t grid[3456];            // DCA_GRID_SIZE rounded up to 64
  float w[FREP + 1];
  float wg[FREP + 1];
  long long frep[FREP + 1];
  long long nfrep[FREP + 1];
  long long chs[NCH + 2];
  // outputs
  long long out_frep[FREP + 1];
  long long reward;
  ops::BackwardOut bout;
  int n, ch, violates;
  float dot;
};
static_assert(DCA_GRID_SIZE <= 3456, "grid staging");

struct OpScratch {
  bool ready = false;
  int device = -1;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  OpIO* d_io = nullptr;       // device
  OpIO* h_io = nullptr;       // pinned host mirror
  RepState* st = nullptr;     // a scratch replica blob (grid masks, cnt2, frep of the caller's grid)
  long long* afreps = nullptr;  // [NCH][FREP]
  float* q = nullptr;         // [72]
  std::mutex mu;
  void release() {
    if (d_io) cudaFree(d_io);
    if (h_io) cudaFreeHost(h_io);
    if (st) cudaFree(st);
    if (afreps) cudaFree(afreps);
    if (q) cudaFree(q);
    if (own_stream && stream) cudaStreamDestroy(stream);
    d_io = h_io = nullptr; st = nullptr; afreps = nullptr; q = nullptr; stream = nullptr;
    ready = false;
  }
};

}  // namespace

struct dca_handle {
  dca_config cfg{};
  int device = 0;
  int n = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  RepState* d_states = nullptr;
  // per-replica parameters
  std::vector<double> mean_new, call_dur, call_dur_hoff, hoff_prob;
  std::vector<float> a_net, a_avg, a_grad;
  double *d_mean_new = nullptr, *d_call_dur = nullptr, *d_call_dur_hoff = nullptr, *d_hoff_prob = nullptr;
  float *d_a_net = nullptr, *d_a_avg = nullptr, *d_a_grad = nullptr;
  // tapes
  std::vector<std::vector<unsigned long long>> tape_w;
  std::vector<std::vector<double>> tape_l;
  bool tape_dirty = false;
  unsigned long long* d_tape_w = nullptr;
  double* d_tape_l = nullptr;
  unsigned long long* d_tape_off = nullptr;
  TraceRec* d_trace = nullptr;
  long long* d_sum = nullptr;
  Summary* d_summary = nullptr;
  HostView* d_view = nullptr;
  StepOut* d_stepout = nullptr;
  float* d_agents = nullptr;  // dca_read_agents: [2][chunk][3479] + [chunk] floats, allocated at first use
  unsigned* d_units = nullptr;  // FAST run kernel: chunks of replica r done in the current launch
  std::string err;
  float last_ms = 0.f;
  bool pending = false;  // dca_run_async enqueued a launch that dca_wait has not collected yet
  long long launches = 0;
  int warps = 4;
  OpScratch scratch;  // staging of dca_write_state (device + pinned host), allocated at first use
};

namespace {

int fail(dca_handle* h, int code, const char* msg) {
  g_last_error = msg;
  if (h) h->err = msg;
  return code;
}

// a launch of dca_run_async still in flight is collected before anything else touches the handle
#define SETTLE(h)                                   \
  do {                                              \
    if ((h) && (h)->pending) {                      \
      const int rc_settle_ = dca_wait(h);           \
      if (rc_settle_) return rc_settle_;            \
    }                                               \
  } while (0)

int scratch_alloc(dca_handle* h, OpScratch& s, int device, cudaStream_t stream);
int scratch_reserve(dca_handle* h) { return scratch_alloc(h, h->scratch, h->device, h->stream); }

int upload_tables(dca_handle* h, int device) {
  std::lock_guard<std::mutex> lk(g_tab_mutex);
  if (device < 64 && g_tab_uploaded[device]) return DCA_OK;
  Tables t;
  build_tables(t);
  CK(cudaMemcpyToSymbol(c_tab, &t, sizeof t));
  if (device < 64) g_tab_uploaded[device] = true;
  return DCA_OK;
}

int set_device(dca_handle* h, int device) {
  CK(cudaSetDevice(device));
  return upload_tables(h, device);
}

template <typename T>
int up(dca_handle* h, T** dptr, const std::vector<T>& v) {
  if (!*dptr) CK(cudaMalloc(dptr, sizeof(T) * v.size()));
  CK(cudaMemcpyAsync(*dptr, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, h->stream));
  return DCA_OK;
}

int sync_tapes(dca_handle* h) {
  if (h->cfg.rng_mode != DCA_RNG_TAPE || !h->tape_dirty) return DCA_OK;
  std::vector<unsigned long long> off(h->n + 1, 0), words;
  std::vector<double> nl;
  for (int i = 0; i < h->n; ++i) {
    off[i] = words.size();
    words.insert(words.end(), h->tape_w[i].begin(), h->tape_w[i].end());
    nl.insert(nl.end(), h->tape_l[i].begin(), h->tape_l[i].end());
  }
  off[h->n] = words.size();
  CK(cudaStreamSynchronize(h->stream));
  if (h->d_tape_w) { cudaFree(h->d_tape_w); h->d_tape_w = nullptr; }
  if (h->d_tape_l) { cudaFree(h->d_tape_l); h->d_tape_l = nullptr; }
  size_t n = words.size() ? words.size() : 1;
  CK(cudaMalloc(&h->d_tape_w, n * sizeof(unsigned long long)));
  CK(cudaMalloc(&h->d_tape_l, n * sizeof(double)));
  if (!h->d_tape_off) CK(cudaMalloc(&h->d_tape_off, (h->n + 1) * sizeof(unsigned long long)));
  if (!words.empty()) {
    CK(cudaMemcpy(h->d_tape_w, words.data(), words.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_tape_l, nl.data(), nl.size() * 8, cudaMemcpyHostToDevice));
  }
  CK(cudaMemcpy(h->d_tape_off, off.data(), off.size() * 8, cudaMemcpyHostToDevice));
  h->tape_dirty = false;
  return DCA_OK;
}

int launch_init(dca_handle* h, int only_replica) {
  int rc = sync_tapes(h);
  if (rc) return rc;
  InitArgs a{};
  a.states = h->d_states;
  a.n_replicas = h->n;
  a.rng_mode = h->cfg.rng_mode == DCA_RNG_PHILOX ? RNG_PHILOX : RNG_TAPE;
  a.key0 = (uint32_t)h->cfg.seed;
  a.key1 = (uint32_t)(h->cfg.seed >> 32);
  a.first_replica = h->cfg.first_replica;
  a.n_events = h->cfg.n_events;
  a.min_loss = h->cfg.min_loss;
  a.call_dur = h->d_call_dur;
  a.call_dur_hoff = h->d_call_dur_hoff;
  a.mean_new = h->d_mean_new;
  a.hoff_prob = h->d_hoff_prob;
  a.alpha_net = h->d_a_net;
  a.alpha_avg = h->d_a_avg;
  a.alpha_grad = h->d_a_grad;
  a.tape_words = h->d_tape_w;
  a.tape_neglog = h->d_tape_l;
  a.tape_off = h->d_tape_off;
  a.only_replica = only_replica;
  int blocks = only_replica >= 0 ? 1 : (h->n + 3) / 4;
  k_init<<<blocks, 128, 0, h->stream>>>(a);
  h->launches++;
  CK(cudaGetLastError());
  return DCA_OK;
}

bool bad_replica(const dca_handle* h, int r) { return !h || r < 0 || r >= h->n; }

// events per unit of the FAST run kernel (see k_run_fast): small enough that the launch ends within a fraction of a
// per cent of the ideal time, large enough that staging a replica in and out (100 KB) is noise
#ifndef DCA_FAST_CHUNK_EVENTS
#define DCA_FAST_CHUNK_EVENTS 4096
#endif
constexpr long long kFastChunkEvents = DCA_FAST_CHUNK_EVENTS;

template <typename K>
int set_smem(dca_handle* h, K kernel, size_t bytes = sizeof(Smem)) {
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return DCA_OK;
}

}  // namespace

extern "C" {

int dca_config_defaults(dca_config* cfg) {  // Opt defaults, Opt.hs:28-76
  if (!cfg) return DCA_ERR_ARG;
  memset(cfg, 0, sizeof *cfg);
  cfg->n_replicas = 1;
  cfg->device = 0;
  cfg->seed = 0;
  cfg->first_replica = 0;
  cfg->rng_mode = DCA_RNG_PHILOX;
  cfg->order = DCA_ORDER_FAST;
  cfg->n_events = 10000;
  cfg->min_loss = 0.f;
  cfg->call_dur = 3.0;
  cfg->call_dur_hoff = 1.0;
  cfg->call_rate = 200.0;
  cfg->hoff_prob = 0.0;
  cfg->alpha_net = 2.52e-6f;
  cfg->alpha_avg = 0.06f;
  cfg->alpha_grad = 5e-6f;
  return DCA_OK;
}

int dca_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

const char* dca_last_error(const dca_handle* h) {
  return h ? h->err.c_str() : g_last_error.c_str();
}

int dca_destroy(dca_handle* h) {
  if (h && h->pending) dca_wait(h);
  if (!h) return DCA_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  h->scratch.release();
  cudaFree(h->d_states);
  cudaFree(h->d_mean_new); cudaFree(h->d_call_dur); cudaFree(h->d_call_dur_hoff); cudaFree(h->d_hoff_prob);
  cudaFree(h->d_a_net); cudaFree(h->d_a_avg); cudaFree(h->d_a_grad);
  cudaFree(h->d_tape_w); cudaFree(h->d_tape_l); cudaFree(h->d_tape_off);
  cudaFree(h->d_agents);
  cudaFree(h->d_units);
  cudaFree(h->d_trace); cudaFree(h->d_sum); cudaFree(h->d_summary); cudaFree(h->d_view); cudaFree(h->d_stepout);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return DCA_OK;
}

int dca_create(const dca_config* cfg, dca_handle** out) {
  dca_handle* h = nullptr;
  if (!cfg || !out) return fail(h, DCA_ERR_ARG, "dca_create: null argument");
  *out = nullptr;
  if (cfg->n_replicas < 1) return fail(h, DCA_ERR_ARG, "dca_create: n_replicas < 1");
  if (cfg->order != DCA_ORDER_REF && cfg->order != DCA_ORDER_FAST)
    return fail(h, DCA_ERR_ARG, "dca_create: bad order");
  if (cfg->rng_mode != DCA_RNG_TAPE && cfg->rng_mode != DCA_RNG_PHILOX)
    return fail(h, DCA_ERR_ARG, "dca_create: bad rng_mode");
  if (cfg->n_events < 0 || cfg->n_events > DCA_MAX_EVENTS)  // event ids are 32-bit on the device, up to three per event
    return fail(h, DCA_ERR_ARG, "dca_create: n_events must be in [0, DCA_MAX_EVENTS = 1.4e9] (the device keeps event ids in 32 bits)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(h, DCA_ERR_NO_DEVICE, "dca_create: no CUDA device (there is no CPU fallback)");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(h, DCA_ERR_ARG, "dca_create: bad device ordinal");
  h = new dca_handle();
  h->cfg = *cfg;
  h->device = cfg->device;
  h->n = cfg->n_replicas;
  int rc = set_device(h, h->device);
  if (rc) { g_last_error = h->err; delete h; return rc; }
  auto bail = [&](int code) { g_last_error = h->err; dca_destroy(h); return code; };
#define CKB(call)                                                                    \
  do {                                                                               \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess) {                                                         \
      h->err = std::string(#call " failed: ") + cudaGetErrorString(e_);              \
      return bail(DCA_ERR_CUDA);                                                     \
    }                                                                                \
  } while (0)
  CKB(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CKB(cudaEventCreate(&h->ev0));
  CKB(cudaEventCreate(&h->ev1));
  CKB(cudaMalloc(&h->d_states, sizeof(RepState) * (size_t)h->n));
  CKB(cudaMalloc(&h->d_sum, sizeof(long long) * 10));
  CKB(cudaMalloc(&h->d_summary, sizeof(Summary) * (size_t)h->n));
  CKB(cudaMalloc(&h->d_view, sizeof(HostView)));
  CKB(cudaMalloc(&h->d_stepout, sizeof(StepOut)));
  if (cfg->trace_capacity > 0) {
    CKB(cudaMalloc(&h->d_trace, sizeof(TraceRec) * (size_t)h->n * cfg->trace_capacity));
    CKB(cudaMemsetAsync(h->d_trace, 0, sizeof(TraceRec) * (size_t)h->n * cfg->trace_capacity, h->stream));
  }
  // per-replica parameters (Opt.hs:10-25); lam' = 1 / (lam / 60) as EventGen.hs:81-82
  auto fill_d = [&](std::vector<double>& v, const double* arr, double scalar) {
    v.resize(h->n);
    for (int i = 0; i < h->n; ++i) v[i] = arr ? arr[i] : scalar;
  };
  auto fill_f = [&](std::vector<float>& v, const float* arr, float scalar) {
    v.resize(h->n);
    for (int i = 0; i < h->n; ++i) v[i] = arr ? arr[i] : scalar;
  };
  std::vector<double> rate;
  fill_d(rate, cfg->call_rate_v, cfg->call_rate);
  h->mean_new.resize(h->n);
  for (int i = 0; i < h->n; ++i) h->mean_new[i] = 1.0 / (rate[i] / 60.0);
  fill_d(h->call_dur, cfg->call_dur_v, cfg->call_dur);
  fill_d(h->call_dur_hoff, cfg->call_dur_hoff_v, cfg->call_dur_hoff);
  fill_d(h->hoff_prob, cfg->hoff_prob_v, cfg->hoff_prob);
  fill_f(h->a_net, cfg->alpha_net_v, cfg->alpha_net);
  fill_f(h->a_avg, cfg->alpha_avg_v, cfg->alpha_avg);
  fill_f(h->a_grad, cfg->alpha_grad_v, cfg->alpha_grad);
  // the config's array pointers are caller-owned: do not keep them
  h->cfg.call_dur_v = h->cfg.call_dur_hoff_v = h->cfg.call_rate_v = h->cfg.hoff_prob_v = nullptr;
  h->cfg.alpha_net_v = h->cfg.alpha_avg_v = h->cfg.alpha_grad_v = nullptr;
  if ((rc = up(h, &h->d_mean_new, h->mean_new)) || (rc = up(h, &h->d_call_dur, h->call_dur)) ||
      (rc = up(h, &h->d_call_dur_hoff, h->call_dur_hoff)) || (rc = up(h, &h->d_hoff_prob, h->hoff_prob)) ||
      (rc = up(h, &h->d_a_net, h->a_net)) || (rc = up(h, &h->d_a_avg, h->a_avg)) ||
      (rc = up(h, &h->d_a_grad, h->a_grad)))
    return bail(rc);
  h->tape_w.resize(h->n);
  h->tape_l.resize(h->n);
  h->tape_dirty = cfg->rng_mode == DCA_RNG_TAPE;
  h->warps = 4;  // both kernels use 128-thread CTAs (warps_per_replica is accepted and ignored)
  if ((rc = set_smem(h, fast::k_run_fast<false>, sizeof(fast::SmemF))) ||
      (rc = set_smem(h, fast::k_run_fast<true>, sizeof(fast::SmemF))) ||
      (rc = set_smem(h, fast::k_run_fast<true, true>, sizeof(fast::SmemLA))) ||
      (rc = set_smem(h, fast::k_step_fast, sizeof(fast::SmemF))) ||
      (rc = set_smem(h, k_run<ORDER_REF, false>, SMEM_RUN_BYTES)) || (rc = set_smem(h, k_run<ORDER_REF, true>, SMEM_RUN_BYTES)) ||
      (rc = set_smem(h, k_step<ORDER_REF>)) || (rc = set_smem(h, k_gf_violates)))
    return bail(rc);
  CKB(cudaMalloc(&h->d_units, sizeof(unsigned) * (size_t)h->n));
  if ((rc = dca_reset(h))) return bail(rc);
  *out = h;
  return DCA_OK;
#undef CKB
}

int dca_reset(dca_handle* h) {
  SETTLE(h);
  if (!h) return DCA_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaEventRecord(h->ev0, h->stream));
  int rc = launch_init(h, -1);
  if (rc) return rc;
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
  return DCA_OK;
}

int dca_set_params(dca_handle* h, const double* call_dur, const double* call_dur_hoff,
                   const double* call_rate, const double* hoff_prob, const float* alpha_net,
                   const float* alpha_avg, const float* alpha_grad) {
  SETTLE(h);
  if (!h) return DCA_ERR_ARG;
  CK(cudaSetDevice(h->device));
  int rc = DCA_OK;
  const size_t n = (size_t)h->n;
  if (call_rate) {
    for (size_t i = 0; i < n; ++i) h->mean_new[i] = 1.0 / (call_rate[i] / 60.0);  // EventGen.hs:81-82
    if ((rc = up(h, &h->d_mean_new, h->mean_new))) return rc;
  }
  if (call_dur) { h->call_dur.assign(call_dur, call_dur + n); if ((rc = up(h, &h->d_call_dur, h->call_dur))) return rc; }
  if (call_dur_hoff) { h->call_dur_hoff.assign(call_dur_hoff, call_dur_hoff + n); if ((rc = up(h, &h->d_call_dur_hoff, h->call_dur_hoff))) return rc; }
  if (hoff_prob) { h->hoff_prob.assign(hoff_prob, hoff_prob + n); if ((rc = up(h, &h->d_hoff_prob, h->hoff_prob))) return rc; }
  if (alpha_net) { h->a_net.assign(alpha_net, alpha_net + n); if ((rc = up(h, &h->d_a_net, h->a_net))) return rc; }
  if (alpha_avg) { h->a_avg.assign(alpha_avg, alpha_avg + n); if ((rc = up(h, &h->d_a_avg, h->a_avg))) return rc; }
  if (alpha_grad) { h->a_grad.assign(alpha_grad, alpha_grad + n); if ((rc = up(h, &h->d_a_grad, h->a_grad))) return rc; }
  CK(cudaStreamSynchronize(h->stream));
  return DCA_OK;
}

int dca_set_rng_tape(dca_handle* h, int32_t replica, const uint64_t* words, const double* neglog,
                     size_t n) {
  SETTLE(h);
  if (bad_replica(h, replica) || !words || !neglog) return fail(h, DCA_ERR_ARG, "dca_set_rng_tape: bad argument");
  if (h->cfg.rng_mode != DCA_RNG_TAPE) return fail(h, DCA_ERR_STATE, "dca_set_rng_tape: handle is not in DCA_RNG_TAPE mode");
  CK(cudaSetDevice(h->device));
  h->tape_w[replica].assign(words, words + n);
  h->tape_l[replica].assign(neglog, neglog + n);
  h->tape_dirty = true;
  int rc = launch_init(h, replica);
  if (rc) return rc;
  CK(cudaStreamSynchronize(h->stream));
  return DCA_OK;
}

int dca_set_rng_mt19937(dca_handle* h, int32_t replica, uint64_t seed, size_t n_draws) {
  SETTLE(h);
  if (bad_replica(h, replica)) return fail(h, DCA_ERR_ARG, "dca_set_rng_mt19937: bad replica");
  // pureMT seed (mersenne-random-pure64) == MT19937-64 init_genrand64(seed);
  // randomDouble = (w >> 11) / 2^53; exponential = -log(u) * mean with the host libm.
  std::mt19937_64 gen(seed);
  std::vector<uint64_t> w(n_draws);
  std::vector<double> nl(n_draws);
  for (size_t i = 0; i < n_draws; ++i) {
    w[i] = gen();
    double u = (double)(w[i] >> 11) * (1.0 / 9007199254740992.0);
    nl[i] = -std::log(u);
  }
  return dca_set_rng_tape(h, replica, w.data(), nl.data(), n_draws);
}

int dca_wait(dca_handle* h) {
  if (!h) return fail(h, DCA_ERR_ARG, "dca_wait: null handle");
  if (!h->pending) return DCA_OK;
  h->pending = false;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
  return DCA_OK;
}

int dca_run_async(dca_handle* h, int64_t max_events) {
  if (!h || max_events < 0) return fail(h, DCA_ERR_ARG, "dca_run: bad argument");
  if (h->pending) {  // one launch in flight per handle: collect the previous one first
    const int rcw = dca_wait(h);
    if (rcw) return rcw;
  }
  CK(cudaSetDevice(h->device));
  if (h->cfg.rng_mode == DCA_RNG_TAPE)  // a replica without its 49 initial draws was never initialised (DCA_UNINITIALIZED)
    for (int i = 0; i < h->n; ++i)
      if (h->tape_w[i].size() < (size_t)NCELL) {
        char buf[160];
        snprintf(buf, sizeof buf, "dca_run: replica %d has no random tape (dca_set_rng_tape / dca_set_rng_mt19937 with >= 49 draws first)", i);
        return fail(h, DCA_ERR_STATE, buf);
      }
  int rc = sync_tapes(h);
  if (rc) return rc;
  RunArgs a{};
  a.states = h->d_states;
  a.n_replicas = h->n;
  a.max_events = max_events;
  a.rng_mode = h->cfg.rng_mode == DCA_RNG_PHILOX ? RNG_PHILOX : RNG_TAPE;
  a.key0 = (uint32_t)h->cfg.seed;
  a.key1 = (uint32_t)(h->cfg.seed >> 32);
  a.verify_rc = h->cfg.verify_reuse_constraint;
  a.verify_nb = h->cfg.verify_nelig_bounds;
  a.tape_words = h->d_tape_w;
  a.tape_neglog = h->d_tape_l;
  a.tape_off = h->d_tape_off;
  a.trace = h->d_trace;
  a.trace_cap = h->cfg.trace_capacity;
  a.lookahead = h->cfg.hoff_lookahead;
  const bool trace = h->d_trace != nullptr;
  // (the TRACE instantiation of the fused FAST kernel also serves the verify flags: it has the extra per-event barrier)
  const bool trace_fast = trace || a.verify_rc || a.verify_nb;
  const dim3 grid(h->n), block(128);
  const size_t smem = SMEM_RUN_BYTES;  // (REF-order run kernel: RepState + Scratch without the step kernel's prod_g)
  CK(cudaEventRecord(h->ev0, h->stream));
  if (h->cfg.order == DCA_ORDER_FAST) {
    // one CTA per unit (replica, chunk of kFastChunkEvents events), numbered chunk-major
    a.chunk_events = kFastChunkEvents;
    const long long chunks = std::max<long long>(1, (max_events + kFastChunkEvents - 1) / kFastChunkEvents);
    if (chunks * (long long)h->n > 0x7fffffffLL) return fail(h, DCA_ERR_ARG, "dca_run: max_events x n_replicas too large for one launch");
    a.n_chunks = (unsigned)chunks;
    a.chunk_done = h->d_units;
    CK(cudaMemsetAsync(h->d_units, 0, sizeof(unsigned) * (size_t)h->n, h->stream));
    const dim3 fgrid((unsigned)(chunks * (long long)h->n));
    // (hand-off look-ahead: its own instantiation, so that the two hot ones carry none of its code; it has the extra barrier too)
    if (a.lookahead) fast::k_run_fast<true, true><<<fgrid, fast::NT, sizeof(fast::SmemLA), h->stream>>>(a);
    else if (trace_fast) fast::k_run_fast<true><<<fgrid, fast::NT, sizeof(fast::SmemF), h->stream>>>(a);
    else fast::k_run_fast<false><<<fgrid, fast::NT, sizeof(fast::SmemF), h->stream>>>(a);
  } else {
    if (trace) k_run<ORDER_REF, true><<<grid, block, smem, h->stream>>>(a);
    else k_run<ORDER_REF, false><<<grid, block, smem, h->stream>>>(a);
  }
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaEventRecord(h->ev1, h->stream));
  h->pending = true;
  return DCA_OK;
}

int dca_run(dca_handle* h, int64_t max_events) {
  const int rc = dca_run_async(h, max_events);
  return rc ? rc : dca_wait(h);
}

// launch on every handle (typically one per GPU of the box), then collect them all: the devices run side by side
// although the caller is a single thread
int dca_run_many(dca_handle** handles, int32_t n, int64_t max_events) {
  dca_handle* h = nullptr;
  if (!handles || n <= 0) return fail(h, DCA_ERR_ARG, "dca_run_many: bad argument");
  for (int i = 0; i < n; ++i)
    if (!handles[i]) return fail(h, DCA_ERR_ARG, "dca_run_many: null handle");
  int first = DCA_OK;
  for (int i = 0; i < n; ++i) {
    const int rc = dca_run_async(handles[i], max_events);
    if (rc && !first) first = rc;
  }
  for (int i = 0; i < n; ++i) {
    const int rc = dca_wait(handles[i]);
    if (rc && !first) first = rc;
  }
  return first;
}

int64_t dca_run_unit_events(void) { return kFastChunkEvents; }

int dca_elapsed_ms(const dca_handle* h, float* ms) {
  if (!h || !ms) return DCA_ERR_ARG;
  *ms = h->last_ms;
  return DCA_OK;
}

int dca_last_launch_count(const dca_handle* h, int64_t* n) {
  if (!h || !n) return DCA_ERR_ARG;
  *n = h->launches;
  return DCA_OK;
}

static int step_launch(dca_handle* h, int replica, int mode, int r, int c, int is_end, int ch,
                       float an, float aa, float ag, StepOut* host_out) {
  if (bad_replica(h, replica) || r < 0 || r >= ROWS || c < 0 || c >= COLS || ch >= NCH)
    return fail(h, DCA_ERR_ARG, "step: bad argument");
  CK(cudaSetDevice(h->device));
  const int cell = r * COLS + c;
  if (h->cfg.order == DCA_ORDER_FAST)
    fast::k_step_fast<<<1, fast::NT, sizeof(fast::SmemF), h->stream>>>(h->d_states, replica, mode, cell, is_end, ch, an, aa, ag, h->d_stepout);
  else
    k_step<ORDER_REF><<<1, 128, sizeof(Smem), h->stream>>>(h->d_states, replica, mode, cell, is_end, ch, an, aa, ag, h->d_stepout);
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(host_out, h->d_stepout, sizeof(StepOut), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (host_out->ch == STEP_BAD_CHANNEL)
    return fail(h, DCA_ERR_ARG, is_end ? "dca_grid_step_train: the channel is not in use at this cell (nothing was changed)"
                                       : "dca_grid_step_train: the channel is not eligible at this cell (nothing was changed)");
  return DCA_OK;
}

int dca_get_action(dca_handle* h, int32_t replica, int32_t r, int32_t c, int32_t is_end, int32_t* ch) {
  SETTLE(h);
  if (!ch) return fail(h, DCA_ERR_ARG, "dca_get_action: null output");
  StepOut o{};
  int rc = step_launch(h, replica, 0, r, c, is_end, -1, 0.f, 0.f, 0.f, &o);
  if (rc) return rc;
  *ch = o.ch;
  return DCA_OK;
}

int dca_grid_step_train(dca_handle* h, int32_t replica, float alpha_net, float alpha_avg,
                        float alpha_grad, int32_t is_end, int32_t r, int32_t c, int32_t ch,
                        float* loss, int32_t* loss_is_nan, int64_t* reward) {
  SETTLE(h);
  StepOut o{};
  int rc = step_launch(h, replica, 1, r, c, is_end, ch, alpha_net, alpha_avg, alpha_grad, &o);
  if (rc) return rc;
  if (loss) *loss = o.loss;
  if (loss_is_nan) *loss_is_nan = o.loss_is_nan;
  if (reward) *reward = o.reward;
  return DCA_OK;
}

int dca_read_state(dca_handle* h, int32_t replica, uint8_t* grid, int64_t* frep, float* wnet,
                   float* wgrad, float* avg_reward, int64_t* iter, dca_event* next_event) {
  SETTLE(h);
  if (bad_replica(h, replica)) return fail(h, DCA_ERR_ARG, "dca_read_state: bad replica");
  CK(cudaSetDevice(h->device));
  k_unpack<<<1, 256, 0, h->stream>>>(h->d_states + replica, h->d_view);
  h->launches++;
  CK(cudaGetLastError());
  std::vector<unsigned char> hv(sizeof(HostView));
  CK(cudaMemcpyAsync(hv.data(), h->d_view, sizeof(HostView), cudaMemcpyDeviceToHost, h->stream));
  // scalars: copy the tail of the blob
  const size_t tail_off = offsetof(RepState, ev_time);
  std::vector<unsigned char> tail(sizeof(RepState) - tail_off);
  CK(cudaMemcpyAsync(tail.data(), reinterpret_cast<const unsigned char*>(h->d_states + replica) + tail_off,
                     tail.size(), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const HostView* v = reinterpret_cast<const HostView*>(hv.data());
  if (grid) memcpy(grid, v->grid, sizeof v->grid);
  if (frep) memcpy(frep, v->frep, sizeof v->frep);
  if (wnet) memcpy(wnet, v->w, sizeof v->w);
  if (wgrad) memcpy(wgrad, v->wg, sizeof v->wg);
  auto field = [&](size_t off, void* dst, size_t sz) { memcpy(dst, tail.data() + (off - tail_off), sz); };
  if (avg_reward) field(offsetof(RepState, avg_reward), avg_reward, sizeof(float));
  if (iter) { long long it; field(offsetof(RepState, iter), &it, sizeof it); *iter = it; }
  if (next_event) {
    int type, cell, ch, hoff;
    field(offsetof(RepState, ev_time), &next_event->time, sizeof(double));
    field(offsetof(RepState, ev_type), &type, 4);
    field(offsetof(RepState, ev_cell), &cell, 4);
    field(offsetof(RepState, ev_ch), &ch, 4);
    field(offsetof(RepState, ev_hoff), &hoff, 4);
    next_event->etype = type;
    next_event->r = cell / COLS;
    next_event->c = cell % COLS;
    next_event->ch = type == EV_END ? ch : -1;
    next_event->hoff_r = (type == EV_END && hoff != HOFF_NONE) ? hoff / COLS : -1;
    next_event->hoff_c = (type == EV_END && hoff != HOFF_NONE) ? hoff % COLS : -1;
  }
  return DCA_OK;
}

int dca_write_state(dca_handle* h, int32_t replica, const uint8_t* grid, const float* wnet,
                    const float* wgrad, const float* avg_reward) {
  SETTLE(h);
  if (bad_replica(h, replica)) return fail(h, DCA_ERR_ARG, "dca_write_state: bad replica");
  CK(cudaSetDevice(h->device));
  RepState* st = h->d_states + replica;
  int rc = scratch_reserve(h);
  if (rc) return rc;
  OpScratch& sc = h->scratch;
  OpIO *hi = sc.h_io, *di = sc.d_io;
  if (grid) memcpy(hi->grid, grid, DCA_GRID_SIZE);
  if (wnet) memcpy(hi->w, wnet, FREP * 4);
  if (wgrad) memcpy(hi->wg, wgrad, FREP * 4);
  CK(cudaMemcpyAsync(di, hi, offsetof(OpIO, frep), cudaMemcpyHostToDevice, h->stream));  // grid, w, wg in one copy
  if (grid) {
    k_gf_load_grid<<<1, 256, 0, h->stream>>>(st, di->grid);
    h->launches++;
    CK(cudaGetLastError());
  }
  if (wnet || wgrad) {
    k_pack_weights<<<1, 256, 0, h->stream>>>(st, wnet ? di->w : nullptr, wgrad ? di->wg : nullptr);
    h->launches++;
    CK(cudaGetLastError());
  }
  CK(cudaStreamSynchronize(h->stream));
  if (avg_reward)
    CK(cudaMemcpy(reinterpret_cast<unsigned char*>(st) + offsetof(RepState, avg_reward), avg_reward, 4, cudaMemcpyHostToDevice));
  CK(cudaGetLastError());
  return DCA_OK;
}

// ---- checkpoint / resume: the replica blobs verbatim behind a small header ----------
namespace {
struct CkptHeader {
  uint64_t magic;      // "DCAB200\2"
  uint32_t version;    // layout version of this header + RepState
  uint32_t blob_bytes; // sizeof(RepState)
  int32_t n_replicas;
  int32_t order, rng_mode;
  int32_t pad;
  uint64_t seed;          // Philox key (the stream continues only under the same key) ...
  uint64_t first_replica; // ... and the same stream ids
  uint64_t tape_hash;     // DCA_RNG_TAPE: hash over every replica's tape length and words (0 in Philox mode)
};
constexpr uint64_t kCkptMagic = 0x0230303242414344ULL;
constexpr uint32_t kCkptVersion = 2;

uint64_t tape_hash(const dca_handle* h) {
  if (h->cfg.rng_mode != DCA_RNG_TAPE) return 0;
  uint64_t x = 0xCBF29CE484222325ULL;  // FNV-1a over (length, words) of every replica
  auto mixin = [&](uint64_t v) {
    for (int b = 0; b < 8; ++b) { x ^= (v >> (8 * b)) & 0xFF; x *= 0x100000001B3ULL; }
  };
  for (int i = 0; i < h->n; ++i) {
    mixin(h->tape_w[i].size());
    for (unsigned long long w : h->tape_w[i]) mixin(w);
  }
  return x ? x : 1;
}
}  // namespace

size_t dca_checkpoint_size(const dca_handle* h) {
  return h ? sizeof(CkptHeader) + sizeof(RepState) * (size_t)h->n : 0;
}

int dca_checkpoint_save(dca_handle* h, void* buf, size_t size) {
  SETTLE(h);
  if (!h || !buf || size < dca_checkpoint_size(h)) return fail(h, DCA_ERR_ARG, "dca_checkpoint_save: bad argument");
  CK(cudaSetDevice(h->device));
  CkptHeader hd{kCkptMagic, kCkptVersion, (uint32_t)sizeof(RepState), h->n, h->cfg.order, h->cfg.rng_mode, 0,
                h->cfg.seed, h->cfg.first_replica, tape_hash(h)};
  memcpy(buf, &hd, sizeof hd);
  CK(cudaMemcpyAsync(static_cast<char*>(buf) + sizeof hd, h->d_states, sizeof(RepState) * (size_t)h->n,
                     cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return DCA_OK;
}

int dca_checkpoint_load(dca_handle* h, const void* buf, size_t size) {
  SETTLE(h);
  if (!h || !buf || size < dca_checkpoint_size(h)) return fail(h, DCA_ERR_ARG, "dca_checkpoint_load: bad argument");
  CkptHeader hd;
  memcpy(&hd, buf, sizeof hd);
  if (hd.magic != kCkptMagic || hd.version != kCkptVersion || hd.blob_bytes != sizeof(RepState) || hd.n_replicas != h->n ||
      hd.order != h->cfg.order || hd.rng_mode != h->cfg.rng_mode)
    return fail(h, DCA_ERR_STATE, "dca_checkpoint_load: the blob does not match this handle (replicas, order, rng mode or layout version)");
  // the random stream lives in the handle (Philox key and stream ids, or the tapes), not in the blob: continuing
  // "bit for bit" needs the same one
  if (hd.seed != h->cfg.seed || hd.first_replica != h->cfg.first_replica)
    return fail(h, DCA_ERR_STATE, "dca_checkpoint_load: the blob was saved under a different seed / first_replica");
  if (hd.tape_hash != tape_hash(h))
    return fail(h, DCA_ERR_STATE, "dca_checkpoint_load: the blob was saved with different random tapes");
  CK(cudaSetDevice(h->device));
  int rc = sync_tapes(h);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->d_states, static_cast<const char*>(buf) + sizeof hd, sizeof(RepState) * (size_t)h->n,
                     cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return DCA_OK;
}

static int fetch_summary(dca_handle* h, std::vector<Summary>& out) {
  CK(cudaSetDevice(h->device));
  k_summary<<<(h->n + 127) / 128, 128, 0, h->stream>>>(h->d_states, h->n, h->d_summary);
  h->launches++;
  CK(cudaGetLastError());
  out.resize(h->n);
  CK(cudaMemcpyAsync(out.data(), h->d_summary, sizeof(Summary) * (size_t)h->n, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return DCA_OK;
}

static int fetch_one(dca_handle* h, int replica, Summary& s) {
  CK(cudaSetDevice(h->device));
  k_summary<<<1, 1, 0, h->stream>>>(h->d_states + replica, 1, h->d_summary);
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(&s, h->d_summary, sizeof(Summary), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return DCA_OK;
}

int dca_read_stats(dca_handle* h, int32_t replica, int64_t stats[DCA_N_STATS]) {
  SETTLE(h);
  if (bad_replica(h, replica) || !stats) return fail(h, DCA_ERR_ARG, "dca_read_stats: bad argument");
  Summary s;
  int rc = fetch_one(h, replica, s);
  if (rc) return rc;
  for (int k = 0; k < 8; ++k) stats[k] = s.stats[k];
  return DCA_OK;
}

int dca_read_status(dca_handle* h, int32_t replica, int32_t* status, int64_t* iter, uint64_t* n_draws) {
  SETTLE(h);
  if (bad_replica(h, replica)) return fail(h, DCA_ERR_ARG, "dca_read_status: bad replica");
  Summary s;
  int rc = fetch_one(h, replica, s);
  if (rc) return rc;
  if (status) *status = s.status;
  if (iter) *iter = s.iter;
  if (n_draws) *n_draws = s.ndraws;
  return DCA_OK;
}

int dca_read_period(dca_handle* h, int32_t replica, float* loss_sum, int64_t* n, int32_t reset) {
  SETTLE(h);
  if (bad_replica(h, replica)) return fail(h, DCA_ERR_ARG, "dca_read_period: bad replica");
  Summary s;
  int rc = fetch_one(h, replica, s);
  if (rc) return rc;
  if (loss_sum) *loss_sum = s.loss_sum;
  if (n) *n = s.loss_n;
  if (reset) {
    k_period_reset<<<1, 1, 0, h->stream>>>(h->d_states, replica);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
  }
  return DCA_OK;
}

int dca_read_trace(dca_handle* h, int32_t replica, dca_trace* out, size_t cap, size_t* n) {
  SETTLE(h);
  if (bad_replica(h, replica) || !out) return fail(h, DCA_ERR_ARG, "dca_read_trace: bad argument");
  if (!h->d_trace) return fail(h, DCA_ERR_STATE, "dca_read_trace: handle created with trace_capacity == 0");
  Summary s;
  int rc = fetch_one(h, replica, s);
  if (rc) return rc;
  size_t have = (size_t)std::min<long long>(s.iter, h->cfg.trace_capacity);
  size_t cnt = std::min(have, cap);
  CK(cudaMemcpy(out, h->d_trace + (size_t)replica * h->cfg.trace_capacity, cnt * sizeof(dca_trace), cudaMemcpyDeviceToHost));
  if (n) *n = cnt;
  return DCA_OK;
}

int dca_read_summary(dca_handle* h, int32_t* status, int64_t* iter, float* avg_reward, int64_t* stats) {
  SETTLE(h);
  if (!h) return DCA_ERR_ARG;
  std::vector<Summary> s;
  int rc = fetch_summary(h, s);
  if (rc) return rc;
  for (int i = 0; i < h->n; ++i) {
    if (status) status[i] = s[i].status;
    if (iter) iter[i] = s[i].iter;
    if (avg_reward) avg_reward[i] = s[i].avg_reward;
    if (stats)
      for (int k = 0; k < 8; ++k) stats[(size_t)i * 8 + k] = s[i].stats[k];
  }
  return DCA_OK;
}

int dca_read_agents(dca_handle* h, float* wnet, float* wgrad, float* avg_reward) {
  SETTLE(h);
  if (!h) return DCA_ERR_ARG;
  CK(cudaSetDevice(h->device));
  const int chunk = 4096;
  const size_t per = (size_t)chunk * FREP;
  if (!h->d_agents) CK(cudaMalloc(&h->d_agents, sizeof(float) * (2 * per + chunk)));
  float *dw = h->d_agents, *dg = h->d_agents + per, *da = h->d_agents + 2 * per;
  for (int first = 0; first < h->n; first += chunk) {
    const int cnt = std::min(chunk, h->n - first);
    k_unpack_agents<<<cnt, 256, 0, h->stream>>>(h->d_states, first, cnt, wnet ? dw : nullptr, wgrad ? dg : nullptr,
                                                avg_reward ? da : nullptr);
    h->launches++;
    CK(cudaGetLastError());
    if (wnet) CK(cudaMemcpyAsync(wnet + (size_t)first * FREP, dw, sizeof(float) * (size_t)cnt * FREP, cudaMemcpyDeviceToHost, h->stream));
    if (wgrad) CK(cudaMemcpyAsync(wgrad + (size_t)first * FREP, dg, sizeof(float) * (size_t)cnt * FREP, cudaMemcpyDeviceToHost, h->stream));
    if (avg_reward) CK(cudaMemcpyAsync(avg_reward + first, da, sizeof(float) * cnt, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));  // (the device buffers are reused by the next chunk)
  }
  return DCA_OK;
}

int dca_sum_stats(dca_handle* h, int64_t out[10], void** dev_ptr) {
  SETTLE(h);
  if (!h) return DCA_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaMemsetAsync(h->d_sum, 0, sizeof(long long) * 10, h->stream));
  int blocks = std::min((h->n + 255) / 256, 148);
  k_sum_stats<<<blocks, 256, 0, h->stream>>>(h->d_states, h->n, h->d_sum);
  h->launches++;
  CK(cudaGetLastError());
  if (out) {
    long long tmp[10];
    CK(cudaMemcpyAsync(tmp, h->d_sum, sizeof tmp, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int k = 0; k < 10; ++k) out[k] = tmp[k];
  } else {
    CK(cudaStreamSynchronize(h->stream));
  }
  if (dev_ptr) *dev_ptr = h->d_sum;
  return DCA_OK;
}

void dca_stats_cumus(const int64_t st[DCA_N_STATS], double out[3]) {  // Stats.hs:83-92
  double rejNew = (double)st[DCA_ST_REJECTED_NEW], arrNew = (double)st[DCA_ST_ARRIVALS_NEW];
  double rejHoff = (double)st[DCA_ST_REJECTED_HOFF], arrHoff = (double)st[DCA_ST_ARRIVALS_HOFF];
  out[0] = rejNew / (arrNew + 1.0);
  out[1] = rejHoff / (arrHoff + 1.0);
  out[2] = (rejNew + rejHoff) / (arrNew + arrHoff + 1.0);
}

// ---- stateless Gridneighs / Gridfuncs operators ------------------------------
int dca_get_neighs(int32_t d, int32_t r, int32_t c, int32_t include_self, int32_t* out_rc, int32_t* n) {
  if (d < 1 || d > 4 || r < 0 || r >= ROWS || c < 0 || c >= COLS || !out_rc || !n) return DCA_ERR_ARG;
  int buf[2 * 64];
  int k = neighs(d, r, c, include_self != 0, buf);
  for (int i = 0; i < 2 * k; ++i) out_rc[i] = buf[i];
  *n = k;
  return DCA_OK;
}

}  // extern "C"

// The pure operators work on caller-owned HOST arrays.  Each device keeps ONE scratch for them, allocated at the
// first call and reused: a device block, a pinned host mirror of its host-visible part, and a stream.  A call
// copies its inputs into the pinned mirror, issues one H2D copy, its kernels and one D2H copy on that stream and
// synchronises once -- no cudaMalloc / cudaFree and no mid-pipeline round trip per call (the candidate count of
// getAction stays on the device).
namespace {

OpScratch g_dev_scratch[64];

int scratch_alloc(dca_handle* h, OpScratch& s, int device, cudaStream_t stream) {
  if (s.ready) return DCA_OK;
  s.device = device;
  if (stream) { s.stream = stream; s.own_stream = false; }
  else { CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking)); s.own_stream = true; }
  CK(cudaMalloc(&s.d_io, sizeof(OpIO)));
  CK(cudaMallocHost(&s.h_io, sizeof(OpIO)));
  CK(cudaMalloc(&s.st, sizeof(RepState)));
  CK(cudaMalloc(&s.afreps, sizeof(long long) * (size_t)NCH * FREP));
  CK(cudaMalloc(&s.q, sizeof(float) * 72));
  CK(cudaMemsetAsync(s.d_io, 0, sizeof(OpIO), s.stream));
  memset(s.h_io, 0, sizeof(OpIO));
  s.ready = true;
  return DCA_OK;
}

// the per-device scratch of the stateless operators (caller holds s.mu)
int dev_scratch(int device, OpScratch** out) {
  dca_handle* h = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(h, DCA_ERR_NO_DEVICE, "no CUDA device (there is no CPU fallback)");
  if (device < 0 || device >= ndev || device >= 64) return fail(h, DCA_ERR_ARG, "bad device ordinal");
  int rc = set_device(h, device);
  if (rc) return rc;
  OpScratch& s = g_dev_scratch[device];
  rc = scratch_alloc(h, s, device, nullptr);
  if (rc) { s.release(); return rc; }
  *out = &s;
  return DCA_OK;
}

#define IO_OFF(field) offsetof(OpIO, field)
// host -> device / device -> host of the byte range [from, to) of the OpIO block
int io_h2d(OpScratch& s, size_t from, size_t to) {
  dca_handle* h = nullptr;
  CK(cudaMemcpyAsync(reinterpret_cast<char*>(s.d_io) + from, reinterpret_cast<char*>(s.h_io) + from, to - from,
                     cudaMemcpyHostToDevice, s.stream));
  return DCA_OK;
}
int io_d2h(OpScratch& s, size_t from, size_t to) {
  dca_handle* h = nullptr;
  CK(cudaMemcpyAsync(reinterpret_cast<char*>(s.h_io) + from, reinterpret_cast<char*>(s.d_io) + from, to - from,
                     cudaMemcpyDeviceToHost, s.stream));
  return DCA_OK;
}
int io_sync(OpScratch& s) {
  dca_handle* h = nullptr;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(s.stream));
  return DCA_OK;
}

// grid bytes -> masks, cnt2, frep in the scratch blob (the caller's grid is already in h_io->grid)
void launch_load_grid(OpScratch& s) { k_gf_load_grid<<<1, 256, 0, s.stream>>>(s.st, s.d_io->grid); }

int gf_chs(int device, const uint8_t* grid, int r, int c, int is_end, int64_t* chs, int32_t* n) {
  dca_handle* h = nullptr;
  if (!grid || r < 0 || r >= ROWS || c < 0 || c >= COLS || !chs || !n) return fail(h, DCA_ERR_ARG, "gridfuncs: bad argument");
  OpScratch* sp = nullptr;
  int rc = dev_scratch(device, &sp);
  if (rc) return rc;
  OpScratch& s = *sp;
  std::lock_guard<std::mutex> lk(s.mu);
  memcpy(s.h_io->grid, grid, DCA_GRID_SIZE);
  if ((rc = io_h2d(s, IO_OFF(grid), IO_OFF(w)))) return rc;
  launch_load_grid(s);
  k_gf_chs<<<1, 32, 0, s.stream>>>(s.st, r * COLS + c, is_end, s.d_io->chs, &s.d_io->n);
  if ((rc = io_d2h(s, IO_OFF(chs), IO_OFF(out_frep))) || (rc = io_d2h(s, IO_OFF(n), IO_OFF(ch))) || (rc = io_sync(s))) return rc;
  const int cnt = s.h_io->n;
  for (int i = 0; i < cnt; ++i) chs[i] = s.h_io->chs[i];
  *n = cnt;
  return DCA_OK;
}
}  // namespace

extern "C" {

int dca_gf_eligible_chs(int32_t device, const uint8_t* grid, int32_t r, int32_t c, int64_t* chs, int32_t* n) {
  return gf_chs(device, grid, r, c, 0, chs, n);
}
int dca_gf_inuse_chs(int32_t device, const uint8_t* grid, int32_t r, int32_t c, int64_t* chs, int32_t* n) {
  return gf_chs(device, grid, r, c, 1, chs, n);
}

int dca_gf_feature_rep(int32_t device, const uint8_t* grid, int64_t* frep) {
  dca_handle* h = nullptr;
  if (!grid || !frep) return fail(h, DCA_ERR_ARG, "dca_gf_feature_rep: null argument");
  OpScratch* sp = nullptr;
  int rc = dev_scratch(device, &sp);
  if (rc) return rc;
  OpScratch& s = *sp;
  std::lock_guard<std::mutex> lk(s.mu);
  memcpy(s.h_io->grid, grid, DCA_GRID_SIZE);
  if ((rc = io_h2d(s, IO_OFF(grid), IO_OFF(w)))) return rc;
  launch_load_grid(s);
  k_gf_store_frep<<<1, 256, 0, s.stream>>>(s.st, s.d_io->out_frep);
  if ((rc = io_d2h(s, IO_OFF(out_frep), IO_OFF(reward))) || (rc = io_sync(s))) return rc;
  static_assert(sizeof(long long) == sizeof(int64_t), "int64");
  memcpy(frep, s.h_io->out_frep, sizeof(int64_t) * FREP);
  return DCA_OK;
}

int dca_gf_inc_afterstate_freps(int32_t device, int32_t r, int32_t c, int32_t is_end, const int64_t* chs,
                                int32_t n, const uint8_t* grid, const int64_t* frep, int64_t* out) {
  dca_handle* h = nullptr;
  if (r < 0 || r >= ROWS || c < 0 || c >= COLS || n < 0 || n > NCH || !grid || !frep || !out || (n && !chs))
    return fail(h, DCA_ERR_ARG, "dca_gf_inc_afterstate_freps: bad argument");
  if (n == 0) return DCA_OK;
  for (int i = 0; i < n; ++i)
    if (chs[i] < 0 || chs[i] >= NCH) return fail(h, DCA_ERR_ARG, "dca_gf_inc_afterstate_freps: channel out of range");
  OpScratch* sp = nullptr;
  int rc = dev_scratch(device, &sp);
  if (rc) return rc;
  OpScratch& s = *sp;
  std::lock_guard<std::mutex> lk(s.mu);
  memcpy(s.h_io->grid, grid, DCA_GRID_SIZE);
  memcpy(s.h_io->frep, frep, sizeof(int64_t) * FREP);
  memcpy(s.h_io->chs, chs, sizeof(int64_t) * n);
  if ((rc = io_h2d(s, IO_OFF(grid), IO_OFF(w))) || (rc = io_h2d(s, IO_OFF(frep), IO_OFF(nfrep))) ||
      (rc = io_h2d(s, IO_OFF(chs), IO_OFF(out_frep))))
    return rc;
  launch_load_grid(s);
  k_gf_inc_afterstate_freps<<<148, 256, 0, s.stream>>>(s.st, r * COLS + c, is_end, s.d_io->chs, n, s.d_io->frep, s.afreps);
  CK(cudaGetLastError());
  // (up to 1.9 MB straight into the caller's buffer: too large for the pinned mirror)
  CK(cudaMemcpyAsync(out, s.afreps, sizeof(long long) * (size_t)n * FREP, cudaMemcpyDeviceToHost, s.stream));
  return io_sync(s);
}

int dca_gf_violates_reuse_constraint(int32_t device, const uint8_t* grid, int32_t* violates) {
  dca_handle* h = nullptr;
  if (!grid || !violates) return fail(h, DCA_ERR_ARG, "dca_gf_violates_reuse_constraint: null argument");
  OpScratch* sp = nullptr;
  int rc = dev_scratch(device, &sp);
  if (rc) return rc;
  OpScratch& s = *sp;
  std::lock_guard<std::mutex> lk(s.mu);
  memcpy(s.h_io->grid, grid, DCA_GRID_SIZE);
  if ((rc = io_h2d(s, IO_OFF(grid), IO_OFF(w)))) return rc;
  launch_load_grid(s);
  CK(cudaFuncSetAttribute(k_gf_violates, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  k_gf_violates<<<1, 128, sizeof(Smem), s.stream>>>(s.st, &s.d_io->violates);
  if ((rc = io_d2h(s, IO_OFF(violates), IO_OFF(dot))) || (rc = io_sync(s))) return rc;
  *violates = s.h_io->violates;
  return DCA_OK;
}

// ---- accFn1 / accFn2 on explicit arrays (SimRunner.hs:61-62) --------------------
int dca_acc_get_action(int32_t device, int32_t order, int32_t r, int32_t c, int32_t is_end, const float* wnet,
                       const uint8_t* grid, const int64_t* frep, int32_t* ch, int64_t* next_frep) {
  dca_handle* h = nullptr;
  if (r < 0 || r >= ROWS || c < 0 || c >= COLS || !wnet || !grid || !frep || !ch || !next_frep ||
      (order != DCA_ORDER_REF && order != DCA_ORDER_FAST))
    return fail(h, DCA_ERR_ARG, "dca_acc_get_action: bad argument");
  OpScratch* sp = nullptr;
  int rc = dev_scratch(device, &sp);
  if (rc) return rc;
  OpScratch& s = *sp;
  std::lock_guard<std::mutex> lk(s.mu);
  OpIO* d = s.d_io;
  const int cell = r * COLS + c;
  memcpy(s.h_io->grid, grid, DCA_GRID_SIZE);
  memcpy(s.h_io->w, wnet, sizeof(float) * FREP);
  memcpy(s.h_io->frep, frep, sizeof(int64_t) * FREP);
  if ((rc = io_h2d(s, IO_OFF(grid), IO_OFF(wg))) || (rc = io_h2d(s, IO_OFF(frep), IO_OFF(nfrep)))) return rc;
  launch_load_grid(s);
  k_gf_chs<<<1, 32, 0, s.stream>>>(s.st, cell, is_end, d->chs, &d->n);  // eligibleChs / inuseChs, Simulator.hs:161-165
  // selectAction, Simulator.hs:185-205: afterstate freps -> q = mvMul -> argpmax; (Nothing, frep) without a candidate
  ops::k_afterstates_dev<<<148, 256, 0, s.stream>>>(s.st, cell, is_end, d->chs, &d->n, d->frep, s.afreps);
  ops::k_dense_dots_dev<<<NCH, ops::NT, 0, s.stream>>>(order, s.afreps, &d->n, d->w, s.q);
  ops::k_select_dev<<<1, 256, 0, s.stream>>>(s.q, &d->n, d->chs, s.afreps, d->frep, &d->ch, d->out_frep);
  if ((rc = io_d2h(s, IO_OFF(out_frep), IO_OFF(reward))) || (rc = io_d2h(s, IO_OFF(n), IO_OFF(dot))) || (rc = io_sync(s)))
    return rc;
  *ch = s.h_io->ch;
  memcpy(next_frep, s.h_io->out_frep, sizeof(int64_t) * FREP);
  return DCA_OK;
}

int dca_acc_grid_step_train(int32_t device, int32_t order, float alpha_net, float alpha_avg, float alpha_grad,
                            int32_t is_end, int32_t r, int32_t c, int32_t ch, const int64_t* frep,
                            const int64_t* next_frep, uint8_t* grid, float* wnet, float* wgrad, float* avg_reward,
                            float* loss, int32_t* loss_is_nan, int64_t* reward) {
  dca_handle* h = nullptr;
  if (r < 0 || r >= ROWS || c < 0 || c >= COLS || ch >= NCH || !frep || !next_frep || !grid || !wnet || !wgrad ||
      !avg_reward || (order != DCA_ORDER_REF && order != DCA_ORDER_FAST))
    return fail(h, DCA_ERR_ARG, "dca_acc_grid_step_train: bad argument");
  OpScratch* sp = nullptr;
  int rc = dev_scratch(device, &sp);
  if (rc) return rc;
  OpScratch& s = *sp;
  std::lock_guard<std::mutex> lk(s.mu);
  OpIO* d = s.d_io;
  memcpy(s.h_io->grid, grid, DCA_GRID_SIZE);
  memcpy(s.h_io->w, wnet, sizeof(float) * FREP);
  memcpy(s.h_io->wg, wgrad, sizeof(float) * FREP);
  memcpy(s.h_io->frep, frep, sizeof(int64_t) * FREP);
  memcpy(s.h_io->nfrep, next_frep, sizeof(int64_t) * FREP);
  if ((rc = io_h2d(s, IO_OFF(grid), IO_OFF(chs)))) return rc;
  // gridStep, Simulator.hs:137-144
  ops::k_grid_step<<<1, 256, 0, s.stream>>>(d->grid, r * COLS + c, is_end, ch, &d->reward);
  // backward, Agent.hs:44-81 with reward' = fromIntegral reward (SimRunner.hs:117)
  ops::k_backward_dev<<<1, ops::NT, 0, s.stream>>>(order, alpha_net, alpha_avg, alpha_grad, d->frep, d->nfrep, &d->reward,
                                                   *avg_reward, d->w, d->wg, &d->bout);
  if ((rc = io_d2h(s, IO_OFF(grid), IO_OFF(frep))) || (rc = io_d2h(s, IO_OFF(reward), IO_OFF(n))) || (rc = io_sync(s))) return rc;
  memcpy(grid, s.h_io->grid, DCA_GRID_SIZE);
  memcpy(wnet, s.h_io->w, sizeof(float) * FREP);
  memcpy(wgrad, s.h_io->wg, sizeof(float) * FREP);
  const ops::BackwardOut o = s.h_io->bout;
  *avg_reward = o.avg_reward;
  if (loss) *loss = o.loss;
  if (loss_is_nan) *loss_is_nan = o.loss_is_nan;
  if (reward) *reward = s.h_io->reward;
  return DCA_OK;
}

int dca_dense_dot(int32_t device, int32_t order, int32_t grad_corr_tree, const int64_t* frep, const float* w, float* out) {
  dca_handle* h = nullptr;
  if (!frep || !w || !out || (order != DCA_ORDER_REF && order != DCA_ORDER_FAST))
    return fail(h, DCA_ERR_ARG, "dca_dense_dot: bad argument");
  OpScratch* sp = nullptr;
  int rc = dev_scratch(device, &sp);
  if (rc) return rc;
  OpScratch& s = *sp;
  std::lock_guard<std::mutex> lk(s.mu);
  OpIO* d = s.d_io;
  memcpy(s.h_io->w, w, sizeof(float) * FREP);
  memcpy(s.h_io->frep, frep, sizeof(int64_t) * FREP);
  if ((rc = io_h2d(s, IO_OFF(w), IO_OFF(wg))) || (rc = io_h2d(s, IO_OFF(frep), IO_OFF(nfrep)))) return rc;
  if (grad_corr_tree) ops::k_dense_dot_g<<<1, ops::NT, 0, s.stream>>>(order, d->frep, d->w, &d->dot);
  else ops::k_dense_dots<<<1, ops::NT, 0, s.stream>>>(order, d->frep, 1, d->w, &d->dot);
  if ((rc = io_d2h(s, IO_OFF(dot), sizeof(OpIO))) || (rc = io_sync(s))) return rc;
  *out = s.h_io->dot;
  return DCA_OK;
}

// ---- NCCL, resolved at call time with dlopen (no link-time dependency) -----------------
// Communicators are created once per set of devices (ncclCommInitAll) and reused by later calls; a set whose
// collective failed is destroyed and forgotten.  Every NCCL return code is checked.
}  // extern "C"
namespace {
typedef void* nccl_comm_t;
struct NcclComms {
  std::vector<int> devs;
  std::vector<nccl_comm_t> comms;
};
struct NcclApi {
  std::mutex mu;
  void* lib = nullptr;
  int (*pInitAll)(nccl_comm_t*, int, const int*) = nullptr;
  int (*pAllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*pGroupStart)() = nullptr;
  int (*pGroupEnd)() = nullptr;
  int (*pDestroy)(nccl_comm_t) = nullptr;
  std::vector<NcclComms> cache;

  const char* load() {
    if (lib) return nullptr;
    void* l = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!l) l = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!l) return "dca_allreduce_stats: libnccl.so.2 not found";
    pInitAll = (int (*)(nccl_comm_t*, int, const int*))dlsym(l, "ncclCommInitAll");
    pAllReduce = (int (*)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(l, "ncclAllReduce");
    pGroupStart = (int (*)())dlsym(l, "ncclGroupStart");
    pGroupEnd = (int (*)())dlsym(l, "ncclGroupEnd");
    pDestroy = (int (*)(nccl_comm_t))dlsym(l, "ncclCommDestroy");
    if (!pInitAll || !pAllReduce || !pGroupStart || !pGroupEnd || !pDestroy) {
      dlclose(l);
      return "dca_allreduce_stats: NCCL symbols missing";
    }
    lib = l;
    return nullptr;
  }
  const char* comms_for(const std::vector<int>& devs, NcclComms** out) {
    for (NcclComms& c : cache)
      if (c.devs == devs) { *out = &c; return nullptr; }
    NcclComms c;
    c.devs = devs;
    c.comms.assign(devs.size(), nullptr);
    if (pInitAll(c.comms.data(), (int)devs.size(), devs.data()) != 0) {
      for (nccl_comm_t x : c.comms)
        if (x) pDestroy(x);
      return "dca_allreduce_stats: ncclCommInitAll failed";
    }
    cache.push_back(std::move(c));
    *out = &cache.back();
    return nullptr;
  }
  void drop(const std::vector<int>& devs) {
    for (size_t i = 0; i < cache.size(); ++i)
      if (cache[i].devs == devs) {
        for (nccl_comm_t x : cache[i].comms)
          if (x) pDestroy(x);
        cache.erase(cache.begin() + i);
        return;
      }
  }
};
NcclApi g_nccl;
}  // namespace
extern "C" {

// ---- multi-GPU stats reduction over NCCL (single process, many handles) --------
// NCCL is resolved at call time with dlopen so that the library has no link-time
// dependency on it (a process that already loaded NCCL, e.g. through
// torch.distributed, should all-reduce dca_sum_stats' device pointer itself).
int dca_allreduce_stats(dca_handle** handles, int32_t n, int64_t out[10]) {
  dca_handle* h = (handles && n > 0) ? handles[0] : nullptr;
  if (!h || !out) return fail(h, DCA_ERR_ARG, "dca_allreduce_stats: bad argument");
  for (int i = 0; i < n; ++i) {
    if (!handles[i]) return fail(h, DCA_ERR_ARG, "dca_allreduce_stats: null handle");
    int rc = dca_sum_stats(handles[i], nullptr, nullptr);
    if (rc) return rc;
  }
  if (n == 1) return dca_sum_stats(h, out, nullptr);
  std::vector<int> devs(n);
  for (int i = 0; i < n; ++i) devs[i] = handles[i]->device;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j)
      if (devs[i] == devs[j]) return fail(h, DCA_ERR_ARG, "dca_allreduce_stats: two handles on the same device (NCCL needs one rank per device)");
  std::lock_guard<std::mutex> lk(g_nccl.mu);
  const char* err = g_nccl.load();
  if (err) return fail(h, DCA_ERR_STATE, err);
  NcclComms* cs = nullptr;
  if ((err = g_nccl.comms_for(devs, &cs))) return fail(h, DCA_ERR_STATE, err);
  const int ncclInt64 = 4, ncclSum = 0;
  int bad = 0, rc_nccl = g_nccl.pGroupStart();
  if (rc_nccl == 0) {
    for (int i = 0; i < n && !bad; ++i) {
      if (cudaSetDevice(devs[i]) != cudaSuccess) { bad = -1; break; }
      bad = g_nccl.pAllReduce(handles[i]->d_sum, handles[i]->d_sum, 10, ncclInt64, ncclSum, cs->comms[i], handles[i]->stream);
    }
    rc_nccl = g_nccl.pGroupEnd();  // (always closes the group that was opened)
  }
  if (rc_nccl != 0 || bad != 0) {
    g_nccl.drop(devs);  // a failed collective leaves the communicators unusable: destroy them
    return fail(h, DCA_ERR_STATE, "dca_allreduce_stats: ncclAllReduce failed");
  }
  for (int i = 0; i < n; ++i) {
    if (cudaSetDevice(devs[i]) != cudaSuccess || cudaStreamSynchronize(handles[i]->stream) != cudaSuccess) {
      g_nccl.drop(devs);
      return fail(h, DCA_ERR_CUDA, "dca_allreduce_stats: stream synchronisation after the all-reduce failed");
    }
  }
  long long tmp[10];
  CK(cudaSetDevice(devs[0]));
  CK(cudaMemcpy(tmp, h->d_sum, sizeof tmp, cudaMemcpyDeviceToHost));
  for (int k = 0; k < 10; ++k) out[k] = tmp[k];
  return DCA_OK;
}

}  // extern "C"

#ifdef DCA_PHASE_CLOCK
extern "C" int dca_debug_cta(unsigned long long* out, int n_blocks) {
  return cudaMemcpyFromSymbol(out, dca::fast::g_cta, sizeof(unsigned long long) * 2 * (size_t)n_blocks) == cudaSuccess ? 0 : -1;
}
extern "C" int dca_debug_ref_phase(unsigned long long* out, int reset) {
  if (cudaMemcpyFromSymbol(out, dca::g_ref_phase, sizeof(unsigned long long) * 8) != cudaSuccess) return -1;
  if (reset) {
    unsigned long long z[8] = {0};
    cudaMemcpyToSymbol(dca::g_ref_phase, z, sizeof z);
  }
  return 0;
}
extern "C" int dca_debug_phase(unsigned long long* out, int reset) {
  if (cudaMemcpyFromSymbol(out, dca::fast::g_phase, sizeof(unsigned long long) * 48) != cudaSuccess) return -1;
  if (reset) {
    unsigned long long z[48] = {0};
    cudaMemcpyToSymbol(dca::fast::g_phase, z, sizeof z);
  }
  return 0;
}
#endif
