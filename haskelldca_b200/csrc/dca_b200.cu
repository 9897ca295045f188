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
  std::string err;
  float last_ms = 0.f;
  long long launches = 0;
  int warps = 4;
};

namespace {

int fail(dca_handle* h, int code, const char* msg) {
  g_last_error = msg;
  if (h) h->err = msg;
  return code;
}

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
  if (!h) return DCA_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->d_states);
  cudaFree(h->d_mean_new); cudaFree(h->d_call_dur); cudaFree(h->d_call_dur_hoff); cudaFree(h->d_hoff_prob);
  cudaFree(h->d_a_net); cudaFree(h->d_a_avg); cudaFree(h->d_a_grad);
  cudaFree(h->d_tape_w); cudaFree(h->d_tape_l); cudaFree(h->d_tape_off);
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
      (rc = set_smem(h, fast::k_step_fast, sizeof(fast::SmemF))) ||
      (rc = set_smem(h, k_run<ORDER_REF, false>)) || (rc = set_smem(h, k_run<ORDER_REF, true>)) ||
      (rc = set_smem(h, k_step<ORDER_REF>)) || (rc = set_smem(h, k_gf_violates)))
    return bail(rc);
  if ((rc = dca_reset(h))) return bail(rc);
  *out = h;
  return DCA_OK;
#undef CKB
}

int dca_reset(dca_handle* h) {
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

int dca_run(dca_handle* h, int64_t max_events) {
  if (!h || max_events < 0) return fail(h, DCA_ERR_ARG, "dca_run: bad argument");
  CK(cudaSetDevice(h->device));
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
  const bool trace = h->d_trace != nullptr;
  // (the TRACE instantiation of the fused FAST kernel also serves the verify flags: it has the extra per-event barrier)
  const bool trace_fast = trace || a.verify_rc || a.verify_nb;
  const dim3 grid(h->n), block(128);
  const size_t smem = sizeof(Smem);
  CK(cudaEventRecord(h->ev0, h->stream));
  if (h->cfg.order == DCA_ORDER_FAST) {
    if (trace_fast) fast::k_run_fast<true><<<grid, fast::NT, sizeof(fast::SmemF), h->stream>>>(a);
    else fast::k_run_fast<false><<<grid, fast::NT, sizeof(fast::SmemF), h->stream>>>(a);
  } else {
    if (trace) k_run<ORDER_REF, true><<<grid, block, smem, h->stream>>>(a);
    else k_run<ORDER_REF, false><<<grid, block, smem, h->stream>>>(a);
  }
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
  return DCA_OK;
}

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
  return DCA_OK;
}

int dca_get_action(dca_handle* h, int32_t replica, int32_t r, int32_t c, int32_t is_end, int32_t* ch) {
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
  if (bad_replica(h, replica)) return fail(h, DCA_ERR_ARG, "dca_write_state: bad replica");
  CK(cudaSetDevice(h->device));
  RepState* st = h->d_states + replica;
  if (grid) {
    uint8_t* d_grid = nullptr;
    CK(cudaMalloc(&d_grid, DCA_GRID_SIZE));
    CK(cudaMemcpyAsync(d_grid, grid, DCA_GRID_SIZE, cudaMemcpyHostToDevice, h->stream));
    k_gf_load_grid<<<1, 256, 0, h->stream>>>(st, d_grid);
    h->launches++;
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(d_grid);
  }
  if (wnet || wgrad) {
    float *dw = nullptr, *dg = nullptr;
    if (wnet) { CK(cudaMalloc(&dw, FREP * 4)); CK(cudaMemcpyAsync(dw, wnet, FREP * 4, cudaMemcpyHostToDevice, h->stream)); }
    if (wgrad) { CK(cudaMalloc(&dg, FREP * 4)); CK(cudaMemcpyAsync(dg, wgrad, FREP * 4, cudaMemcpyHostToDevice, h->stream)); }
    k_pack_weights<<<1, 256, 0, h->stream>>>(st, dw, dg);
    h->launches++;
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(dw);
    cudaFree(dg);
  }
  if (avg_reward)
    CK(cudaMemcpy(reinterpret_cast<unsigned char*>(st) + offsetof(RepState, avg_reward), avg_reward, 4, cudaMemcpyHostToDevice));
  CK(cudaGetLastError());
  return DCA_OK;
}

// ---- checkpoint / resume: the replica blobs verbatim behind a small header ----------
namespace {
struct CkptHeader {
  uint64_t magic;      // "DCAB200\1"
  uint32_t blob_bytes; // sizeof(RepState)
  int32_t n_replicas;
  int32_t order, rng_mode;
  uint64_t reserved;
};
constexpr uint64_t kCkptMagic = 0x0130303242414344ULL;
}  // namespace

size_t dca_checkpoint_size(const dca_handle* h) {
  return h ? sizeof(CkptHeader) + sizeof(RepState) * (size_t)h->n : 0;
}

int dca_checkpoint_save(dca_handle* h, void* buf, size_t size) {
  if (!h || !buf || size < dca_checkpoint_size(h)) return fail(h, DCA_ERR_ARG, "dca_checkpoint_save: bad argument");
  CK(cudaSetDevice(h->device));
  CkptHeader hd{kCkptMagic, (uint32_t)sizeof(RepState), h->n, h->cfg.order, h->cfg.rng_mode, 0};
  memcpy(buf, &hd, sizeof hd);
  CK(cudaMemcpyAsync(static_cast<char*>(buf) + sizeof hd, h->d_states, sizeof(RepState) * (size_t)h->n,
                     cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return DCA_OK;
}

int dca_checkpoint_load(dca_handle* h, const void* buf, size_t size) {
  if (!h || !buf || size < dca_checkpoint_size(h)) return fail(h, DCA_ERR_ARG, "dca_checkpoint_load: bad argument");
  CkptHeader hd;
  memcpy(&hd, buf, sizeof hd);
  if (hd.magic != kCkptMagic || hd.blob_bytes != sizeof(RepState) || hd.n_replicas != h->n ||
      hd.order != h->cfg.order || hd.rng_mode != h->cfg.rng_mode)
    return fail(h, DCA_ERR_STATE, "dca_checkpoint_load: the blob does not match this handle (replicas, order, rng mode or layout)");
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
  if (bad_replica(h, replica) || !stats) return fail(h, DCA_ERR_ARG, "dca_read_stats: bad argument");
  Summary s;
  int rc = fetch_one(h, replica, s);
  if (rc) return rc;
  for (int k = 0; k < 8; ++k) stats[k] = s.stats[k];
  return DCA_OK;
}

int dca_read_status(dca_handle* h, int32_t replica, int32_t* status, int64_t* iter, uint64_t* n_draws) {
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

int dca_sum_stats(dca_handle* h, int64_t out[10], void** dev_ptr) {
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

namespace {
struct GfScratch {  // a scratch replica blob + staging buffers on `device`
  RepState* st = nullptr;
  uint8_t* grid = nullptr;
  long long* frep = nullptr;
  long long* chs = nullptr;
  int* n = nullptr;
  ~GfScratch() { cudaFree(st); cudaFree(grid); cudaFree(frep); cudaFree(chs); cudaFree(n); }
};
int gf_begin(int device, const uint8_t* grid, GfScratch& g) {
  dca_handle* h = nullptr;
  if (!grid) return fail(h, DCA_ERR_ARG, "gridfuncs: null grid");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(h, DCA_ERR_NO_DEVICE, "gridfuncs: no CUDA device (there is no CPU fallback)");
  int rc = set_device(h, device);
  if (rc) return rc;
  CK(cudaMalloc(&g.st, sizeof(RepState)));
  CK(cudaMalloc(&g.grid, DCA_GRID_SIZE));
  CK(cudaMalloc(&g.frep, sizeof(long long) * FREP));
  CK(cudaMalloc(&g.chs, sizeof(long long) * NCH));
  CK(cudaMalloc(&g.n, sizeof(int)));
  CK(cudaMemcpy(g.grid, grid, DCA_GRID_SIZE, cudaMemcpyHostToDevice));
  k_gf_load_grid<<<1, 256>>>(g.st, g.grid);
  CK(cudaGetLastError());
  return DCA_OK;
}
int gf_chs(int device, const uint8_t* grid, int r, int c, int is_end, int64_t* chs, int32_t* n) {
  dca_handle* h = nullptr;
  if (r < 0 || r >= ROWS || c < 0 || c >= COLS || !chs || !n) return fail(h, DCA_ERR_ARG, "gridfuncs: bad argument");
  GfScratch g;
  int rc = gf_begin(device, grid, g);
  if (rc) return rc;
  k_gf_chs<<<1, 32>>>(g.st, r * COLS + c, is_end, g.chs, g.n);
  CK(cudaGetLastError());
  int cnt = 0;
  CK(cudaMemcpy(&cnt, g.n, sizeof(int), cudaMemcpyDeviceToHost));
  long long tmp[NCH];
  CK(cudaMemcpy(tmp, g.chs, sizeof(long long) * NCH, cudaMemcpyDeviceToHost));
  for (int i = 0; i < cnt; ++i) chs[i] = tmp[i];
  *n = cnt;
  return DCA_OK;
}
}  // namespace

int dca_gf_eligible_chs(int32_t device, const uint8_t* grid, int32_t r, int32_t c, int64_t* chs, int32_t* n) {
  return gf_chs(device, grid, r, c, 0, chs, n);
}
int dca_gf_inuse_chs(int32_t device, const uint8_t* grid, int32_t r, int32_t c, int64_t* chs, int32_t* n) {
  return gf_chs(device, grid, r, c, 1, chs, n);
}

int dca_gf_feature_rep(int32_t device, const uint8_t* grid, int64_t* frep) {
  dca_handle* h = nullptr;
  if (!frep) return fail(h, DCA_ERR_ARG, "dca_gf_feature_rep: null output");
  GfScratch g;
  int rc = gf_begin(device, grid, g);
  if (rc) return rc;
  k_gf_store_frep<<<1, 256>>>(g.st, g.frep);
  CK(cudaGetLastError());
  static_assert(sizeof(long long) == sizeof(int64_t), "int64");
  CK(cudaMemcpy(frep, g.frep, sizeof(long long) * FREP, cudaMemcpyDeviceToHost));
  return DCA_OK;
}

int dca_gf_inc_afterstate_freps(int32_t device, int32_t r, int32_t c, int32_t is_end, const int64_t* chs,
                                int32_t n, const uint8_t* grid, const int64_t* frep, int64_t* out) {
  dca_handle* h = nullptr;
  if (r < 0 || r >= ROWS || c < 0 || c >= COLS || n < 0 || n > NCH || !frep || !out || (n && !chs))
    return fail(h, DCA_ERR_ARG, "dca_gf_inc_afterstate_freps: bad argument");
  if (n == 0) return DCA_OK;
  for (int i = 0; i < n; ++i)
    if (chs[i] < 0 || chs[i] >= NCH) return fail(h, DCA_ERR_ARG, "dca_gf_inc_afterstate_freps: channel out of range");
  GfScratch g;
  int rc = gf_begin(device, grid, g);
  if (rc) return rc;
  CK(cudaMemcpy(g.frep, frep, sizeof(long long) * FREP, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(g.chs, chs, sizeof(long long) * n, cudaMemcpyHostToDevice));
  long long* d_out = nullptr;
  CK(cudaMalloc(&d_out, sizeof(long long) * (size_t)n * FREP));
  k_gf_inc_afterstate_freps<<<148, 256>>>(g.st, r * COLS + c, is_end, g.chs, n, g.frep, d_out);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpy(out, d_out, sizeof(long long) * (size_t)n * FREP, cudaMemcpyDeviceToHost);
  cudaFree(d_out);
  if (e != cudaSuccess) return fail(h, DCA_ERR_CUDA, cudaGetErrorString(e));
  return DCA_OK;
}

int dca_gf_violates_reuse_constraint(int32_t device, const uint8_t* grid, int32_t* violates) {
  dca_handle* h = nullptr;
  if (!violates) return fail(h, DCA_ERR_ARG, "dca_gf_violates_reuse_constraint: null output");
  GfScratch g;
  int rc = gf_begin(device, grid, g);
  if (rc) return rc;
  CK(cudaFuncSetAttribute(k_gf_violates, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  k_gf_violates<<<1, 128, sizeof(Smem)>>>(g.st, g.n);
  CK(cudaGetLastError());
  int v = 0;
  CK(cudaMemcpy(&v, g.n, sizeof(int), cudaMemcpyDeviceToHost));
  *violates = v;
  return DCA_OK;
}


}  // extern "C"

// ---- accFn1 / accFn2 on explicit arrays (SimRunner.hs:61-62) --------------------
namespace {
struct DevBuf {  // RAII device scratch
  void* p = nullptr;
  ~DevBuf() { cudaFree(p); }
  template <typename T> T* as() { return static_cast<T*>(p); }
};
int dev_alloc(DevBuf& b, size_t bytes) {
  dca_handle* h = nullptr;
  CK(cudaMalloc(&b.p, bytes ? bytes : 1));
  return DCA_OK;
}
}  // namespace

extern "C" {

int dca_acc_get_action(int32_t device, int32_t order, int32_t r, int32_t c, int32_t is_end, const float* wnet,
                       const uint8_t* grid, const int64_t* frep, int32_t* ch, int64_t* next_frep) {
  dca_handle* h = nullptr;
  if (r < 0 || r >= ROWS || c < 0 || c >= COLS || !wnet || !frep || !ch || !next_frep ||
      (order != DCA_ORDER_REF && order != DCA_ORDER_FAST))
    return fail(h, DCA_ERR_ARG, "dca_acc_get_action: bad argument");
  GfScratch g;
  int rc = gf_begin(device, grid, g);
  if (rc) return rc;
  const int cell = r * COLS + c;
  k_gf_chs<<<1, 32>>>(g.st, cell, is_end, g.chs, g.n);  // eligibleChs / inuseChs, Simulator.hs:161-165
  CK(cudaGetLastError());
  int n = 0;
  CK(cudaMemcpy(&n, g.n, sizeof(int), cudaMemcpyDeviceToHost));
  if (n == 0) {  // (Nothing, frep), Simulator.hs:170-171
    *ch = -1;
    memcpy(next_frep, frep, sizeof(int64_t) * FREP);
    return DCA_OK;
  }
  DevBuf afreps, w, q, idx;
  if ((rc = dev_alloc(afreps, sizeof(long long) * (size_t)n * FREP)) || (rc = dev_alloc(w, sizeof(float) * FREP)) ||
      (rc = dev_alloc(q, sizeof(float) * NCH)) || (rc = dev_alloc(idx, sizeof(int))))
    return rc;
  CK(cudaMemcpy(g.frep, frep, sizeof(long long) * FREP, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(w.p, wnet, sizeof(float) * FREP, cudaMemcpyHostToDevice));
  // selectAction, Simulator.hs:185-205: afterstate freps -> q = mvMul -> argpmax
  k_gf_inc_afterstate_freps<<<148, 256>>>(g.st, cell, is_end, g.chs, n, g.frep, afreps.as<long long>());
  ops::k_dense_dots<<<n, ops::NT>>>(order, afreps.as<long long>(), n, w.as<float>(), q.as<float>());
  ops::k_argpmax<<<1, 32>>>(q.as<float>(), n, idx.as<int>());
  CK(cudaGetLastError());
  int k = 0;
  CK(cudaMemcpy(&k, idx.p, sizeof(int), cudaMemcpyDeviceToHost));
  long long chs[NCH];
  CK(cudaMemcpy(chs, g.chs, sizeof(long long) * n, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(next_frep, afreps.as<long long>() + (size_t)k * FREP, sizeof(long long) * FREP, cudaMemcpyDeviceToHost));
  *ch = (int32_t)chs[k];
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
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(h, DCA_ERR_NO_DEVICE, "dca_acc_grid_step_train: no CUDA device (there is no CPU fallback)");
  int rc = set_device(h, device);
  if (rc) return rc;
  DevBuf dgrid, df, dnf, dw, dg, drew, dout;
  if ((rc = dev_alloc(dgrid, DCA_GRID_SIZE)) || (rc = dev_alloc(df, sizeof(long long) * FREP)) ||
      (rc = dev_alloc(dnf, sizeof(long long) * FREP)) || (rc = dev_alloc(dw, sizeof(float) * FREP)) ||
      (rc = dev_alloc(dg, sizeof(float) * FREP)) || (rc = dev_alloc(drew, sizeof(long long))) ||
      (rc = dev_alloc(dout, sizeof(ops::BackwardOut))))
    return rc;
  CK(cudaMemcpy(dgrid.p, grid, DCA_GRID_SIZE, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(df.p, frep, sizeof(long long) * FREP, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dnf.p, next_frep, sizeof(long long) * FREP, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw.p, wnet, sizeof(float) * FREP, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dg.p, wgrad, sizeof(float) * FREP, cudaMemcpyHostToDevice));
  // gridStep, Simulator.hs:137-144
  ops::k_grid_step<<<1, 256>>>(dgrid.as<uint8_t>(), r * COLS + c, is_end, ch, drew.as<long long>());
  CK(cudaGetLastError());
  long long rew = 0;
  CK(cudaMemcpy(&rew, drew.p, sizeof rew, cudaMemcpyDeviceToHost));
  // backward, Agent.hs:44-81 with reward' = fromIntegral reward (SimRunner.hs:117)
  ops::k_backward<<<1, ops::NT>>>(order, alpha_net, alpha_avg, alpha_grad, df.as<long long>(), dnf.as<long long>(),
                                  (float)rew, *avg_reward, dw.as<float>(), dg.as<float>(), dout.as<ops::BackwardOut>());
  CK(cudaGetLastError());
  ops::BackwardOut o;
  CK(cudaMemcpy(&o, dout.p, sizeof o, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(grid, dgrid.p, DCA_GRID_SIZE, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(wnet, dw.p, sizeof(float) * FREP, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(wgrad, dg.p, sizeof(float) * FREP, cudaMemcpyDeviceToHost));
  *avg_reward = o.avg_reward;
  if (loss) *loss = o.loss;
  if (loss_is_nan) *loss_is_nan = o.loss_is_nan;
  if (reward) *reward = rew;
  return DCA_OK;
}

int dca_dense_dot(int32_t device, int32_t order, int32_t grad_corr_tree, const int64_t* frep, const float* w, float* out) {
  dca_handle* h = nullptr;
  if (!frep || !w || !out || (order != DCA_ORDER_REF && order != DCA_ORDER_FAST))
    return fail(h, DCA_ERR_ARG, "dca_dense_dot: bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(h, DCA_ERR_NO_DEVICE, "dca_dense_dot: no CUDA device (there is no CPU fallback)");
  int rc = set_device(h, device);
  if (rc) return rc;
  DevBuf df, dw, dq;
  if ((rc = dev_alloc(df, sizeof(long long) * FREP)) || (rc = dev_alloc(dw, sizeof(float) * FREP)) ||
      (rc = dev_alloc(dq, sizeof(float))))
    return rc;
  CK(cudaMemcpy(df.p, frep, sizeof(long long) * FREP, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw.p, w, sizeof(float) * FREP, cudaMemcpyHostToDevice));
  if (grad_corr_tree) ops::k_dense_dot_g<<<1, ops::NT>>>(order, df.as<long long>(), dw.as<float>(), dq.as<float>());
  else ops::k_dense_dots<<<1, ops::NT>>>(order, df.as<long long>(), 1, dw.as<float>(), dq.as<float>());
  CK(cudaGetLastError());
  CK(cudaMemcpy(out, dq.p, sizeof(float), cudaMemcpyDeviceToHost));
  return DCA_OK;
}

// ---- multi-GPU stats reduction over NCCL (single process, many handles) --------
// NCCL is resolved at call time with dlopen so that the library has no link-time
// dependency on it (a process that already loaded NCCL, e.g. through
// torch.distributed, should all-reduce dca_sum_stats' device pointer itself).
int dca_allreduce_stats(dca_handle** handles, int32_t n, int64_t out[10]) {
  dca_handle* h = (handles && n > 0) ? handles[0] : nullptr;
  if (!h || !out) return fail(h, DCA_ERR_ARG, "dca_allreduce_stats: bad argument");
  for (int i = 0; i < n; ++i) {
    int rc = dca_sum_stats(handles[i], nullptr, nullptr);
    if (rc) return rc;
  }
  if (n == 1) return dca_sum_stats(h, out, nullptr);
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return fail(h, DCA_ERR_STATE, "dca_allreduce_stats: libnccl.so.2 not found");
  typedef void* comm_t;
  auto pInitAll = (int (*)(comm_t*, int, const int*))dlsym(lib, "ncclCommInitAll");
  auto pAllReduce = (int (*)(const void*, void*, size_t, int, int, comm_t, cudaStream_t))dlsym(lib, "ncclAllReduce");
  auto pGroupStart = (int (*)())dlsym(lib, "ncclGroupStart");
  auto pGroupEnd = (int (*)())dlsym(lib, "ncclGroupEnd");
  auto pDestroy = (int (*)(comm_t))dlsym(lib, "ncclCommDestroy");
  if (!pInitAll || !pAllReduce || !pGroupStart || !pGroupEnd || !pDestroy)
    return fail(h, DCA_ERR_STATE, "dca_allreduce_stats: NCCL symbols missing");
  std::vector<comm_t> comms(n);
  std::vector<int> devs(n);
  for (int i = 0; i < n; ++i) devs[i] = handles[i]->device;
  if (pInitAll(comms.data(), n, devs.data()) != 0) return fail(h, DCA_ERR_STATE, "ncclCommInitAll failed");
  const int ncclInt64 = 4, ncclSum = 0;
  pGroupStart();
  for (int i = 0; i < n; ++i) {
    cudaSetDevice(devs[i]);
    pAllReduce(handles[i]->d_sum, handles[i]->d_sum, 10, ncclInt64, ncclSum, comms[i], handles[i]->stream);
  }
  if (pGroupEnd() != 0) return fail(h, DCA_ERR_STATE, "ncclAllReduce failed");
  for (int i = 0; i < n; ++i) {
    cudaSetDevice(devs[i]);
    cudaStreamSynchronize(handles[i]->stream);
  }
  long long tmp[10];
  cudaSetDevice(devs[0]);
  CK(cudaMemcpy(tmp, h->d_sum, sizeof tmp, cudaMemcpyDeviceToHost));
  for (int k = 0; k < 10; ++k) out[k] = tmp[k];
  for (int i = 0; i < n; ++i) pDestroy(comms[i]);
  return DCA_OK;
}

}  // extern "C"

#ifdef DCA_PHASE_CLOCK
extern "C" int dca_debug_cta(unsigned long long* out, int n_blocks) {
  return cudaMemcpyFromSymbol(out, dca::fast::g_cta, sizeof(unsigned long long) * 2 * (size_t)n_blocks) == cudaSuccess ? 0 : -1;
}
extern "C" int dca_debug_phase(unsigned long long* out, int reset) {
  if (cudaMemcpyFromSymbol(out, dca::fast::g_phase, sizeof(unsigned long long) * 32) != cudaSuccess) return -1;
  if (reset) {
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbol(dca::fast::g_phase, z, sizeof z);
  }
  return 0;
}
#endif
