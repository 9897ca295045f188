// dca_host.hpp -- C++ host layer over the C ABI of libdca_b200 (include/dca_b200.h).
//
// The reference's host side is Haskell; no GHC exists in this build environment, so besides the
// Haskell bindings under hs/ (written, not compilable here) this header mirrors the reference's
// host modules for the hot path in C++, same names and meaning:
//   Opt / getOpts            src/Opt.hs:10-94      (every flag, default and help text)
//   Stats reports            src/Stats.hs:83-152   (statsCumus, statsReportPeriod, statsReportEnd)
//   SimStop                  src/SimRunner.hs:45-50
//   Replicas                 SimState + the operators on it (src/Simulator.hs:59-97)
//   runSim                   src/SimRunner.hs:206-243 (periods of logIter events, reports, speed line)
// Header-only; link with -ldca_b200.  Errors of the library surface as dcahost::Error.
#pragma once
#include <dca_b200.h>

#include <chrono>
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

namespace dcahost {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// ---- Opt (src/Opt.hs:10-25) -------------------------------------------------------------------
struct Opt {
  double callDurNew = 3.0;
  double callDurHoff = 1.0;
  double callRate = 200.0;
  double hoffProb = 0.0;
  long long nEvents = 10000;
  long long logIter = 1000;
  float alphaNet = 2.52e-6f;
  float alphaAvg = 0.06f;
  float alphaGrad = 5e-6f;
  bool verifyReuseConstraint = false;
  bool verifyNEligBounds = false;
  std::string backend = "gpu";  // the reference's `interp` / `cpu` Accelerate backends are not part of this build
  float minLoss = 0.0f;
  uint64_t rngSeed = 0;
  // additions of the batched build
  int nReplicas = 1;
  std::string order = "fast";   // fast | ref  (fp32 reduction order)
  std::string rng = "philox";   // philox | mt19937 (replay of the reference's stream)
  int device = 0;
  bool hoffLookahead = false;  // extension (the reference's TODO, README.md:71-72); --order ref only
};

inline const char* usage() {
  return "DCA - Dynamic Channel Allocation by Reinforcement Learning\n\n"
         "Usage: dca-exe [--call_dur MINUTES] [--call_dur_hoff MINUTES] [--call_rate PER_HOUR]\n"
         "               [--hoff_prob PROBABILITY] [--n_events N] [--log_iter N] [--learning_rate F]\n"
         "               [--learning_rate_avg F] [--learning_rate_grad F] [--verify_reuse_constraint]\n"
         "               [--verify_nelig_bounds] [--backend ARG] [--min_loss F] [--fixed_rng]\n"
         "               [--n_replicas N] [--order fast|ref] [--rng philox|mt19937] [--device N]\n"
         "               [--hoff_lookahead]\n\n"
         "Available options:\n"
         "  --call_dur MINUTES       Call duration for new calls. (default: 3.0)\n"
         "  --call_dur_hoff MINUTES  Call duration for handed-off calls. (default: 1.0)\n"
         "  --call_rate PER_HOUR     Call arrival rate (new calls). (default: 200.0)\n"
         "  --hoff_prob PROBABILITY  Hand-off probability. Set to 0 to disable hand-offs. (default: 0.0)\n"
         "  --n_events N             Simulation duration, in number of processed events. (default: 10000)\n"
         "  --log_iter N             How often to show run time statistics such as call blocking probability. (default: 1000)\n"
         "  --learning_rate F        For neural net, i.e. state value update. (default: 2.52e-6)\n"
         "  --learning_rate_avg F    Learning rate for the average reward estimate. (default: 6.0e-2)\n"
         "  --learning_rate_grad F   Learning rate for gradient correction. (default: 5.0e-6)\n"
         "  --verify_reuse_constraint\n"
         "  --verify_nelig_bounds\n"
         "  --backend ARG            'gpu' = libdca_b200 (sm_100a). (default: gpu)\n"
         "  --min_loss F             Abort simulation if loss goes below given absolute value. Set to 0 to disable. (default: 0.0)\n"
         "  --fixed_rng              Use a fixed (at 0) seed for the RNG. If this switch is not enabled, the seed is selected at random.\n"
         "  --n_replicas N           independent simulator replicas on the device (default: 1)\n"
         "  --order fast|ref         fp32 reduction order (default: fast)\n"
         "  --rng philox|mt19937     random source; mt19937 replays the reference's stream (default: philox)\n"
         "  --device N               CUDA device ordinal (default: 0)\n"
         "  --hoff_lookahead         extension: look ahead at the hand-off arrival when freeing a channel\n"
         "                           on the END half of a hand-off (--order ref only)\n"
         "  -h,--help                Show this help text\n";
}

// getOpts (src/Opt.hs:28-82): `rand` is the seed used unless --fixed_rng is given
inline Opt getOpts(int argc, const char* const* argv, uint64_t rand) {
  Opt o;
  bool fixed = false;
  auto need = [&](int& i) -> const char* {
    if (i + 1 >= argc) throw Error(DCA_ERR_ARG, std::string("The option `") + argv[i] + "` expects an argument.");
    return argv[++i];
  };
  auto num = [&](const char* flag, const char* s) -> double {
    char* end = nullptr;
    double v = std::strtod(s, &end);
    if (end == s || *end) throw Error(DCA_ERR_ARG, std::string("option ") + flag + ": cannot parse value `" + s + "'");
    return v;
  };
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--call_dur") o.callDurNew = num("--call_dur", need(i));
    else if (a == "--call_dur_hoff") o.callDurHoff = num("--call_dur_hoff", need(i));
    else if (a == "--call_rate") o.callRate = num("--call_rate", need(i));
    else if (a == "--hoff_prob") o.hoffProb = num("--hoff_prob", need(i));
    else if (a == "--n_events") o.nEvents = (long long)num("--n_events", need(i));
    else if (a == "--log_iter") o.logIter = (long long)num("--log_iter", need(i));
    else if (a == "--learning_rate") o.alphaNet = (float)num("--learning_rate", need(i));
    else if (a == "--learning_rate_avg") o.alphaAvg = (float)num("--learning_rate_avg", need(i));
    else if (a == "--learning_rate_grad") o.alphaGrad = (float)num("--learning_rate_grad", need(i));
    else if (a == "--verify_reuse_constraint") o.verifyReuseConstraint = true;
    else if (a == "--verify_nelig_bounds") o.verifyNEligBounds = true;
    else if (a == "--backend") {
      std::string b = need(i);
      for (auto& c : b) c = (char)std::tolower((unsigned char)c);
      if (b != "gpu")  // bkendOpt, src/Opt.hs:90-94
        throw Error(DCA_ERR_ARG, "option --backend: Accepted backend is 'gpu' (libdca_b200, sm_100a); the reference's "
                                 "'interp' and 'cpu' Accelerate backends are not part of this build.");
      o.backend = b;
    } else if (a == "--min_loss") o.minLoss = (float)num("--min_loss", need(i));
    else if (a == "--fixed_rng") fixed = true;
    else if (a == "--n_replicas") o.nReplicas = (int)num("--n_replicas", need(i));
    else if (a == "--order") o.order = need(i);
    else if (a == "--rng") o.rng = need(i);
    else if (a == "--device") o.device = (int)num("--device", need(i));
    else if (a == "--hoff_lookahead") o.hoffLookahead = true;
    else if (a == "-h" || a == "--help") throw Error(0, usage());
    else throw Error(DCA_ERR_ARG, "Invalid option `" + a + "'\n\n" + usage());
  }
  if (o.order != "fast" && o.order != "ref") throw Error(DCA_ERR_ARG, "option --order: fast | ref");
  if (o.rng != "philox" && o.rng != "mt19937") throw Error(DCA_ERR_ARG, "option --rng: philox | mt19937");
  o.rngSeed = fixed ? 0 : rand;  // seedParser, src/Opt.hs:79-82
  return o;
}

// ---- Stats (src/Stats.hs) -----------------------------------------------------------------------
struct Cumus { double cumuNew, cumuHoff, cumuTot; };
inline Cumus statsCumus(const int64_t st[DCA_N_STATS]) {  // :83-92
  double out[3];
  dca_stats_cumus(st, out);
  return {out[0], out[1], out[2]};
}
inline std::string fmt(const char* f, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, f);
  vsnprintf(buf, sizeof buf, f, ap);
  va_end(ap);
  return buf;
}
// statsReportPeriod (:94-124); the caller resets the period counters
inline std::string statsReportPeriod(const int64_t st[DCA_N_STATS], long long i, long long logIter, float lossSum,
                                     long long lossN, float avgReward) {
  const Cumus c = statsCumus(st);
  const double bp = (double)st[DCA_ST_CURR_REJECTED_NEW] / (double)(st[DCA_ST_CURR_ARRIVALS_NEW] + 1);
  const float avgLoss = lossSum / (float)lossN;
  return fmt("Blocking probability events %lld-%lld: %.4f, cumulative %.4f, avg. loss: %.1f, avg. reward %.4f",
             std::max(i - logIter, 0LL), i, bp, c.cumuNew, (double)avgLoss, (double)avgReward);
}
// statsReportEnd (:126-152)
inline std::string statsReportEnd(const int64_t st[DCA_N_STATS], long long nCallsInProgress, long long i) {
  const Cumus c = statsCumus(st);
  std::string s = fmt("Finished %lld events. Blocking probability %.4f for new calls", i, c.cumuNew);
  s += st[DCA_ST_REJECTED_HOFF] > 0 ? fmt(", %.4f for hand-offs, %.4f total.", c.cumuHoff, c.cumuTot) : std::string(".");
  const long long reported = st[DCA_ST_ARRIVALS_NEW] + st[DCA_ST_ARRIVALS_HOFF] - st[DCA_ST_REJECTED_NEW] -
                             st[DCA_ST_REJECTED_HOFF] - st[DCA_ST_ENDED];
  s += fmt("\n%lld new arrivals of which %lld have been rejected. %lld calls in progress at simulation end.",
           (long long)st[DCA_ST_ARRIVALS_NEW], (long long)st[DCA_ST_REJECTED_NEW], nCallsInProgress);
  if (reported != nCallsInProgress)
    s += fmt("\n\n\nSOME CALLS WERE LOST. This should never happen on runs that exits with 'Success'. According to"
             " reported calls there should be %lld calls currently in progress at simulation end but there are %lld.",
             reported, nCallsInProgress);
  return s;
}

// ---- SimStop (src/SimRunner.hs:45-50) --------------------------------------------------------------
inline const char* simStop(int status) {
  switch (status) {
    case DCA_RUNNING: return "Paused";
    case DCA_SUCCESS: return "Success";
    case DCA_ZERO_LOSS: return "ZeroLoss";
    case DCA_NAN_LOSS: return "NaNLoss";
    case DCA_REUSE_VIOLATED: return "InternalError ReuseConstraintViolated";
    case DCA_NELIG_OVERFLOW: return "InternalError NEligOverflow";
    case DCA_TAPE_EXHAUSTED: return "InternalError TapeExhausted";
    case DCA_UNINITIALIZED: return "InternalError Uninitialized";
    case DCA_QUEUE_OVERFLOW: return "InternalError EventQueueOverflow";
  }
  return "?";
}

// ---- Replicas: mkSimState for n replicas + the operators (src/Simulator.hs:59-97) --------------------
class Replicas {
 public:
  explicit Replicas(const Opt& o) : opt_(o) {
    dca_config cfg;
    check(dca_config_defaults(&cfg), "dca_config_defaults");
    cfg.n_replicas = o.nReplicas;
    cfg.device = o.device;
    cfg.hoff_lookahead = o.hoffLookahead ? 1 : 0;
    cfg.seed = o.rngSeed;
    cfg.rng_mode = o.rng == "philox" ? DCA_RNG_PHILOX : DCA_RNG_TAPE;
    cfg.order = o.order == "fast" ? DCA_ORDER_FAST : DCA_ORDER_REF;
    cfg.n_events = o.nEvents;
    cfg.min_loss = o.minLoss;
    cfg.verify_reuse_constraint = o.verifyReuseConstraint;
    cfg.verify_nelig_bounds = o.verifyNEligBounds;
    cfg.call_dur = o.callDurNew;
    cfg.call_dur_hoff = o.callDurHoff;
    cfg.call_rate = o.callRate;
    cfg.hoff_prob = o.hoffProb;
    cfg.alpha_net = o.alphaNet;
    cfg.alpha_avg = o.alphaAvg;
    cfg.alpha_grad = o.alphaGrad;
    check(dca_create(&cfg, &h_), "dca_create");
    if (o.rng == "mt19937")  // --fixed_rng replay: the reference's MT19937-64 stream (EventGen/Internal.hs:57-66)
      for (int r = 0; r < o.nReplicas; ++r)
        check(dca_set_rng_mt19937(h_, r, o.rngSeed + (uint64_t)r, (size_t)(4 * o.nEvents + 64)), "dca_set_rng_mt19937");
  }
  ~Replicas() { dca_destroy(h_); }
  Replicas(const Replicas&) = delete;
  Replicas& operator=(const Replicas&) = delete;

  float run(long long maxEvents) {  // runPeriod's loop of runStep on the device; returns device ms
    check(dca_run(h_, maxEvents), "dca_run");
    float ms = 0;
    dca_elapsed_ms(h_, &ms);
    return ms;
  }
  void runAsync(long long maxEvents) { check(dca_run_async(h_, maxEvents), "dca_run_async"); }  // enqueue and return
  float wait() {  // collect the launch of runAsync; returns device ms
    check(dca_wait(h_), "dca_wait");
    float ms = 0;
    dca_elapsed_ms(h_, &ms);
    return ms;
  }
  dca_handle* handle() const { return h_; }
  int getAction(int replica, int r, int c, bool isEnd) {  // accFn1
    int32_t ch = -1;
    check(dca_get_action(h_, replica, r, c, isEnd, &ch), "dca_get_action");
    return ch;
  }
  struct Status { int status; long long iter; unsigned long long nDraws; };
  Status status(int replica) {
    int32_t s; int64_t it; uint64_t nd;
    check(dca_read_status(h_, replica, &s, &it, &nd), "dca_read_status");
    return {s, it, nd};
  }
  void stats(int replica, int64_t out[DCA_N_STATS]) { check(dca_read_stats(h_, replica, out), "dca_read_stats"); }
  struct Period { float lossSum; long long n; };
  Period period(int replica, bool reset) {
    float s; int64_t n;
    check(dca_read_period(h_, replica, &s, &n, reset), "dca_read_period");
    return {s, n};
  }
  struct State {
    std::vector<uint8_t> grid;
    std::vector<int64_t> frep;
    std::vector<float> wNet, wGradCorr;
    float avgReward;
    long long iter;
    dca_event event;
  };
  State state(int replica) {
    State s;
    s.grid.resize(DCA_GRID_SIZE); s.frep.resize(DCA_FREP_SIZE); s.wNet.resize(DCA_FREP_SIZE); s.wGradCorr.resize(DCA_FREP_SIZE);
    int64_t it;
    check(dca_read_state(h_, replica, s.grid.data(), s.frep.data(), s.wNet.data(), s.wGradCorr.data(), &s.avgReward, &it, &s.event),
          "dca_read_state");
    s.iter = it;
    return s;
  }
  void sumStats(int64_t out[10]) { check(dca_sum_stats(h_, out, nullptr), "dca_sum_stats"); }
  dca_handle* handle() { return h_; }
  const Opt& opt() const { return opt_; }

 private:
  void check(int rc, const char* what) {
    if (rc != DCA_OK) throw Error(rc, std::string(what) + ": " + dca_last_error(h_));
  }
  Opt opt_;
  dca_handle* h_ = nullptr;
};

// ---- runSim (src/SimRunner.hs:206-243) ---------------------------------------------------------------
// Periods of logIter events (one dca_run each), the reference's period / end reports for replica 0, and
// its closing speed line.  Returns the SimStop status of replica 0.
inline int runSim(const Opt& opt, const std::function<void(const std::string&)>& out) {
  const auto t0 = std::chrono::steady_clock::now();
  Replicas sims(opt);
  int status = DCA_RUNNING;
  long long lastIter = -1;
  while (status == DCA_RUNNING) {
    sims.run(opt.logIter);  // one period (SimRunner.hs:167-168)
    const Replicas::Status st = sims.status(0);
    status = st.status;
    if (st.iter == lastIter)  // a period without progress can only be a library fault: never spin on it
      throw std::runtime_error(fmt("replica 0 made no progress in a period (status %d, iter %lld)", status, st.iter));
    lastIter = st.iter;
    int64_t counters[DCA_N_STATS];
    sims.stats(0, counters);
    const Replicas::Period p = sims.period(0, true);
    const float avgR = sims.state(0).avgReward;
    out(statsReportPeriod(counters, st.iter, opt.logIter, p.lossSum, p.n, avgR));
    if (status != DCA_RUNNING) out(simStop(status));
  }
  const Replicas::State s = sims.state(0);
  int64_t counters[DCA_N_STATS];
  sims.stats(0, counters);
  long long inProgress = 0;
  for (uint8_t b : s.grid) inProgress += b;  // boolSum of the final grid (SimRunner.hs:224)
  out(statsReportEnd(counters, inProgress, s.iter));
  int64_t tot[10];
  sims.sumStats(tot);
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  const double tSim = s.event.time;
  out(fmt("\nSimulation duration: %dm%ds in sim time, %dm%ds wall clock with speed %lld events at %.0f events/second",
          (int)tSim, (int)((tSim - (int)tSim) * 60.0), (int)(dt / 60.0), (int)dt - 60 * (int)(dt / 60.0),
          (long long)tot[8], (double)tot[8] / dt));
  return status;
}

}  // namespace dcahost
