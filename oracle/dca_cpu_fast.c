/*
 * dca_cpu_fast.c -- the OPTIMISED CPU competitor ("port-optimised") for bench.py's cpu_baseline.
 *
 * TEST / BENCH INFRASTRUCTURE, like the rest of oracle/: only tests/ and bench.py's cpu_baseline leg
 * load it; nothing under haskelldca_b200/ links, imports or executes it.
 *
 * What it is: the same per-event hot path as dca_oracle.c (runStep, src/SimRunner.hs:60-97) in the
 * FAST arithmetic order (dca_oracle.h), written the way one would write it for a CPU that has to
 * compete -- instead of the reference's literal formulation (materialise n x 3479 afterstates,
 * src/Gridfuncs.hs:186-252, and run n dense dots, src/Simulator.hs:190-198):
 *   - grid as 70-bit masks, cnt2[cell][ch] = users of ch within distance 2 (eligibility = cnt2 == 0,
 *     src/Gridfuncs.hs:69-76; the nested scan of :220-232 becomes cnt2 == 0 / == 1);
 *   - features as fp32 in a [plane][column][row] layout, so that the 7 rows of a plane sit in one AVX2
 *     vector and the FAST order's row dots (even / odd column fma chains) vectorise over rows;
 *   - q-values in DELTA FORM: the plane sums of x.w are kept from the previous update, a candidate
 *     re-evaluates only the two planes its action touches (channel plane, n_elig plane) and substitutes
 *     them into the V tree -- bit for bit the dense dot of the materialised afterstate in the FAST tree;
 *   - one fused dense pass per event: w' = w - g, wg' = fma(sg, x, wg) (src/Agent.hs:67-76) together with
 *     the row dots of x'.w' and x'.wg' for the next event;
 *   - OpenMP over replicas.
 * The event generator, RNG and statistics are the oracle's own (orc_eg_*, orc_rng_*).
 *
 * Parity: tests/test_oracle_cpu_fast.py runs it against ORC_ORDER_FAST of the literal oracle -- every
 * trace record and the final state bit for bit.  It is therefore "the same program, made fast", and
 * its events/s is the fair CPU number next to the GPU's.
 */
#include <immintrin.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "dca_oracle.h"

#define NR 7
#define NC 7
#define NCHAN 70
#define NPLANE 71
#define NCELLS 49
#define PV (NC * 8)            /* floats per plane: 7 column vectors of 8 row lanes */
#define VIDX(f, c) (((f) * NC + (c)) * 8)

typedef struct {
  float n4f[NCELLS][PV];       /* N4(cell) \ {cell} as 0/1 floats in [column][row] layout */
  uint8_t n2list[NCELLS][20];  /* N2(cell) u {cell}: cell ids */
  uint8_t n2cnt[NCELLS];
  uint64_t n4mask[NCELLS];
  int ready;
} cpf_tables;
static cpf_tables T;

static void cpf_build_tables(void) {
  if (T.ready) return;
#ifdef _OPENMP
#pragma omp critical(cpf_tables)
#endif
  {
    if (!T.ready) {
      int buf[2 * 64];
      memset(&T, 0, sizeof T);
      for (int r = 0; r < NR; ++r)
        for (int c = 0; c < NC; ++c) {
          const int cell = r * NC + c;
          int n4 = orc_get_neighs(4, r, c, 0, buf); /* Gridneighs.hs:92-98 */
          for (int i = 0; i < n4; ++i) {
            const int rr = buf[2 * i], cc = buf[2 * i + 1];
            T.n4f[cell][cc * 8 + rr] = 1.0f;
            T.n4mask[cell] |= 1ULL << (rr * NC + cc);
          }
          int n2 = orc_get_neighs(2, r, c, 1, buf);
          T.n2cnt[cell] = (uint8_t)n2;
          for (int i = 0; i < n2; ++i) T.n2list[cell][i] = (uint8_t)(buf[2 * i] * NC + buf[2 * i + 1]);
        }
      T.ready = 1;
    }
  }
}

struct cpf_sim {
  /* [plane][column][row lane]; row lane 7 and nothing else is padding (x = 0 there) */
  float xf[NPLANE * PV] __attribute__((aligned(32)));
  float w[NPLANE * PV] __attribute__((aligned(32)));
  float wg[NPLANE * PV] __attribute__((aligned(32)));
  float P[72];                 /* plane sums of x.w (current state, current weights); P[71] = 0 */
  float dotg;                  /* x.wGradCorr */
  uint8_t cnt2[NCELLS][NCHAN + 2];
  uint64_t gm[NCELLS][2];      /* in-use channel masks */
  int ncalls;
  float avg_reward;
  orc_opt opt;
  orc_event event;
  orc_eventgen *eg;
  int64_t stats[8];
  int64_t iter;
  int status;
  float loss_sum;
  int64_t loss_n;
};
typedef struct cpf_sim cpf_sim;

static inline __m256 lane_mask(void) {
  return _mm256_castsi256_ps(_mm256_setr_epi32(-1, -1, -1, -1, -1, -1, -1, 0));
}
/* rowdot of the FAST order for the 7 rows of a plane at once (lane 7 masked to 0):
 * e = x0*w0; e = fma(x2,w2,e); e = fma(x4,w4,e); e = fma(x6,w6,e); o = x1*w1; o = fma(x3,w3,o);
 * o = fma(x5,w5,o); e + o */
static inline __m256 rowdot8(const __m256 x[7], const __m256 w[7]) {
  __m256 e = _mm256_mul_ps(x[0], w[0]);
  __m256 o = _mm256_mul_ps(x[1], w[1]);
  e = _mm256_fmadd_ps(x[2], w[2], e);
  o = _mm256_fmadd_ps(x[3], w[3], o);
  e = _mm256_fmadd_ps(x[4], w[4], e);
  o = _mm256_fmadd_ps(x[5], w[5], o);
  e = _mm256_fmadd_ps(x[6], w[6], e);
  return _mm256_and_ps(_mm256_add_ps(e, o), lane_mask());
}
/* ((a0+a1)+(a2+a3)) + ((a4+a5)+(a6+a7)): the xor-butterfly 1,2,4 over the rows of a plane */
static inline float tree8v(__m256 v) {
  __m256 t = _mm256_add_ps(v, _mm256_permute_ps(v, 0xB1));        /* swap neighbours */
  t = _mm256_add_ps(t, _mm256_permute_ps(t, 0x4E));               /* swap pairs */
  t = _mm256_add_ps(t, _mm256_permute2f128_ps(t, t, 0x01));       /* swap halves */
  return _mm256_cvtss_f32(t);
}

/* V tree of the FAST order: v[l] = (P[l] + P[l+32]) + (l < 6 ? P[l+64] : 0); balanced tree over the
 * 32 leaves (what the xor-butterfly 1..16 computes in every lane); + P[70].  The inner nodes are kept
 * so that one leaf can be substituted in five additions. */
typedef struct {
  float L[5][32];  /* L[0] = leaves, L[k] = sums of 2^k leaves */
  float val;
} vtree;
static void vtree_build(const float *P, vtree *t) {
  for (int l = 0; l < 32; ++l) t->L[0][l] = (P[l] + P[l + 32]) + (l < 6 ? P[l + 64] : 0.0f);
  for (int k = 1; k < 5; ++k)
    for (int i = 0; i < (32 >> k); ++i) t->L[k][i] = t->L[k - 1][2 * i] + t->L[k - 1][2 * i + 1];
  t->val = (t->L[4][0] + t->L[4][1]) + P[NCHAN];
}
static float vtree_subst(const float *P, const vtree *t, int ch, float Pp, float Ep) {
  const int lc = ch & 31, hi = ch >> 5;
  const float p0 = hi == 0 ? Pp : P[lc], p1 = hi == 1 ? Pp : P[lc + 32];
  const float p2 = hi == 2 ? Pp : (lc < 6 ? P[lc + 64] : 0.0f);
  float b = (p0 + p1) + p2;
  int i = lc;
  for (int k = 0; k < 5; ++k) {
    b = b + t->L[k][i ^ 1];
    i >>= 1;
  }
  return b + Ep;
}

/* the sums of the current state: P[f], dotg (no update) */
static void cpf_sums(cpf_sim *s) {
  __m256 acc[12];
  for (int f = 0; f < NPLANE; ++f) {
    __m256 x[7], w[7], g[7];
    for (int c = 0; c < NC; ++c) {
      x[c] = _mm256_load_ps(&s->xf[VIDX(f, c)]);
      w[c] = _mm256_load_ps(&s->w[VIDX(f, c)]);
      g[c] = _mm256_load_ps(&s->wg[VIDX(f, c)]);
    }
    s->P[f] = tree8v(rowdot8(x, w));
    const __m256 rg = rowdot8(x, g);
    acc[f % 12] = f < 12 ? rg : _mm256_add_ps(acc[f % 12], rg);
  }
  s->P[71] = 0.0f;
  float Tg[12], W[3];
  for (int g = 0; g < 12; ++g) Tg[g] = tree8v(acc[g]);
  for (int wp = 0; wp < 3; ++wp) W[wp] = (Tg[4 * wp] + Tg[4 * wp + 1]) + (Tg[4 * wp + 2] + Tg[4 * wp + 3]);
  s->dotg = (W[0] + W[1]) + W[2];
}

cpf_sim *cpf_create(const orc_opt *opt) { /* mkSimState, Simulator.hs:88-97 */
  cpf_build_tables();
  cpf_sim *s = (cpf_sim *)aligned_alloc(64, (sizeof(cpf_sim) + 63) / 64 * 64);
  memset(s, 0, sizeof *s);
  s->opt = *opt;
  s->eg = orc_eg_create(opt, 1);
  orc_eg_pop(s->eg, &s->event);
  for (int c = 0; c < NC; ++c)
    for (int r = 0; r < NR; ++r) s->xf[VIDX(NCHAN, c) + r] = (float)NCHAN; /* featureRep emptyGrid */
  cpf_sums(s);
  s->status = ORC_RUNNING;
  return s;
}
void cpf_destroy(cpf_sim *s) {
  if (!s) return;
  orc_eg_destroy(s->eg);
  free(s);
}

/* environmentStep, Simulator.hs:101-134 (same statements as the oracle's) */
static int cpf_environment_step(cpf_sim *s, int act) {
  const orc_event ev = s->event;
  const double time = ev.time;
  int64_t *st = s->stats;
  orc_rng *rng = orc_eg_rng(s->eg);
  switch (ev.etype) {
    case ORC_NEW:
      st[6]++; st[0]++;
      if (act >= 0) st[1]++;
      else { st[2]++; st[7]++; }
      orc_eg_generate_new(s->eg, time, ev.r, ev.c);
      if (act >= 0) {
        const double p = orc_rng_double(rng);
        if (p < s->opt.hoff_prob) orc_eg_generate_hoff_new(s->eg, time, ev.r, ev.c, act);
        else orc_eg_generate_end(s->eg, time, ev.r, ev.c, act);
      }
      break;
    case ORC_HOFF:
      st[3]++;
      if (act < 0) { st[2]++; st[7]++; } /* Simulator.hs:113 */
      else orc_eg_generate_hoff_end(s->eg, time, ev.r, ev.c, act);
      break;
    default:
      st[4]++;
      if (act < 0) return -1;
      if (ev.ch != act && orc_eg_reassign(s->eg, ev.r, ev.c, act, ev.ch) != 0) return -1;
  }
  s->iter += 1;
  return orc_eg_pop(s->eg, &s->event);
}

static void cpf_unpack(const cpf_sim *s, uint8_t *grid, int64_t *frep, float *w, float *wg);

static int cpf_step(cpf_sim *s, orc_trace *tr) {
  const orc_opt *o = &s->opt;
  const orc_event ev = s->event;
  const int is_end = ev.etype == ORC_END;
  const int cell = ev.r * NC + ev.c;
  /* ---- accFn1 = getAction, Simulator.hs:154-172 ---- */
  uint8_t cand[NCHAN];
  int n = 0;
  if (is_end) { /* inuseChs, ascending */
    for (int h = 0; h < 2; ++h) {
      uint64_t m = s->gm[cell][h];
      while (m) {
        cand[n++] = (uint8_t)(64 * h + __builtin_ctzll(m));
        m &= m - 1;
      }
    }
  } else { /* eligibleChs <=> cnt2 == 0 */
    for (int ch = 0; ch < NCHAN; ++ch)
      if (s->cnt2[cell][ch] == 0) cand[n++] = (uint8_t)ch;
  }
  vtree vt;
  vtree_build(s->P, &vt);
  const float val = vt.val;
  const __m256 sgn = _mm256_set1_ps(is_end ? -1.0f : 1.0f);
  const uint8_t want = is_end ? 1 : 0;
  const float ediff = is_end ? 1.0f : -1.0f;
  float q[NCHAN];
  __m256 w70[7], e0[7];
  for (int c = 0; c < NC; ++c) {
    w70[c] = _mm256_load_ps(&s->w[VIDX(NCHAN, c)]);
    e0[c] = _mm256_load_ps(&s->xf[VIDX(NCHAN, c)]);
  }
  for (int k = 0; k < n; ++k) { /* selectAction, Simulator.hs:185-205, in delta form */
    const int ch = cand[k];
    __m256 x[7], w[7];
    for (int c = 0; c < NC; ++c) { /* channel plane: x' = x +- [N4 \ {cell}], Gridfuncs.hs:212-217 */
      x[c] = _mm256_fmadd_ps(sgn, _mm256_load_ps(&T.n4f[cell][c * 8]), _mm256_load_ps(&s->xf[VIDX(ch, c)]));
      w[c] = _mm256_load_ps(&s->w[VIDX(ch, c)]);
    }
    const float Pp = tree8v(rowdot8(x, w));
    float e[PV] __attribute__((aligned(32))); /* n_elig plane: e' = e -+ [ch eligible there on grid'], :220-252 */
    for (int c = 0; c < NC; ++c) _mm256_store_ps(&e[c * 8], e0[c]);
    for (int i = 0; i < T.n2cnt[cell]; ++i) {
      const int c2 = T.n2list[cell][i];
      if (s->cnt2[c2][ch] == want) e[(c2 % NC) * 8 + c2 / NC] += ediff;
    }
    for (int c = 0; c < NC; ++c) x[c] = _mm256_load_ps(&e[c * 8]);
    const float Ep = tree8v(rowdot8(x, w70));
    q[k] = vtree_subst(s->P, &vt, ch, Pp, Ep);
  }
  int act = -1;
  float next_val = val;
  if (n > 0) {
    const int idx = orc_argpmax(q, n); /* AccUtils.hs:203-208 */
    act = cand[idx];
    next_val = q[idx];
  }
  if (cpf_environment_step(s, act) != 0) return ORC_HOST_ERROR;
  /* ---- accFn2 = gridStepTrain, SimRunner.hs:100-118 ---- */
  float enew[PV] __attribute__((aligned(32)));
  if (act >= 0) {
    for (int c = 0; c < NC; ++c) _mm256_store_ps(&enew[c * 8], e0[c]);
    for (int i = 0; i < T.n2cnt[cell]; ++i) {
      const int c2 = T.n2list[cell][i];
      if (s->cnt2[c2][act] == want) enew[(c2 % NC) * 8 + c2 / NC] += ediff;
      s->cnt2[c2][act] = (uint8_t)(s->cnt2[c2][act] + (is_end ? -1 : 1));
    }
    if (is_end) s->gm[cell][act >> 6] &= ~(1ULL << (act & 63)); /* executeAction, Gridfuncs.hs:105-110 */
    else s->gm[cell][act >> 6] |= 1ULL << (act & 63);
    s->ncalls += is_end ? -1 : 1;
  }
  const float reward = (float)s->ncalls; /* boolSum nextGrid */
  const float avgR = s->avg_reward;
  const float td = ((reward - avgR) + next_val) - val; /* Agent.hs:65 */
  const float cc = -2.0f * o->alpha_net;
  const float ctd = cc * td, cavg = cc * avgR, ncdot = -(cc * s->dotg);
  const float sg = o->alpha_grad * (td - s->dotg);
  s->avg_reward = avgR + o->alpha_avg * td;
  {
    const __m256 vctd = _mm256_set1_ps(ctd), vcavg = _mm256_set1_ps(cavg), vncdot = _mm256_set1_ps(ncdot),
                 vsg = _mm256_set1_ps(sg);
    __m256 acc[12];
    for (int f = 0; f < NPLANE; ++f) {
      __m256 xn[7], wv[7], gv[7];
      const int touched = act >= 0 && (f == act || f == NCHAN);
      for (int c = 0; c < NC; ++c) {
        const __m256 xo = _mm256_load_ps(&s->xf[VIDX(f, c)]);
        __m256 x1 = xo;
        if (touched) {
          x1 = f == act ? _mm256_fmadd_ps(sgn, _mm256_load_ps(&T.n4f[cell][c * 8]), xo) : _mm256_load_ps(&enew[c * 8]);
          _mm256_store_ps(&s->xf[VIDX(f, c)], x1);
        }
        /* g = fma(-cdot, x', fma(ctd, x, cavg)); w' = w - g; wg' = fma(sg, x, wg)   (Agent.hs:67-76) */
        const __m256 t = _mm256_fmadd_ps(vctd, xo, vcavg);
        const __m256 g = _mm256_fmadd_ps(vncdot, x1, t);
        wv[c] = _mm256_sub_ps(_mm256_load_ps(&s->w[VIDX(f, c)]), g);
        gv[c] = _mm256_fmadd_ps(vsg, xo, _mm256_load_ps(&s->wg[VIDX(f, c)]));
        _mm256_store_ps(&s->w[VIDX(f, c)], wv[c]);
        _mm256_store_ps(&s->wg[VIDX(f, c)], gv[c]);
        xn[c] = x1;
      }
      s->P[f] = tree8v(rowdot8(xn, wv));
      const __m256 rg = rowdot8(xn, gv);
      acc[f % 12] = f < 12 ? rg : _mm256_add_ps(acc[f % 12], rg);
    }
    float Tg[12], W[3];
    for (int g = 0; g < 12; ++g) Tg[g] = tree8v(acc[g]);
    for (int wp = 0; wp < 3; ++wp) W[wp] = (Tg[4 * wp] + Tg[4 * wp + 1]) + (Tg[4 * wp + 2] + Tg[4 * wp + 3]);
    s->dotg = (W[0] + W[1]) + W[2];
  }
  const int is_nan = isnan(td) ? 1 : 0;
  if (tr) {
    uint8_t grid[ORC_GRID_SIZE];
    int64_t frep[ORC_FREP_SIZE];
    cpf_unpack(s, grid, frep, NULL, NULL);
    memset(tr, 0, sizeof *tr);
    tr->time = ev.time;
    tr->loss = td;
    tr->avg_reward = s->avg_reward;
    tr->reward = s->ncalls;
    tr->etype = (uint8_t)ev.etype;
    tr->cell = (uint8_t)cell;
    tr->ch = (int8_t)act;
    tr->ncand = (uint8_t)n;
    tr->end_ch = (int8_t)(is_end ? ev.ch : -1);
    tr->grid_hash = orc_grid_hash(grid);
    tr->frep_hash = orc_frep_hash(frep);
  }
  if (is_nan) return ORC_NAN_LOSS; /* SimRunner.hs:152 */
  s->loss_sum = s->loss_sum + td;
  s->loss_n++;
  if (o->min_loss > 0 && s->iter > 50 && fabsf(td) < o->min_loss) return ORC_ZERO_LOSS;
  if (s->iter == o->n_events) return ORC_SUCCESS;
  return ORC_RUNNING;
}

int cpf_run(cpf_sim *s, int64_t max_steps, orc_trace *trace, int64_t trace_cap, int64_t *steps_done) {
  int64_t n = 0;
  while (n < max_steps && s->status == ORC_RUNNING) {
    s->status = cpf_step(s, (trace && n < trace_cap) ? &trace[n] : NULL);
    ++n;
  }
  if (steps_done) *steps_done = n;
  return s->status;
}

static void cpf_unpack(const cpf_sim *s, uint8_t *grid, int64_t *frep, float *w, float *wg) {
  for (int r = 0; r < NR; ++r)
    for (int c = 0; c < NC; ++c) {
      const int cell = r * NC + c;
      for (int f = 0; f < NPLANE; ++f) {
        const int src = VIDX(f, c) + r, dst = cell * NPLANE + f;
        if (frep) frep[dst] = (int64_t)s->xf[src];
        if (w) w[dst] = s->w[src];
        if (wg) wg[dst] = s->wg[src];
      }
      if (grid)
        for (int ch = 0; ch < NCHAN; ++ch) grid[cell * NCHAN + ch] = (uint8_t)((s->gm[cell][ch >> 6] >> (ch & 63)) & 1ULL);
    }
}

void cpf_read(const cpf_sim *s, uint8_t *grid, int64_t *frep, float *w, float *wg, float *avg_reward,
              int64_t *iter, orc_event *next_event, int64_t stats8[8]) {
  cpf_unpack(s, grid, frep, w, wg);
  if (avg_reward) *avg_reward = s->avg_reward;
  if (iter) *iter = s->iter;
  if (next_event) *next_event = s->event;
  if (stats8) memcpy(stats8, s->stats, sizeof s->stats);
}
uint64_t cpf_ndraws(const cpf_sim *s) { return orc_eg_ndraws(s->eg); }

/* n_replicas independent simulators over OpenMP (Philox stream id = first_replica + i) */
int64_t cpf_run_replicas(const orc_opt *base, int64_t first_replica, int n_replicas, int64_t n_events,
                         int n_threads, int64_t stats_sum8[8], float *avg_reward_out, double *bp_out) {
  int64_t total = 0;
  int64_t sums[8] = {0};
  cpf_build_tables();
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
#endif
  for (int i = 0; i < n_replicas; ++i) {
    orc_opt o = *base;
    o.replica = (uint64_t)(first_replica + i);
    o.n_events = n_events;
    o.order = ORC_ORDER_FAST;
    cpf_sim *s = cpf_create(&o);
    int64_t done = 0;
    cpf_run(s, n_events, NULL, 0, &done);
    total += done;
    if (avg_reward_out) avg_reward_out[i] = s->avg_reward;
    if (bp_out) {
      double cum[3];
      orc_stats_cumus(s->stats, cum);
      bp_out[i] = cum[0];
    }
#ifdef _OPENMP
#pragma omp critical(cpf_sum)
#endif
    for (int k = 0; k < 8; ++k) sums[k] += s->stats[k];
    cpf_destroy(s);
  }
  if (stats_sum8) memcpy(stats_sum8, sums, sizeof sums);
  return total;
}
