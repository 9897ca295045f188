/*
 * dca_oracle.c -- CPU ORACLE (test infrastructure; see dca_oracle.h header).
 *
 * Plain-C restatement of tsoernes/haskelldca's per-event hot path.  Citations
 * are file:line into the reference tree.  Compile with -ffp-contract=off: the
 * only fused multiply-adds are the explicit fmaf() calls of ORC_ORDER_FAST.
 */
#include "dca_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define R_ ORC_ROWS
#define C_ ORC_COLS
#define CH_ ORC_CHANNELS
#define NF_ ORC_NFEATS
#define GIDX(r, c, ch) (((r) * C_ + (c)) * CH_ + (ch))
#define FIDX(r, c, f) (((r) * C_ + (c)) * NF_ + (f))

/* ------------------------------------------------------------------------ */
/* Opt defaults: src/Opt.hs:28-76                                           */
/* ------------------------------------------------------------------------ */
void orc_opt_defaults(orc_opt *o) {
  memset(o, 0, sizeof *o);
  o->call_dur = 3.0;
  o->call_dur_hoff = 1.0;
  o->call_rate = 200.0;
  o->hoff_prob = 0.0;
  o->n_events = 10000;
  o->log_iter = 1000;
  o->alpha_net = 2.52e-6f;
  o->alpha_avg = 0.06f;
  o->alpha_grad = 5e-6f;
  o->min_loss = 0.0f;
  o->seed = 0;
  o->order = ORC_ORDER_REF;
  o->rng_mode = ORC_RNG_MT19937_64;
  o->replica = 0;
}

/* ------------------------------------------------------------------------ */
/* Neighbourhoods: src/Gridneighs.hs:92-128                                 */
/* ------------------------------------------------------------------------ */
static int is_valid(int r, int c) { /* Gridneighs.hs:124-128 */
  return r >= 0 && r < R_ && c >= 0 && c < C_;
}

/* Gridneighs.hs:111-121.  Six hexagon sides of length d in axial coordinates;
 * concatenated ps6++ps5++ps4++ps3++ps2++ps1, then bounds-filtered. */
int orc_periphery(int d, int rF, int cF, int *out) {
  /* start points and steps of ps1..ps6 */
  const int sr[6] = {rF - d, rF - d, rF, rF + d, rF + d, rF};
  const int sc[6] = {cF, cF + d, cF + d, cF, cF - d, cF - d};
  const int dr[6] = {0, 1, 1, 0, -1, -1};
  const int dc[6] = {1, 0, -1, -1, 0, 1};
  int n = 0;
  for (int side = 5; side >= 0; --side) { /* ps6 first, ps1 last */
    int r = sr[side], c = sc[side];
    for (int k = 0; k < d; ++k) {
      if (is_valid(r, c)) {
        out[2 * n] = r;
        out[2 * n + 1] = c;
        ++n;
      }
      r += dr[side];
      c += dc[side];
    }
  }
  return n;
}

/* Gridneighs.hs:92-98 */
int orc_get_neighs(int d, int r, int c, int include_self, int *out) {
  int n = 0;
  if (include_self) {
    out[0] = r;
    out[1] = c;
    n = 1;
  }
  for (int dd = 1; dd <= d; ++dd) n += orc_periphery(dd, r, c, out + 2 * n);
  return n;
}

/* Lookup tables V (_neighs), V' (_neighsO), M (_nNeighs), _neighs2:
 * Gridneighs.hs:57-89, Gridconsts.hs:24-45 */
static int g_tables_ready = 0;
static int g_V[49 * 61][2], g_nV;     /* d<=4 incl. self; <= 61 cells per hood */
static int g_VO[49 * 60][2], g_nVO;
static int g_V2[49 * 19][2], g_nV2;   /* d<=2 incl. self */
static int g_M[49][5];                /* [start, n1..n4] (n_d counts self) */
static int g_M2[49][2];               /* start,len in V2 */

static void build_tables(void) {
  if (g_tables_ready) return;
#ifdef _OPENMP
#pragma omp critical(orc_tables)
#endif
  {
    if (!g_tables_ready) {
      int buf[2 * 64];
      g_nV = g_nVO = g_nV2 = 0;
      for (int r = 0; r < R_; ++r)
        for (int c = 0; c < C_; ++c) {
          int cell = r * C_ + c;
          /* neighborhood 4 idx, Gridneighs.hs:102-107 */
          g_M[cell][0] = g_nV;
          g_V[g_nV][0] = r;
          g_V[g_nV][1] = c;
          g_nV++;
          int cum = 1;
          for (int d = 1; d <= 4; ++d) {
            int k = orc_periphery(d, r, c, buf);
            for (int i = 0; i < k; ++i) {
              g_V[g_nV][0] = buf[2 * i];
              g_V[g_nV][1] = buf[2 * i + 1];
              g_nV++;
              g_VO[g_nVO][0] = buf[2 * i];
              g_VO[g_nVO][1] = buf[2 * i + 1];
              g_nVO++;
            }
            cum += k;
            g_M[cell][d] = cum; /* tail (scanl (+) 1 lengths) */
          }
          /* neighborhood 2 idx, Gridneighs.hs:80-89 */
          g_M2[cell][0] = g_nV2;
          int k2 = orc_get_neighs(2, r, c, 1, buf);
          for (int i = 0; i < k2; ++i) {
            g_V2[g_nV2][0] = buf[2 * i];
            g_V2[g_nV2][1] = buf[2 * i + 1];
            g_nV2++;
          }
          g_M2[cell][1] = k2;
        }
      g_tables_ready = 1;
    }
  }
}

void orc_table_sizes(int *n_neighs, int *n_neighs_o, int *n_neighs2, int64_t *m_out) {
  build_tables();
  if (n_neighs) *n_neighs = g_nV;
  if (n_neighs_o) *n_neighs_o = g_nVO;
  if (n_neighs2) *n_neighs2 = g_nV2;
  if (m_out)
    for (int i = 0; i < 49; ++i)
      for (int j = 0; j < 5; ++j) m_out[i * 5 + j] = g_M[i][j];
}

/* Gridconsts.hs:52-71: slit (start + !self) (n_d - !self) V */
static int hood(int d, int r, int c, int include_self, const int (**cells)[2]) {
  build_tables();
  int cell = r * C_ + c;
  int start = g_M[cell][0] + (include_self ? 0 : 1);
  int n = g_M[cell][d] - (include_self ? 0 : 1);
  *cells = &g_V[start];
  return n;
}

int orc_get_neighborhood_acc(int d, int r, int c, int include_self, int *out) {
  const int(*cells)[2];
  int n = hood(d, r, c, include_self, &cells);
  for (int i = 0; i < n; ++i) {
    out[2 * i] = cells[i][0];
    out[2 * i + 1] = cells[i][1];
  }
  return n;
}

/* ------------------------------------------------------------------------ */
/* Grid functions: src/Gridfuncs.hs                                         */
/* ------------------------------------------------------------------------ */

/* eligibleMap, Gridfuncs.hs:69-76: not (OR over {cell} u N2(cell)) */
static void eligible_map(const uint8_t *grid, int r, int c, uint8_t *emap) {
  const int(*nb)[2];
  int n = hood(2, r, c, 1, &nb);
  for (int ch = 0; ch < CH_; ++ch) {
    int used = 0;
    for (int i = 0; i < n; ++i) used |= grid[GIDX(nb[i][0], nb[i][1], ch)];
    emap[ch] = !used;
  }
}

/* indicesOf, AccUtils.hs:233-240: ascending indices of True entries */
static int indices_of(const uint8_t *bits, int n, int64_t *out) {
  int k = 0;
  for (int i = 0; i < n; ++i)
    if (bits[i]) out[k++] = i;
  return k;
}

int orc_eligible_chs(const uint8_t *grid, int r, int c, int64_t *chs) { /* :81-82 */
  uint8_t emap[CH_];
  eligible_map(grid, r, c, emap);
  return indices_of(emap, CH_, chs);
}

int orc_inuse_chs(const uint8_t *grid, int r, int c, int64_t *chs) { /* :86-87 */
  return indices_of(&grid[GIDX(r, c, 0)], CH_, chs);
}

/* Gridfuncs.hs:94-101: a channel used at a focal cell and at any d<=2 neighbour */
int orc_violates_reuse_constraint(const uint8_t *grid) {
  build_tables();
  for (int cell = 0; cell < 49; ++cell) {
    const int(*nb)[2] = &g_V2[g_M2[cell][0]];
    int n = g_M2[cell][1];
    for (int ch = 0; ch < CH_; ++ch) {
      int focal = grid[GIDX(nb[0][0], nb[0][1], ch)];
      int any = 0;
      for (int i = 1; i < n; ++i) any |= grid[GIDX(nb[i][0], nb[i][1], ch)];
      if (focal && any) return 1;
    }
  }
  return 0;
}

void orc_execute_action(int is_end, int r, int c, int ch, uint8_t *grid) { /* :105-110 */
  grid[GIDX(r, c, ch)] = is_end ? 0 : 1;
}

void orc_afterstates(const uint8_t *grid, int r, int c, int is_end, const int64_t *chs, int n,
                     uint8_t *out) { /* :116-123 */
  for (int k = 0; k < n; ++k) {
    memcpy(out + (size_t)k * ORC_GRID_SIZE, grid, ORC_GRID_SIZE);
    out[(size_t)k * ORC_GRID_SIZE + GIDX(r, c, (int)chs[k])] = is_end ? 0 : 1;
  }
}

int64_t orc_bool_sum(const uint8_t *grid) { /* AccUtils.hs:166-167 */
  int64_t s = 0;
  for (int i = 0; i < ORC_GRID_SIZE; ++i) s += grid[i] ? 1 : 0;
  return s;
}

/* featureRep, Gridfuncs.hs:153-170 (segmented folds over _neighsO / _neighs2) */
void orc_feature_rep(const uint8_t *grid, int64_t *frep) {
  build_tables();
  for (int r = 0; r < R_; ++r)
    for (int c = 0; c < C_; ++c) {
      int cell = r * C_ + c;
      /* segment of V' for this cell: d<=4 neighbours, focal cell excluded */
      const int(*nb)[2] = &g_V[g_M[cell][0] + 1];
      int n = g_M[cell][4] - 1;
      for (int ch = 0; ch < CH_; ++ch) {
        int64_t s = 0;
        for (int i = 0; i < n; ++i) s += grid[GIDX(nb[i][0], nb[i][1], ch)];
        frep[FIDX(r, c, ch)] = s;
      }
      const int(*nb2)[2] = &g_V2[g_M2[cell][0]];
      int n2 = g_M2[cell][1];
      int64_t nelig = 0;
      for (int ch = 0; ch < CH_; ++ch) {
        int used = 0;
        for (int i = 0; i < n2; ++i) used |= grid[GIDX(nb2[i][0], nb2[i][1], ch)];
        nelig += !used;
      }
      frep[FIDX(r, c, CH_)] = nelig;
    }
}

/* featureRep', Gridfuncs.hs:132-150: one cell at a time from host-side getNeighs */
void orc_feature_rep_slow(const uint8_t *grid, int64_t *frep) {
  int buf[2 * 64];
  memset(frep, 0, sizeof(int64_t) * ORC_FREP_SIZE);
  for (int r = 0; r < R_; ++r)
    for (int c = 0; c < C_; ++c) {
      int n = orc_get_neighs(4, r, c, 0, buf);
      for (int i = 0; i < n; ++i)
        for (int ch = 0; ch < CH_; ++ch)
          frep[FIDX(r, c, ch)] += grid[GIDX(buf[2 * i], buf[2 * i + 1], ch)] ? 1 : 0;
      uint8_t emap[CH_];
      eligible_map(grid, r, c, emap);
      int64_t s = 0;
      for (int ch = 0; ch < CH_; ++ch) s += emap[ch];
      frep[FIDX(r, c, CH_)] = s;
    }
}

/* incAfterStateFreps, Gridfuncs.hs:186-252 -- literal, including the nested
 * neighbour scan of :220-232 and the zgrid of :202-203 */
void orc_inc_afterstate_freps(int r, int c, int is_end, const int64_t *chs, int n,
                              const uint8_t *grid, const int64_t *frep, int64_t *out) {
  build_tables();
  const int diff = is_end ? -1 : 1; /* :201 */
  uint8_t gridp[ORC_GRID_SIZE];     /* grid' :202-203 */
  memcpy(gridp, grid, ORC_GRID_SIZE);
  if (is_end)
    for (int k = 0; k < n; ++k) gridp[GIDX(r, c, (int)chs[k])] = 0;
  const int(*n2)[2];
  const int(*n4)[2];
  int k2 = hood(2, r, c, 1, &n2); /* :207 */
  int k4 = hood(4, r, c, 0, &n4); /* :208 */
  for (int k = 0; k < n; ++k) {
    int64_t *af = out + (size_t)k * ORC_FREP_SIZE;
    memcpy(af, frep, sizeof(int64_t) * ORC_FREP_SIZE); /* replicate :206 */
    int ch = (int)chs[k];
    for (int i = 0; i < k4; ++i) af[FIDX(n4[i][0], n4[i][1], ch)] += diff; /* :212-217 */
    for (int i = 0; i < k2; ++i) {                                       /* :220-232 */
      int cell2 = n2[i][0] * C_ + n2[i][1];
      int start = g_M[cell2][0], cnt = g_M[cell2][2];
      int acc = 0;
      for (int j = 0; j < cnt; ++j) acc |= gridp[GIDX(g_V[start + j][0], g_V[start + j][1], ch)];
      if (!acc) af[FIDX(n2[i][0], n2[i][1], CH_)] += -diff; /* :222, :243-252 */
    }
  }
}

/* ------------------------------------------------------------------------ */
/* fp32 reductions                                                          */
/* ------------------------------------------------------------------------ */

/* REF: Accelerate Interpreter `fold (+) 0 (zipWith (*) xs ys)` (AccUtils.hs:219-228)
 * = right fold x0 + (x1 + (... + (x_{n-1} + 0))), product rounded, then sum. */
static float dot_ref(const int64_t *x, const float *w) {
  float acc = 0.0f;
  for (int i = ORC_FREP_SIZE - 1; i >= 0; --i) {
    float p = (float)x[i] * w[i];
    acc = p + acc;
  }
  return acc;
}

/* FAST canonical trees (see header). Flatten index i = (r*7+c)*71 + f.  The
 * shapes follow the thread layout of the sm_100a kernel (96 "agent" threads =
 * 12 groups of 8 lanes; lane r of a group owns row r of one feature plane per
 * slot, plane f = 12*slot + group), so that every partial sum is either
 * thread-local or one xor-shuffle away:
 *   rowdot(f,r) = e + o with two fma chains over the even / odd columns
 *                 (what one f32x2 accumulator pair computes):
 *                 e = x0*w0; e = fma(x2,w2,e); e = fma(x4,w4,e); e = fma(x6,w6,e)
 *                 o = x1*w1; o = fma(x3,w3,o); o = fma(x5,w5,o)
 *   plane(f)    = bfly8(rowdot(f,0..6), 0)        -- xor-butterfly 1,2,4 =
 *                 ((a0+a1)+(a2+a3)) + ((a4+a5)+(a6+a7))
 *   x.w   (V):  v[l] = (plane(l) + plane(l+32)) + (l < 6 ? plane(l+64) : 0), l = 0..31
 *               xor-butterfly 1,2,4,8,16 over v[0..31], then + plane(70)
 *   x.wg:       thread (group g, row r): s = rowdot(g,r); s += rowdot(12j+g, r),
 *               j = 1..5 while 12j+g <= 70; row 7 = 0;  warp sums by
 *               xor-butterfly 1..16 over the 32 lanes (4 groups) of each of the
 *               3 warps; total = (W0 + W1) + W2. */
static float tree8(const float a[8]) {
  return ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
}
static float rowdot_fast(const int64_t *x, const float *w, int f, int r) {
  float e = (float)x[FIDX(r, 0, f)] * w[FIDX(r, 0, f)];
  float o = (float)x[FIDX(r, 1, f)] * w[FIDX(r, 1, f)];
  e = fmaf((float)x[FIDX(r, 2, f)], w[FIDX(r, 2, f)], e);
  o = fmaf((float)x[FIDX(r, 3, f)], w[FIDX(r, 3, f)], o);
  e = fmaf((float)x[FIDX(r, 4, f)], w[FIDX(r, 4, f)], e);
  o = fmaf((float)x[FIDX(r, 5, f)], w[FIDX(r, 5, f)], o);
  e = fmaf((float)x[FIDX(r, 6, f)], w[FIDX(r, 6, f)], e);
  return e + o;
}
static float plane_fast(const int64_t *x, const float *w, int f) {
  float a[8];
  for (int r = 0; r < R_; ++r) a[r] = rowdot_fast(x, w, f, r);
  a[7] = 0.0f;
  return tree8(a);
}
static void bfly32(float v[32]) { /* xor-butterfly 1,2,4,8,16: every lane ends with the same sum */
  for (int s = 1; s < 32; s <<= 1) {
    float t[32];
    for (int l = 0; l < 32; ++l) t[l] = v[l] + v[l ^ s];
    memcpy(v, t, sizeof t);
  }
}
static float dot_fast(const int64_t *x, const float *w) {
  float v[32];
  for (int l = 0; l < 32; ++l) {
    float p2 = (l < 6) ? plane_fast(x, w, l + 64) : 0.0f;
    v[l] = (plane_fast(x, w, l) + plane_fast(x, w, l + 32)) + p2;
  }
  bfly32(v);
  return v[0] + plane_fast(x, w, CH_);
}
static float dot_fast_g(const int64_t *x, const float *wg) {
  float W[3];
  for (int wp = 0; wp < 3; ++wp) {
    float v[32];
    for (int l = 0; l < 32; ++l) {
      const int g = 4 * wp + (l >> 3), r = l & 7;
      float s = 0.0f;
      if (r < R_) {
        s = rowdot_fast(x, wg, g, r);
        for (int j = 1; j < 6 && 12 * j + g <= CH_; ++j) s = s + rowdot_fast(x, wg, 12 * j + g, r);
      }
      v[l] = s;
    }
    bfly32(v);
    W[wp] = v[0];
  }
  return (W[0] + W[1]) + W[2];
}

float orc_dense_dot(int order, const int64_t *frep, const float *w) {
  return order == ORC_ORDER_FAST ? dot_fast(frep, w) : dot_ref(frep, w);
}
/* x . wGradCorr (Agent.hs:66): same right fold in REF, its own tree in FAST */
float orc_dense_dot_g(int order, const int64_t *frep, const float *wg) {
  return order == ORC_ORDER_FAST ? dot_fast_g(frep, wg) : dot_ref(frep, wg);
}

float orc_forward(int order, const int64_t *frep, const float *w) { /* Agent.hs:15-23 */
  return orc_dense_dot(order, frep, w);
}

/* argpmax, AccUtils.hs:203-208: fn a b = cond (snd a > snd b) a b folded over
 * (index,value) pairs => the maximum value, and among equal maxima the LAST
 * index.  (Right fold as the Interpreter does; left fold gives the same for
 * non-NaN input.) */
int orc_argpmax(const float *q, int n) {
  int best = n - 1;
  for (int i = n - 2; i >= 0; --i)
    if (q[i] > q[best]) best = i; /* a = element i, b = running result */
  return best;
}

/* selectAction, Simulator.hs:185-205 */
int orc_select_action(int order, int r, int c, int is_end, const int64_t *chs, int n,
                      const float *w, const uint8_t *grid, const int64_t *frep,
                      float *qval_out, int64_t *afrep_out, float *qvals_out) {
  int64_t *afreps = (int64_t *)malloc(sizeof(int64_t) * ORC_FREP_SIZE * (size_t)n);
  float q[CH_];
  orc_inc_afterstate_freps(r, c, is_end, chs, n, grid, frep, afreps); /* :188 */
  for (int k = 0; k < n; ++k)                                          /* mvMul :198 */
    q[k] = orc_dense_dot(order, afreps + (size_t)k * ORC_FREP_SIZE, w);
  int idx = orc_argpmax(q, n); /* :202 */
  if (qval_out) *qval_out = q[idx];
  if (qvals_out) memcpy(qvals_out, q, sizeof(float) * (size_t)n);
  if (afrep_out)
    memcpy(afrep_out, afreps + (size_t)idx * ORC_FREP_SIZE, sizeof(int64_t) * ORC_FREP_SIZE);
  free(afreps);
  return idx;
}

/* getAction, Simulator.hs:154-172 */
int orc_get_action(int order, int r, int c, int is_end, const float *w, const uint8_t *grid,
                   const int64_t *frep, int64_t *next_frep, int *ncand_out) {
  int64_t chs[CH_];
  int n = is_end ? orc_inuse_chs(grid, r, c, chs) : orc_eligible_chs(grid, r, c, chs); /* :161-165 */
  if (ncand_out) *ncand_out = n;
  if (n == 0) { /* :170-171 */
    memcpy(next_frep, frep, sizeof(int64_t) * ORC_FREP_SIZE);
    return -1;
  }
  int idx = orc_select_action(order, r, c, is_end, chs, n, w, grid, frep, NULL, next_frep, NULL);
  return (int)chs[idx];
}

/* hand-off look-ahead (extension, see dca_oracle.h) */
int orc_get_action_lookahead(int order, int r, int c, int hr, int hc, const float *w,
                             const uint8_t *grid, const int64_t *frep, int64_t *next_frep,
                             int *ncand_out, float *qvals_out) {
  int64_t chs[CH_], hs[CH_];
  const int n = orc_inuse_chs(grid, r, c, chs);
  if (ncand_out) *ncand_out = n;
  if (n == 0) {
    memcpy(next_frep, frep, sizeof(int64_t) * ORC_FREP_SIZE);
    return -1;
  }
  int64_t *afreps = (int64_t *)malloc(sizeof(int64_t) * ORC_FREP_SIZE * (size_t)n);
  int64_t *freps2 = (int64_t *)malloc(sizeof(int64_t) * ORC_FREP_SIZE * (size_t)CH_);
  float q[CH_];
  orc_inc_afterstate_freps(r, c, 1, chs, n, grid, frep, afreps);
  for (int k = 0; k < n; ++k) {
    const int64_t *af = afreps + (size_t)k * ORC_FREP_SIZE;
    uint8_t gk[ORC_GRID_SIZE];
    memcpy(gk, grid, ORC_GRID_SIZE);
    orc_execute_action(1, r, c, (int)chs[k], gk);
    const int m = orc_eligible_chs(gk, hr, hc, hs);
    if (m == 0) {
      q[k] = orc_dense_dot(order, af, w);
    } else {
      orc_inc_afterstate_freps(hr, hc, 0, hs, m, gk, af, freps2);
      float best = orc_dense_dot(order, freps2, w);
      for (int j = 1; j < m; ++j) {
        const float v = orc_dense_dot(order, freps2 + (size_t)j * ORC_FREP_SIZE, w);
        if (v > best) best = v;
      }
      q[k] = best;
    }
  }
  const int idx = orc_argpmax(q, n);
  if (qvals_out) memcpy(qvals_out, q, sizeof(float) * (size_t)n);
  memcpy(next_frep, afreps + (size_t)idx * ORC_FREP_SIZE, sizeof(int64_t) * ORC_FREP_SIZE);
  free(afreps);
  free(freps2);
  return (int)chs[idx];
}

/* backward, Agent.hs:44-81 */
float orc_backward(int order, float aN, float aA, float aG, const int64_t *frep,
                   const int64_t *next_frep, float reward, float *w, float *wg,
                   float *avg_reward, int *is_nan) {
  const float avgR = *avg_reward;
  const float val = orc_dense_dot(order, frep, w);           /* :62 */
  const float next_val = orc_dense_dot(order, next_frep, w); /* :63 */
  const float td = ((reward - avgR) + next_val) - val;       /* :65 */
  const float dot = orc_dense_dot_g(order, frep, wg);        /* :66 */
  const float c = -2.0f * aN;                                /* :67 */
  const float ctd = c * td, cavg = c * avgR, cdot = c * dot;
  const float s = aG * (td - dot); /* :75 */
  if (order == ORC_ORDER_FAST) {
    const float ncdot = -cdot;
    for (int i = 0; i < ORC_FREP_SIZE; ++i) {
      const float x = (float)frep[i], xn = (float)next_frep[i];
      const float g = fmaf(ncdot, xn, fmaf(ctd, x, cavg));
      w[i] = w[i] - g;
      wg[i] = fmaf(s, x, wg[i]);
    }
  } else {
    for (int i = 0; i < ORC_FREP_SIZE; ++i) {
      const float x = (float)frep[i], xn = (float)next_frep[i];
      const float a1 = ctd * x;        /* :68 */
      const float a3 = cdot * xn;      /* :70 */
      const float g = (a1 + cavg) - a3; /* :71 */
      w[i] = w[i] - g;                 /* :73 */
      const float corr = s * x;        /* :75 */
      wg[i] = wg[i] + corr;            /* :76 */
    }
  }
  *avg_reward = avgR + aA * td; /* :78 */
  if (is_nan) *is_nan = isnan(td) ? 1 : 0; /* :80 */
  return td;
}

/* gridStep + gridStepTrain, Simulator.hs:137-144, SimRunner.hs:100-118 */
float orc_grid_step_train(int order, float aN, float aA, float aG, int is_end, int r, int c,
                          int ch, const int64_t *frep, const int64_t *next_frep, uint8_t *grid,
                          float *w, float *wg, float *avg_reward, int *is_nan,
                          int64_t *reward_out) {
  if (ch >= 0) orc_execute_action(is_end, r, c, ch, grid);
  int64_t reward = orc_bool_sum(grid);
  if (reward_out) *reward_out = reward;
  return orc_backward(order, aN, aA, aG, frep, next_frep, (float)reward, w, wg, avg_reward,
                      is_nan);
}

/* ------------------------------------------------------------------------ */
/* RNG                                                                      */
/* ------------------------------------------------------------------------ */
/* mersenne-random-pure64 0.2.2.0 (pinned by lts-12.10, stack.yaml:23; NOT in the
 * reference tree): standard MT19937-64 of Matsumoto & Nishimura (init_genrand64,
 * genrand64_int64); randomDouble = (w >> 11) / 2^53.  Call sites:
 * EventGen/Internal.hs:21,57-66. */
#define MT_NN 312
#define MT_MM 156
struct orc_rng {
  int mode;
  uint64_t mt[MT_NN];
  int mti;
  uint32_t key[2];
  uint64_t replica;
  uint64_t ndraws;
};

static void mt_seed(orc_rng *g, uint64_t seed) {
  g->mt[0] = seed;
  for (int i = 1; i < MT_NN; ++i)
    g->mt[i] = 6364136223846793005ULL * (g->mt[i - 1] ^ (g->mt[i - 1] >> 62)) + (uint64_t)i;
  g->mti = MT_NN;
}

static uint64_t mt_next(orc_rng *g) {
  static const uint64_t mag01[2] = {0ULL, 0xB5026F5AA96619E9ULL};
  const uint64_t UM = 0xFFFFFFFF80000000ULL, LM = 0x7FFFFFFFULL;
  uint64_t x;
  if (g->mti >= MT_NN) {
    int i;
    for (i = 0; i < MT_NN - MT_MM; ++i) {
      x = (g->mt[i] & UM) | (g->mt[i + 1] & LM);
      g->mt[i] = g->mt[i + MT_MM] ^ (x >> 1) ^ mag01[(int)(x & 1ULL)];
    }
    for (; i < MT_NN - 1; ++i) {
      x = (g->mt[i] & UM) | (g->mt[i + 1] & LM);
      g->mt[i] = g->mt[i + (MT_MM - MT_NN)] ^ (x >> 1) ^ mag01[(int)(x & 1ULL)];
    }
    x = (g->mt[MT_NN - 1] & UM) | (g->mt[0] & LM);
    g->mt[MT_NN - 1] = g->mt[MT_MM - 1] ^ (x >> 1) ^ mag01[(int)(x & 1ULL)];
    g->mti = 0;
  }
  x = g->mt[g->mti++];
  x ^= (x >> 29) & 0x5555555555555555ULL;
  x ^= (x << 17) & 0x71D67FFFEDA60000ULL;
  x ^= (x << 37) & 0xFFF7EEE000000000ULL;
  x ^= (x >> 43);
  return x;
}

/* Philox4x32-10 (Salmon et al., SC'11): counter-based generator for the
 * batched build; word(d) = out[0] | out[1]<<32 of
 * philox(ctr = (d_lo, d_hi, replica_lo, replica_hi), key = (seed_lo, seed_hi)). */
void orc_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
  uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
  uint32_t k0 = key_in[0], k1 = key_in[1];
  for (int round = 0; round < 10; ++round) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

orc_rng *orc_rng_create(int mode, uint64_t seed, uint64_t replica) {
  orc_rng *g = (orc_rng *)calloc(1, sizeof *g);
  g->mode = mode;
  g->replica = replica;
  g->key[0] = (uint32_t)seed;
  g->key[1] = (uint32_t)(seed >> 32);
  if (mode == ORC_RNG_MT19937_64) mt_seed(g, seed); /* pureMT seed */
  return g;
}
void orc_rng_destroy(orc_rng *g) { free(g); }
uint64_t orc_rng_ndraws(const orc_rng *g) { return g->ndraws; }

uint64_t orc_rng_word64(orc_rng *g) { /* getRandomWord64, Internal.hs:60 */
  uint64_t d = g->ndraws++;
  if (g->mode == ORC_RNG_MT19937_64) return mt_next(g);
  uint32_t ctr[4] = {(uint32_t)d, (uint32_t)(d >> 32), (uint32_t)g->replica,
                     (uint32_t)(g->replica >> 32)};
  uint32_t out[4];
  orc_philox4x32_10(ctr, g->key, out);
  return (uint64_t)out[0] | ((uint64_t)out[1] << 32);
}

static double word_to_double(uint64_t w) { /* randomDouble: (w `div` 2048) / 2^53 */
  return (double)(w >> 11) * (1.0 / 9007199254740992.0);
}

double orc_rng_double(orc_rng *g) { /* getRandomDouble, Internal.hs:61 */
  return word_to_double(orc_rng_word64(g));
}

/* -log(u) for the Philox mode, specified operation by operation so that the CUDA
 * kernel can reproduce it bit for bit (libm `log` differs between host and
 * device):  u = k * 2^-53, k = word >> 11;  k == 0 -> +inf.
 *   u = m * 2^e with m in [sqrt(1/2), sqrt(2)), via the IEEE bit pattern;
 *   s = (m-1)/(m+1); z = s*s;
 *   p = Horner over z with fma of c_j = 1/(2j+1), j = 10..0;
 *   t = s*p; t = t+t;              (log m)
 *   result = -fma((double)e, LN2, t)                                         */
double orc_neglog_u53(uint64_t word) {
  uint64_t k = word >> 11;
  if (k == 0) return INFINITY;
  double u = (double)k * (1.0 / 9007199254740992.0);
  uint64_t bits;
  memcpy(&bits, &u, 8);
  int e = (int)((bits >> 52) & 0x7FF) - 1022; /* u = m * 2^e, m in [0.5,1) */
  bits = (bits & 0x000FFFFFFFFFFFFFULL) | 0x3FE0000000000000ULL;
  double m;
  memcpy(&m, &bits, 8);
  if (m < 0.70710678118654752440) {
    m = m + m;
    e -= 1;
  }
  double s = (m - 1.0) / (m + 1.0);
  double z = s * s;
  double p = 1.0 / 21.0;
  p = fma(p, z, 1.0 / 19.0);
  p = fma(p, z, 1.0 / 17.0);
  p = fma(p, z, 1.0 / 15.0);
  p = fma(p, z, 1.0 / 13.0);
  p = fma(p, z, 1.0 / 11.0);
  p = fma(p, z, 1.0 / 9.0);
  p = fma(p, z, 1.0 / 7.0);
  p = fma(p, z, 1.0 / 5.0);
  p = fma(p, z, 1.0 / 3.0);
  p = fma(p, z, 1.0);
  double t = s * p;
  t = t + t;
  return -fma((double)e, 0.69314718055994530942, t);
}

/* random-fu 0.2.7.0 `exponential mu` (NOT in the reference tree):
 * floatingExponential = do x <- stdUniformT; return (negate (log x) * mu), one
 * getRandomDouble.  Call sites: EventGen.hs:83,102, Internal.hs:85. */
double orc_rng_exponential(orc_rng *g, double mean) {
  uint64_t w = orc_rng_word64(g);
  if (g->mode == ORC_RNG_MT19937_64) {
    double x = word_to_double(w);
    return (-log(x)) * mean;
  }
  return orc_neglog_u53(w) * mean;
}

/* random-fu `uniform lo hi :: Int` = integralUniform (NOT in the reference
 * tree; call site EventGen.hs:104): m = hi-lo+1 <= 256 needs one random byte z,
 * redrawn while z < 256 mod m, result lo + z mod m.  ASSUMPTION (unpinned, see
 * SURVEY 8c): the byte is the low 8 bits of one getRandomWord64. */
int orc_rng_uniform_int(orc_rng *g, int lo, int hi) {
  if (lo > hi) {
    int t = lo; lo = hi; hi = t;
  }
  int m = hi - lo + 1;
  int n_reject = 256 % m;
  for (;;) {
    int z = (int)(orc_rng_word64(g) & 0xFF);
    if (z >= n_reject) return lo + z % m;
  }
}

/* ------------------------------------------------------------------------ */
/* EventGen: src/EventGen.hs, src/EventGen/Internal.hs                      */
/* ------------------------------------------------------------------------ */
typedef struct {
  double time;
  int64_t id;
  int slot;
} heap_ent;

struct orc_eventgen {
  int64_t next_id;        /* egId, Internal.hs:46 */
  heap_ent *heap;         /* egQueue (MinHeap EventKey) */
  int heap_n, heap_cap;
  orc_event *slots;       /* egEvents (Map EventId Event) as a slot pool */
  int *free_slots;
  int n_free, slot_cap, n_live;
  int end_slot[49][CH_];  /* egEndIds (Map (Cell,Ch) EventId) */
  int n_endids;
  orc_rng *rng;           /* egGen */
  orc_opt opt;
};

static int key_less(const heap_ent *a, const heap_ent *b) { /* Internal.hs:29-34 */
  if (a->time < b->time) return 1;
  if (a->time > b->time) return 0;
  return a->id < b->id;
}

static void eg_grow(orc_eventgen *eg) {
  int old = eg->slot_cap;
  int cap = old ? old * 2 : 1024;
  eg->slots = (orc_event *)realloc(eg->slots, sizeof(orc_event) * (size_t)cap);
  eg->free_slots = (int *)realloc(eg->free_slots, sizeof(int) * (size_t)cap);
  eg->heap = (heap_ent *)realloc(eg->heap, sizeof(heap_ent) * (size_t)cap);
  for (int i = cap - 1; i >= old; --i) eg->free_slots[eg->n_free++] = i;
  eg->slot_cap = cap;
  eg->heap_cap = cap;
}

orc_eventgen *orc_eg_create(const orc_opt *opt, int with_initial) {
  orc_eventgen *eg = (orc_eventgen *)calloc(1, sizeof *eg);
  eg->opt = *opt;
  eg->rng = orc_rng_create(opt->rng_mode, opt->seed, opt->replica);
  for (int i = 0; i < 49; ++i)
    for (int ch = 0; ch < CH_; ++ch) eg->end_slot[i][ch] = -1;
  eg_grow(eg);
  if (with_initial) /* mkEventGen, EventGen.hs:54-58: mapM_ (generateNewEvent 0.0) gridIdxs */
    for (int r = 0; r < R_; ++r)
      for (int c = 0; c < C_; ++c) orc_eg_generate_new(eg, 0.0, r, c);
  return eg;
}

void orc_eg_destroy(orc_eventgen *eg) {
  if (!eg) return;
  orc_rng_destroy(eg->rng);
  free(eg->slots);
  free(eg->free_slots);
  free(eg->heap);
  free(eg);
}

int orc_eg_push(orc_eventgen *eg, const orc_event *ev) { /* Internal.hs:70-78 */
  if (eg->n_free == 0) eg_grow(eg);
  int slot = eg->free_slots[--eg->n_free];
  eg->slots[slot] = *ev;
  eg->n_live++;
  heap_ent e = {ev->time, eg->next_id, slot};
  int i = eg->heap_n++;
  while (i > 0) {
    int p = (i - 1) / 2;
    if (!key_less(&e, &eg->heap[p])) break;
    eg->heap[i] = eg->heap[p];
    i = p;
  }
  eg->heap[i] = e;
  if (ev->etype == ORC_END) {
    int cell = ev->r * C_ + ev->c;
    if (eg->end_slot[cell][ev->ch] < 0) eg->n_endids++;
    eg->end_slot[cell][ev->ch] = slot;
  }
  eg->next_id++;
  return 0;
}

int orc_eg_pop(orc_eventgen *eg, orc_event *out) { /* EventGen.hs:39-50 */
  if (eg->heap_n == 0) return -1; /* fromJust . Heap.view on an empty heap */
  heap_ent top = eg->heap[0];
  heap_ent last = eg->heap[--eg->heap_n];
  int i = 0;
  for (;;) {
    int l = 2 * i + 1, r = l + 1, m = i;
    const heap_ent *best = &last;
    if (l < eg->heap_n && key_less(&eg->heap[l], best)) {
      best = &eg->heap[l];
      m = l;
    }
    if (r < eg->heap_n && key_less(&eg->heap[r], best)) {
      best = &eg->heap[r];
      m = r;
    }
    if (m == i) break;
    eg->heap[i] = eg->heap[m];
    i = m;
  }
  if (eg->heap_n > 0) eg->heap[i] = last;
  *out = eg->slots[top.slot];
  eg->free_slots[eg->n_free++] = top.slot;
  eg->n_live--;
  if (out->etype == ORC_END) { /* :46-48 */
    int cell = out->r * C_ + out->c;
    if (eg->end_slot[cell][out->ch] >= 0) eg->n_endids--;
    eg->end_slot[cell][out->ch] = -1;
  }
  return 0;
}

int orc_eg_reassign(orc_eventgen *eg, int r, int c, int from_ch, int to_ch) { /* :62-72 */
  int cell = r * C_ + c;
  int slot = eg->end_slot[cell][from_ch];
  if (slot < 0) return -1; /* Map.! on a missing key */
  eg->end_slot[cell][from_ch] = -1;
  if (eg->end_slot[cell][to_ch] >= 0) eg->n_endids--; /* Map.insert overwrites */
  eg->end_slot[cell][to_ch] = slot;
  eg->slots[slot].ch = to_ch; /* END toCh (hoffCell kept) */
  return 0;
}

int orc_eg_generate_new(orc_eventgen *eg, double time, int r, int c) { /* :75-85 */
  double lam = eg->opt.call_rate;
  double lamp = 1.0 / (lam / 60.0);
  double dt = orc_rng_exponential(eg->rng, lamp);
  orc_event ev = {time + dt, ORC_NEW, r, c, -1, -1, -1};
  return orc_eg_push(eg, &ev);
}

int orc_eg_generate_hoff_new(orc_eventgen *eg, double time, int r, int c, int ch) { /* :94-109 */
  double dt = orc_rng_exponential(eg->rng, eg->opt.call_dur); /* callDurNew, :101-102 */
  int nb[2 * 32];
  int n = orc_get_neighs(2, r, c, 0, nb);                     /* :103 */
  int i = orc_rng_uniform_int(eg->rng, 0, n - 1);             /* :104 */
  double t = time + dt;
  orc_event e1 = {t, ORC_END, r, c, ch, nb[2 * i], nb[2 * i + 1]};
  orc_event e2 = {t, ORC_HOFF, nb[2 * i], nb[2 * i + 1], -1, -1, -1};
  orc_eg_push(eg, &e1);
  return orc_eg_push(eg, &e2);
}

static int generate_end(orc_eventgen *eg, double time, int r, int c, int ch, double lam) {
  double dt = orc_rng_exponential(eg->rng, lam); /* Internal.hs:81-86 */
  orc_event ev = {time + dt, ORC_END, r, c, ch, -1, -1};
  return orc_eg_push(eg, &ev);
}
int orc_eg_generate_end(orc_eventgen *eg, double time, int r, int c, int ch) {
  return generate_end(eg, time, r, c, ch, eg->opt.call_dur);
}
int orc_eg_generate_hoff_end(orc_eventgen *eg, double time, int r, int c, int ch) {
  return generate_end(eg, time, r, c, ch, eg->opt.call_dur_hoff);
}
int64_t orc_eg_id(const orc_eventgen *eg) { return eg->next_id; }
int orc_eg_sizes(const orc_eventgen *eg, int *heap_n, int *events_n, int *endids_n) {
  if (heap_n) *heap_n = eg->heap_n;
  if (events_n) *events_n = eg->n_live;
  if (endids_n) *endids_n = eg->n_endids;
  return 0;
}
uint64_t orc_eg_ndraws(const orc_eventgen *eg) { return orc_rng_ndraws(eg->rng); }
orc_rng *orc_eg_rng(orc_eventgen *eg) { return eg->rng; }

/* ------------------------------------------------------------------------ */
/* Simulator: src/Simulator.hs:59-134, src/SimRunner.hs:60-178              */
/* ------------------------------------------------------------------------ */
struct orc_sim {
  orc_opt opt;
  uint8_t grid[ORC_GRID_SIZE];
  int64_t frep[ORC_FREP_SIZE];
  int64_t next_frep[ORC_FREP_SIZE];
  float w[ORC_FREP_SIZE], wg[ORC_FREP_SIZE];
  float avg_reward;
  orc_event event;     /* ssEvent */
  orc_eventgen *eg;
  int64_t stats[8];    /* Stats.hs:15-30 */
  int64_t iter;        /* ssIter */
  int status;
  float loss_sum;
  int64_t loss_n;
};

static uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
uint64_t orc_grid_hash(const uint8_t *grid) {
  uint64_t h = 0;
  for (int i = 0; i < ORC_GRID_SIZE; ++i)
    if (grid[i]) h += mix64((uint64_t)i);
  return h;
}
uint64_t orc_frep_hash(const int64_t *frep) {
  uint64_t h = 0;
  for (int i = 0; i < ORC_FREP_SIZE; ++i) h += (uint64_t)frep[i] * mix64(0x10000ULL + (uint64_t)i);
  return h;
}

orc_sim *orc_sim_create(const orc_opt *opt) { /* mkSimState, Simulator.hs:88-97 */
  orc_sim *s = (orc_sim *)calloc(1, sizeof *s);
  s->opt = *opt;
  s->eg = orc_eg_create(opt, 1);
  orc_eg_pop(s->eg, &s->event);   /* :95 */
  orc_feature_rep(s->grid, s->frep); /* :96 */
  s->status = ORC_RUNNING;
  return s;
}
void orc_sim_destroy(orc_sim *s) {
  if (!s) return;
  orc_eg_destroy(s->eg);
  free(s);
}

/* environmentStep, Simulator.hs:101-134.  Stats hooks: Stats.hs:61-81. */
static int environment_step(orc_sim *s, int act) {
  const orc_event ev = s->event;
  const double time = ev.time;
  int64_t *st = s->stats;
  switch (ev.etype) {
    case ORC_NEW:
      st[6]++; st[0]++;                    /* statsEventArrivalNew */
      if (act >= 0) st[1]++;               /* statsEventAcceptNew */
      else { st[2]++; st[7]++; }           /* statsEventRejectNew */
      break;
    case ORC_HOFF:
      st[3]++;                             /* statsEventArrivalHoff */
      if (act < 0) { st[2]++; st[7]++; }   /* :113 books it as a rejected NEW */
      break;
    default:
      st[4]++;                             /* statsEventEnd */
  }
  switch (ev.etype) {
    case ORC_NEW:
      orc_eg_generate_new(s->eg, time, ev.r, ev.c); /* :119 */
      if (act >= 0) {                               /* :120-127 */
        double p = orc_rng_double(s->eg->rng);      /* stdUniform, drawn even if hoffProb==0 */
        if (p < s->opt.hoff_prob) orc_eg_generate_hoff_new(s->eg, time, ev.r, ev.c, act);
        else orc_eg_generate_end(s->eg, time, ev.r, ev.c, act);
      }
      break;
    case ORC_HOFF:
      if (act >= 0) orc_eg_generate_hoff_end(s->eg, time, ev.r, ev.c, act); /* :128 */
      break;
    default: { /* END toCh _ :129-130 */
      if (act < 0) return -1; /* fromJust Nothing */
      if (ev.ch != act)
        if (orc_eg_reassign(s->eg, ev.r, ev.c, act, ev.ch) != 0) return -1;
    }
  }
  s->iter += 1;                                    /* :131 */
  if (orc_eg_pop(s->eg, &s->event) != 0) return -1; /* :132-134 */
  return 0;
}

/* runStep (SimRunner.hs:60-97) followed by the checks of runPeriod's runStep'
 * (:132-169).  Returns a SimStop code; ORC_RUNNING = keep going. */
static int run_step(orc_sim *s, orc_trace *tr) {
  const orc_opt *o = &s->opt;
  const orc_event ev = s->event;
  const int is_end = ev.etype == ORC_END;
  int ncand = 0;
  int ch;
  if (o->hoff_lookahead && is_end && ev.hoff_r >= 0) /* extension: the departure half of a hand-off */
    ch = orc_get_action_lookahead(o->order, ev.r, ev.c, ev.hoff_r, ev.hoff_c, s->w, s->grid, s->frep,
                                  s->next_frep, &ncand, NULL);
  else
    ch = orc_get_action(o->order, ev.r, ev.c, is_end, s->w, s->grid, s->frep, s->next_frep,
                        &ncand); /* accFn1 :75 */
  if (environment_step(s, ch) != 0) return ORC_HOST_ERROR; /* :78 */
  int is_nan = 0;
  int64_t reward = 0;
  float loss = orc_grid_step_train(o->order, o->alpha_net, o->alpha_avg, o->alpha_grad, is_end,
                                   ev.r, ev.c, ch, s->frep, s->next_frep, s->grid, s->w, s->wg,
                                   &s->avg_reward, &is_nan, &reward); /* accFn2 :85-86 */
  memcpy(s->frep, s->next_frep, sizeof s->frep); /* ssFrep .= nextFrep :90 */
  if (tr) {
    memset(tr, 0, sizeof *tr);
    tr->time = ev.time;
    tr->loss = loss;
    tr->avg_reward = s->avg_reward;
    tr->reward = (int32_t)reward;
    tr->etype = (uint8_t)ev.etype;
    tr->cell = (uint8_t)(ev.r * C_ + ev.c);
    tr->ch = (int8_t)ch;
    tr->ncand = (uint8_t)ncand;
    tr->end_ch = (int8_t)(is_end ? ev.ch : -1);
    tr->grid_hash = orc_grid_hash(s->grid);
    tr->frep_hash = orc_frep_hash(s->frep);
  }
  if (is_nan) return ORC_NAN_LOSS; /* :152 */
  s->loss_sum = s->loss_sum + loss;
  s->loss_n++;
  if (o->verify_reuse_constraint && orc_violates_reuse_constraint(s->grid))
    return ORC_REUSE_VIOLATED; /* :155-156 */
  if (o->verify_nelig_bounds) { /* :150,158-159: maximum over the whole frep */
    int64_t mx = s->frep[0];
    for (int i = 1; i < ORC_FREP_SIZE; ++i)
      if (s->frep[i] > mx) mx = s->frep[i];
    if (mx > 70) return ORC_NELIG_OVERFLOW;
  }
  if (o->min_loss > 0 && s->iter > 50 && fabsf(loss) < o->min_loss) return ORC_ZERO_LOSS; /* :161 */
  if (s->iter == o->n_events) return ORC_SUCCESS; /* :164 */
  return ORC_RUNNING;
}

int orc_sim_run(orc_sim *s, int64_t max_steps, orc_trace *trace, int64_t trace_cap,
                int64_t *steps_done) {
  int64_t n = 0;
  while (n < max_steps && s->status == ORC_RUNNING) {
    orc_trace *tr = (trace && n < trace_cap) ? &trace[n] : NULL;
    s->status = run_step(s, tr);
    ++n;
  }
  if (steps_done) *steps_done = n;
  return s->status;
}

void orc_sim_read(const orc_sim *s, uint8_t *grid, int64_t *frep, float *w, float *wg,
                  float *avg_reward, int64_t *iter, orc_event *next_event, int64_t stats8[8]) {
  if (grid) memcpy(grid, s->grid, sizeof s->grid);
  if (frep) memcpy(frep, s->frep, sizeof s->frep);
  if (w) memcpy(w, s->w, sizeof s->w);
  if (wg) memcpy(wg, s->wg, sizeof s->wg);
  if (avg_reward) *avg_reward = s->avg_reward;
  if (iter) *iter = s->iter;
  if (next_event) *next_event = s->event;
  if (stats8) memcpy(stats8, s->stats, sizeof s->stats);
}

void orc_sim_period(orc_sim *s, float *loss_sum, int64_t *loss_n, int reset) {
  if (loss_sum) *loss_sum = s->loss_sum;
  if (loss_n) *loss_n = s->loss_n;
  if (reset) { /* statsReportPeriod resets the period counters, Stats.hs:117-118 */
    s->loss_sum = 0.0f;
    s->loss_n = 0;
    s->stats[6] = 0;
    s->stats[7] = 0;
  }
}
uint64_t orc_sim_ndraws(const orc_sim *s) { return orc_eg_ndraws(s->eg); }

void orc_stats_cumus(const int64_t st[8], double out[3]) { /* Stats.hs:83-92 */
  double rejNew = (double)st[2], arrNew = (double)st[0];
  double rejHoff = (double)st[5], arrHoff = (double)st[3];
  out[0] = rejNew / (arrNew + 1.0);
  out[1] = rejHoff / (arrHoff + 1.0);
  out[2] = (rejNew + rejHoff) / (arrNew + arrHoff + 1.0);
}

/* ------------------------------------------------------------------------ */
/* Batched CPU baseline                                                     */
/* ------------------------------------------------------------------------ */
int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int64_t orc_run_replicas(const orc_opt *base, int64_t first_replica, int n_replicas,
                         int64_t n_events, int n_threads, int64_t stats_sum8[8],
                         float *avg_reward_out, double *bp_out) {
  int64_t total = 0;
  int64_t sums[8] = {0};
  build_tables();
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
#endif
  for (int i = 0; i < n_replicas; ++i) {
    orc_opt o = *base;
    o.replica = (uint64_t)(first_replica + i);
    o.n_events = n_events;
    orc_sim *s = orc_sim_create(&o);
    int64_t done = 0;
    orc_sim_run(s, n_events, NULL, 0, &done);
    total += done;
    if (avg_reward_out) avg_reward_out[i] = s->avg_reward;
    if (bp_out) {
      double cum[3];
      orc_stats_cumus(s->stats, cum);
      bp_out[i] = cum[0];
    }
#ifdef _OPENMP
#pragma omp critical(orc_sum)
#endif
    for (int k = 0; k < 8; ++k) sums[k] += s->stats[k];
    orc_sim_destroy(s);
  }
  if (stats_sum8) memcpy(stats_sum8, sums, sizeof sums);
  return total;
}
