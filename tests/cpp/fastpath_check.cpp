// fastpath_check.cpp — CPU check of the fast edge kernel's arithmetic (batching_pomp_b200/csrc/fastpath.cuh,
// fastpath_host.hpp): builds the very tables the device uses and replays, edge by edge, what the kernel's
// phases compute, against a brute-force evaluation with the reference's FP64 arithmetic
// (/root/reference/src/BatchingPOMP.cpp:435-461, 496-544).  Checked per edge the fast path accepts:
//   1. broad phase: every box that contains ANY sample of the edge is in the AND of the mask rows;
//   2. OUTER range: every sample truly inside a candidate box has its coordinate s = k+1 in [smin, smax];
//   3. INNER range: every sample the kernel would accept without the exact test is truly inside;
//   4. the first invalid slot the kernel's phase C arrives at equals the brute-force one.
// Inputs: config-4-like random edges, axis-aligned edges, and edges constructed so that samples land on or
// within an ulp of box faces.  Exit code 0 = all checks passed.  Test infrastructure only.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <random>
#include <vector>

#include "fastpath_host.hpp"

static std::vector<int> bisect_perm(int n) {  // util/BisectPerm.hpp:45-89 as a gap heap (same order)
  struct Gap {
    long val, idx, a, b;
    bool operator<(const Gap& o) const { return val != o.val ? val < o.val : idx > o.idx; }
  };
  std::priority_queue<Gap> pq;
  std::vector<int> out;
  auto push = [&](long a, long b) {
    long v = (b - a) / 2;
    if (v >= 1) pq.push(Gap{v, a + v, a, b});
  };
  push(-1, n);
  while (!pq.empty()) {
    Gap g = pq.top();
    pq.pop();
    out.push_back((int)g.idx);
    push(g.a, g.idx);
    push(g.idx, g.b);
  }
  return out;
}

struct World {
  int D, nb;
  std::vector<double> lo, hi;
  FpTables t;
};

static bool inside_exact(const World& w, int b, const double* x) {
  for (int j = 0; j < w.D; ++j)
    if (!(w.lo[(size_t)b * w.D + j] <= x[j] && x[j] <= w.hi[(size_t)b * w.D + j])) return false;
  return true;
}

static double g_thr[FP_NTHR + 2u];
static void build_thr(double L) {  // as ctx.cu builds the table
  auto nst = [L](double s) { return std::floor(std::sqrt(s) / L); };
  for (uint32_t k = 0; k < FP_NTHR + 2u; ++k) {
    g_thr[k] = 0.0;
    if (k < 3) continue;
    double hi = ((double)k + 2.0) * L;
    hi *= hi;
    uint64_t a = 0, b;
    memcpy(&b, &hi, 8);
    while (b - a > 1) {
      const uint64_t m = a + (b - a) / 2;
      double sm;
      memcpy(&sm, &m, 8);
      if (nst(sm) >= (double)k) b = m;
      else a = m;
    }
    memcpy(&g_thr[k], &b, 8);
  }
}

struct Stats {
  long edges = 0, fast = 0, slow = 0, pairs = 0, hits = 0, exact_tests = 0, blocked = 0, fails = 0;
};

static void fail(Stats& st, const char* what, long e, int b, int k) {
  if (st.fails < 20) fprintf(stderr, "FAIL %s: edge %ld box %d k %d\n", what, e, b, k);
  st.fails++;
}

static void check_edge(const World& w, const double* f, const double* to, double L, long eidx, Stats& st,
                       std::vector<std::vector<int>>& perms) {
  const int D = w.D;
  st.edges++;
  volatile double dist2 = 0.0;
  double d[8];
  for (int j = 0; j < D; ++j) {
    volatile double df = f[j] - to[j];
    volatile double sq = df * df;
    dist2 = dist2 + sq;
    volatile double dd = to[j] - f[j];
    d[j] = dd;
  }
  double q = std::floor(std::sqrt((double)dist2) / L);
  // nStates through the exact thresholds on the squared distance must equal the FP64 formula
  {
    const uint32_t nt = fp_nstates_thr((double)dist2, (float)(1.0 / L), g_thr);
    const uint32_t nref = !(q >= 2.0) ? 2u : (q > 65535.0 ? 0u : (uint32_t)q);
    if (nt != 0u && nt != nref) fail(st, "threshold nStates differs", eidx, (int)nt, (int)nref);
    if (nt == 0u && nref != 0u && nref < FP_NTHR - 4u) fail(st, "threshold nStates gave up early", eidx, 0, (int)nref);
  }
  if (!(q >= 2.0)) return;
  if (q > 65535.0) return;
  const uint32_t n = (uint32_t)q;
  if (n <= 2) return;
  double F[8], T[8];
  bool safe = true;
  int spanmax = 0;
  int row[8];
  for (int j = 0; j < D; ++j) {
    F[j] = fp_q(f[j], w.t.qs[j], w.t.qo[j]);
    T[j] = fp_q(to[j], w.t.qs[j], w.t.qo[j]);
    safe = safe && F[j] >= FP_QOFF - 1.0 && F[j] <= FP_QOFF + FP_QSCALE + 1.0 && T[j] >= FP_QOFF - 1.0 &&
           T[j] <= FP_QOFF + FP_QSCALE + 1.0;
    const int cf = fp_cell(F[j]), ct = fp_cell(T[j]);
    const int c0 = std::min(cf, ct), span = std::max(cf, ct) - c0;
    spanmax = std::max(spanmax, span);
    row[j] = std::min(span, FP_CLASSES - 1) * FP_CELLS + c0;
  }
  if (!safe || n > FP_NMAX || spanmax > FP_SPAN_MAX) {
    st.slow++;
    return;
  }
  st.fast++;
  if ((int)perms.size() <= (int)n) perms.resize(n + 1);
  if (perms[n].empty()) perms[n] = bisect_perm((int)n);
  const std::vector<int>& perm = perms[n];
  // samples in k order with the reference's arithmetic
  std::vector<double> xs((size_t)n * D);
  for (uint32_t k = 0; k < n; ++k) {
    const double tt = 1.0 * (1 + (int)k) / (n + 1);
    for (int j = 0; j < D; ++j) {
      volatile double m = d[j] * tt;
      volatile double x = f[j] + m;
      xs[(size_t)k * D + j] = x;
    }
  }
  // brute-force first invalid slot
  int first_exact = -1;
  for (uint32_t i = 1; i + 1 < n && first_exact < 0; ++i)
    for (int b = 0; b < w.nb; ++b)
      if (inside_exact(w, b, &xs[(size_t)perm[i] * D])) {
        first_exact = (int)i;
        break;
      }
  if (first_exact >= 0) st.blocked++;
  // 1. broad phase: AND of the mask rows as the kernel addresses them
  std::vector<uint32_t> cand(16, ~0u);  // 4 groups x 4 words
  for (int j = 0; j < D; ++j)
    for (int g = 0; g < 4; ++g)
      for (int wd = 0; wd < 4; ++wd)
        cand[g * 4 + wd] &= w.t.tmask[fp_row_offset(D, j, (uint32_t)row[j]) / 4 + (size_t)g * 4 + wd];
  for (int b = 0; b < w.nb; ++b) {
    bool any = false;
    for (uint32_t k = 0; k < n && !any; ++k) any = inside_exact(w, b, &xs[(size_t)k * D]);
    const bool in_cand = (cand[(b >> 7) * 4 + ((b >> 5) & 3)] >> (b & 31)) & 1u;
    if (any && !in_cand) fail(st, "box with an inside sample is not a candidate", eidx, b, -1);
  }
  // per-edge slab data
  float inv[8], cc[8];
  bool has_static = false;
  for (int j = 0; j < D; ++j) {
    FpDim fd = fp_edge_dim(F[j], T[j]);
    inv[j] = fd.inv;
    cc[j] = fd.c;
    has_static = has_static || fd.stat;
  }
  // phase C replay
  int first_fast = 0x7fffffff;
  for (int b = 0; b < w.nb; ++b) {
    if (!((cand[(b >> 7) * 4 + ((b >> 5) & 3)] >> (b & 31)) & 1u)) continue;
    st.pairs++;
    const uint16_t* qb = &w.t.qbox[(size_t)b * 16];
    float smin = -3.0e38f, smax = 3.0e38f, imin = smin, imax = smax;
    for (int j = 0; j < D; ++j) {
      const uint32_t wl = (uint32_t)qb[j & ~1] | ((uint32_t)qb[(j & ~1) + 1] << 16);
      const uint32_t wh = (uint32_t)qb[8 + (j & ~1)] | ((uint32_t)qb[8 + (j & ~1) + 1] << 16);
      const float lo_m = (j & 1) ? fp_magic<1>(wl) : fp_magic<0>(wl);
      const float hi_m = (j & 1) ? fp_magic<1>(wh) : fp_magic<0>(wh);
      fp_slab(lo_m, hi_m, inv[j], cc[j], smin, smax, imin, imax);
    }
    fp_range_to_s(n, smin, smax, imin, imax);
    // 2. / 3. against every sample
    for (uint32_t k = 0; k < n; ++k) {
      const float s = (float)(k + 1);
      const bool truly = inside_exact(w, b, &xs[(size_t)k * D]);
      const bool outer = s >= smin && s <= smax;
      const bool inner = !has_static && s >= imin && s <= imax;
      if (truly && !outer) fail(st, "inside sample outside the OUTER range", eidx, b, (int)k);
      if (inner && !truly) fail(st, "sample in the INNER range is not inside", eidx, b, (int)k);
      if (inner && !outer) fail(st, "INNER range not within OUTER range", eidx, b, (int)k);
    }
    // the kernel's slot search
    if (smax >= smin && std::floor(smax) >= smin) {
      const int lim = std::min((int)n - 1, first_fast);
      for (int i = 1; i < lim; ++i) {
        const float s = (float)(1 + perm[i]);
        if (s >= smin && s <= smax) {
          bool ins = !has_static && s >= imin && s <= imax;
          if (!ins) {
            st.exact_tests++;
            ins = inside_exact(w, b, &xs[(size_t)perm[i] * D]);
          }
          if (ins) {
            first_fast = std::min(first_fast, i);
            st.hits++;
            break;
          }
        }
      }
    }
  }
  const int ff = first_fast == 0x7fffffff ? -1 : first_fast;
  if (ff != first_exact) fail(st, "first invalid slot differs", eidx, ff, first_exact);
}

static World make_world(int D, int nb, unsigned seed, double cmin, double cmax, double hmin, double hmax,
                        bool bounds, double shift) {
  World w;
  w.D = D;
  w.nb = nb;
  std::mt19937_64 rng(seed);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  w.lo.resize((size_t)nb * D);
  w.hi.resize((size_t)nb * D);
  for (int b = 0; b < nb; ++b)
    for (int j = 0; j < D; ++j) {
      const double c = cmin + (cmax - cmin) * U(rng) + shift, h = hmin + (hmax - hmin) * U(rng);
      w.lo[(size_t)b * D + j] = c - h;
      w.hi[(size_t)b * D + j] = c + h;
    }
  double blo[8], bhi[8];
  for (int j = 0; j < 8; ++j) {
    blo[j] = shift;
    bhi[j] = shift + 1.0;
  }
  w.t = fp_build_tables(D, nb, w.lo.data(), w.hi.data(), bounds, blo, bhi);
  return w;
}

int main(int argc, char** argv) {
  const long scale = argc > 1 ? atol(argv[1]) : 1;
  const bool pure = argc > 2 && !strcmp(argv[2], "pure");  // random edges only (statistics of the common case)
  Stats total;
  struct Cfg {
    int D, nb;
    double cmin, cmax, hmin, hmax;
    bool bounds;
    double shift, frac, maxlen;
    long edges;
  };
  const Cfg cfgs[] = {
      {7, 500, 0.0, 1.0, 0.10, 0.25, false, 0.0, 0.01, 0.3474, 60000},   // BASELINE config 4 world
      {7, 500, 0.0, 1.0, 0.10, 0.25, true, 0.0, 0.01, 0.3474, 60000},    // the same with state-space bounds
      {7, 500, 0.0, 1.0, 0.10, 0.25, true, 0.0, 0.002, 0.30, 8000},      // finer resolution: n up to ~56
      {2, 40, 0.0, 1.0, 0.02, 0.08, true, 0.0, 0.002, 0.20, 30000},
      {3, 100, 0.2, 0.8, 0.01, 0.05, false, 0.0, 0.005, 0.40, 30000},
      {5, 300, 0.0, 1.0, 0.05, 0.30, true, 1000.0, 0.01, 0.50, 20000},   // world far from the origin
      {8, 200, 0.0, 1.0, 0.10, 0.30, false, -3.0, 0.01, 0.40, 20000},
      {1, 5, 0.0, 1.0, 0.01, 0.05, true, 0.0, 0.001, 0.30, 5000},
      {6, 512, 0.0, 1.0, 0.02, 0.10, true, 0.0, 0.004, 0.25, 10000},     // table full: 512 boxes
  };
  int ci = 0;
  for (const Cfg& c : cfgs) {
    World w = make_world(c.D, c.nb, 100 + ci, c.cmin, c.cmax, c.hmin, c.hmax, c.bounds, c.shift);
    ++ci;
    if (!w.t.enabled) {
      fprintf(stderr, "config %d: tables not enabled\n", ci);
      return 2;
    }
    const int D = c.D;
    double ext2 = 0.0;
    for (int j = 0; j < D; ++j) ext2 += 1.0;
    const double L = std::sqrt(ext2) * c.frac;
    build_thr(L);
    Stats st;
    std::vector<std::vector<int>> perms;
    std::mt19937_64 rng(7 * ci);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    std::normal_distribution<double> N(0.0, 1.0);
    double f[8], t[8];
    long e = 0;
    const long nrand = c.edges * scale;
    for (long r = 0; r < nrand; ++r, ++e) {  // random edges as workloads.random_edges draws them
      double g[8], nrm = 0.0;
      for (int j = 0; j < D; ++j) {
        f[j] = c.shift + U(rng);
        g[j] = N(rng);
        nrm += g[j] * g[j];
      }
      nrm = std::sqrt(nrm);
      const double ell = U(rng) * c.maxlen;
      for (int j = 0; j < D; ++j)
        t[j] = std::min(c.shift + 1.0, std::max(c.shift, f[j] + ell * g[j] / nrm));
      if (pure) {
        check_edge(w, f, t, L, e, st, perms);
        continue;
      }
      if (r % 16 == 3)  // some edges leave the unit cube / the region
        for (int j = 0; j < D; ++j) t[j] = f[j] + 1.3 * ell * g[j] / nrm;
      if (r % 16 == 5) {  // axis-aligned: every other dimension static
        const int ja = (int)(U(rng) * D) % D;
        for (int j = 0; j < D; ++j)
          if (j != ja) t[j] = f[j];
        t[ja] = f[ja] + (U(rng) < 0.5 ? -1 : 1) * ell;
      }
      if (r % 16 == 7) {  // nearly static dimension: moves by a few 1e-6
        const int ja = (int)(U(rng) * D) % D;
        t[ja] = f[ja] + (U(rng) - 0.5) * 1e-5;
      }
      check_edge(w, f, t, L, e, st, perms);
    }
    // adversarial: a sample lands exactly on (or an ulp beside) a box face
    for (long r = 0; r < (pure ? 0 : nrand / 4); ++r, ++e) {
      const int b = (int)(U(rng) * w.nb) % w.nb, ja = (int)(U(rng) * D) % D;
      for (int j = 0; j < D; ++j) {  // start inside the box in the other dimensions, short move
        const double l = w.lo[(size_t)b * D + j], h = w.hi[(size_t)b * D + j];
        f[j] = l + (h - l) * U(rng);
        t[j] = f[j] + (U(rng) - 0.5) * 0.2 * c.maxlen;
      }
      double len = 0.0;
      for (int j = 0; j < D; ++j) len += (t[j] - f[j]) * (t[j] - f[j]);
      const int n = (int)std::floor(std::sqrt(len) / L);
      const double face = (r & 1) ? w.lo[(size_t)b * D + ja] : w.hi[(size_t)b * D + ja];
      const int mode = (int)(r % 5);
      if (mode == 0) {  // static dimension exactly on the face (closed box: inside)
        f[ja] = t[ja] = face;
      } else if (mode == 1) {  // static dimension one ulp outside
        f[ja] = t[ja] = std::nextafter(face, (r & 1) ? -1e300 : 1e300);
      } else if (n >= 3) {  // sample k on the face: to = f + (face - f)/t_k, then nudged by -1, 0, +1 ulp
        const int k = (int)(U(rng) * n) % n;
        const double tk = 1.0 * (1 + k) / (n + 1);
        double tj = f[ja] + (face - f[ja]) / tk;
        if (mode == 3) tj = std::nextafter(tj, 1e300);
        if (mode == 4) tj = std::nextafter(tj, -1e300);
        t[ja] = tj;
      }
      check_edge(w, f, t, L, e, st, perms);
    }
    printf("config %d (D=%d, %d boxes%s): %ld edges, %ld fast, %ld slow, %ld blocked, %ld pairs, %ld hits, %ld exact tests, "
           "%ld failures\n", ci, c.D, c.nb, c.bounds ? ", bounds" : "", st.edges, st.fast, st.slow, st.blocked, st.pairs,
           st.hits, st.exact_tests, st.fails);
    total.fails += st.fails;
    total.fast += st.fast;
  }
  if (total.fails) {
    fprintf(stderr, "%ld failures\n", total.fails);
    return 1;
  }
  printf("fastpath_check ok (%ld fast-path edges)\n", total.fast);
  return 0;
}
