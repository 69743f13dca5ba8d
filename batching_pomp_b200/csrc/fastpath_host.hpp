// fastpath_host.hpp — host-side builder of the fast kernel's tables (fastpath.cuh): the quantisation q_j over
// the region of interest, the 16-bit box bounds and the broad-phase masks in the team layout.  Plain C++ so
// that the CPU check of the fast path's error analysis (tests/cpp/fastpath_check.cpp) builds the very same
// tables the device uses.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <vector>

#include "fastpath.cuh"

struct FpTables {
  int enabled = 0;
  int nps = 0;
  double qs[8] = {0}, qo[8] = {0};
  std::vector<uint16_t> qbox;   // [nb][16]: lo_q[8], hi_q[8]
  std::vector<uint32_t> tmask;  // fp_row_offset() layout, fp_tmask_bytes(D) bytes
};

// Index (in 32-bit words) of word `word` of group g of the row (class k, first cell c0) of dimension j.
inline size_t fp_tmask_index(int D, int j, int k, int c0, int g, int word) {
  return (size_t)fp_row_offset(D, j, (uint32_t)(k * FP_CELLS + c0)) / 4 + (size_t)g * 4 + (size_t)word;
}

// Region of interest per dimension: the bounding region of the boxes, intersected with the state-space bounds
// when given.  Worlds the fast kernel does not cover (no boxes or more than FP_MAXBOXES, non-finite or
// degenerate extents, coordinates huge relative to the extent) return enabled = 0.
inline FpTables fp_build_tables(int D, int nb, const double* lo, const double* hi, bool have_bounds,
                                const double* blo, const double* bhi) {
  FpTables t;
  t.nps = (D + 1) / 2;
  if (nb < 1 || nb > FP_MAXBOXES || D < 1 || D > 8) return t;
  for (int j = 0; j < D; ++j) {
    double mn = DBL_MAX, mx = -DBL_MAX;
    for (int b = 0; b < nb; ++b) {
      const double l = lo[(size_t)b * D + j], h = hi[(size_t)b * D + j];
      if (!std::isfinite(l) || !std::isfinite(h)) return t;
      if (l < mn) mn = l;
      if (h > mx) mx = h;
    }
    if (have_bounds && blo[j] < bhi[j]) {
      const double a = std::max(mn, blo[j]), b2 = std::min(mx, bhi[j]);
      if (a < b2) {
        mn = a;
        mx = b2;
      }
    }
    const double ext = mx - mn;
    if (!(ext > 0.0) || !std::isfinite(ext)) return t;
    // the FP64 rounding of a sample, 2^-52 |x| qs, must stay far below the 0.35 q-units of margin
    if ((std::fabs(mn) + std::fabs(mx)) / ext > 1e6) return t;
    t.qs[j] = FP_QSCALE / ext;
    t.qo[j] = FP_QOFF - mn * t.qs[j];
    if (!std::isfinite(t.qs[j]) || !std::isfinite(t.qo[j])) return t;
  }
  t.qbox.assign((size_t)nb * 16, 0);
  t.tmask.assign(fp_tmask_bytes(D) / 4, 0u);
  for (int b = 0; b < nb; ++b) {
    bool empty = false;
    for (int j = 0; j < D; ++j)
      if (!(lo[(size_t)b * D + j] <= hi[(size_t)b * D + j])) empty = true;
    for (int j = 0; j < 8; ++j) {
      double ql = 0.0, qh = 65535.0;
      if (j < D) {  // widened by FP_QSLACK outwards; bounds beyond the region clamp
        ql = std::floor(fp_q(lo[(size_t)b * D + j], t.qs[j], t.qo[j])) - FP_QSLACK;
        qh = std::ceil(fp_q(hi[(size_t)b * D + j], t.qs[j], t.qo[j])) + FP_QSLACK;
      }
      t.qbox[(size_t)b * 16 + j] = (uint16_t)std::min(65535.0, std::max(0.0, ql));
      t.qbox[(size_t)b * 16 + 8 + j] = (uint16_t)std::min(65535.0, std::max(0.0, qh));
    }
    if (empty) continue;  // contains nothing: never a candidate
    const int g = b >> 7, word = (b >> 5) & 3;
    const uint32_t bit = 1u << (b & 31);
    for (int j = 0; j < D; ++j) {
      const int cl = fp_cell(fp_q(lo[(size_t)b * D + j], t.qs[j], t.qo[j]));
      const int ch = fp_cell(fp_q(hi[(size_t)b * D + j], t.qs[j], t.qo[j]));
      for (int k = 0; k < FP_CLASSES; ++k)
        for (int c0 = 0; c0 < FP_CELLS; ++c0) {
          if (!(cl <= c0 + fp_class_span(k) && ch >= c0)) continue;
          t.tmask[fp_tmask_index(D, j, k, c0, g, word)] |= bit;
        }
    }
  }
  t.enabled = 1;
  return t;
}
