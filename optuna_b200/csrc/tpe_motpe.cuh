// Exact hypervolume arithmetic for MOTPE (fp64), `__host__ __device__` so that the host
// instantiation can be checked against the oracle without a GPU (tests/csrc/math_shim.cu).
//
//   hv_2d / hv_3d / hv_nd / hypervolume   optuna/_hypervolume/wfg.py:8-181
//   front_sorted                          optuna/study/_multi_objective.py:127-168
//   reference_point                       optuna/samplers/_tpe/sampler.py:679-683
//
// All point sets are small here (the below set of TPE, <= n_below points): one thread evaluates
// one hypervolume, working in a caller-provided scratch arena.  Differences from the reference are
// limited to summation order inside its BLAS dot products (2-D / 3-D cases) -- relative 1e-16.
#pragma once
#include "tpe_math.cuh"

namespace tpe {

// lexicographic compare of two M-vectors: <0, 0, >0
TPE_HD int lex_cmp(const double* a, const double* b, int M) {
  for (int j = 0; j < M; ++j) {
    if (a[j] < b[j]) return -1;
    if (a[j] > b[j]) return 1;
  }
  return 0;
}

// Non-dominated mask of a lexsorted (by coordinate 0, ties allowed) [n, M] array, following
// _is_pareto_front_nd / _2d: walk the rows, keep a row, drop every later row that is not strictly
// better in some coordinate 1..M-1.  alive: n bytes of scratch.
TPE_HD void front_sorted(const double* v, int n, int M, uint8_t* mask, uint8_t* alive) {
  if (M == 1) {
    for (int i = 0; i < n; ++i) mask[i] = (i == 0);
    return;
  }
  if (M == 2) {
    double run = 0.0;
    for (int i = 0; i < n; ++i) {
      const double y = v[i * 2 + 1];
      if (i == 0) { mask[i] = 1; run = y; }
      else { const double nr = y < run ? y : run; mask[i] = nr < run; run = nr; }
    }
    return;
  }
  for (int i = 0; i < n; ++i) { alive[i] = 1; mask[i] = 0; }
  for (int h = 0; h < n; ++h) {
    if (!alive[h]) continue;
    mask[h] = 1;
    alive[h] = 0;
    for (int i = h + 1; i < n; ++i) {
      if (!alive[i]) continue;
      bool better = false;
      for (int j = 1; j < M; ++j) better = better || (v[i * M + j] < v[h * M + j]);
      if (!better) alive[i] = 0;
    }
  }
}

// in-place stable insertion sort of rows by full lexicographic order; returns number of unique rows
// after dropping duplicates (np.unique(axis=0)).
TPE_HD int unique_lexsort(double* v, int n, int M, double* tmp) {
  for (int i = 1; i < n; ++i) {
    for (int j = 0; j < M; ++j) tmp[j] = v[i * M + j];
    int p = i;
    while (p > 0 && lex_cmp(tmp, v + (p - 1) * M, M) < 0) {
      for (int j = 0; j < M; ++j) v[p * M + j] = v[(p - 1) * M + j];
      --p;
    }
    for (int j = 0; j < M; ++j) v[p * M + j] = tmp[j];
  }
  int u = 0;
  for (int i = 0; i < n; ++i) {
    if (u > 0 && lex_cmp(v + i * M, v + (u - 1) * M, M) == 0) continue;
    if (u != i)
      for (int j = 0; j < M; ++j) v[u * M + j] = v[i * M + j];
    ++u;
  }
  return u;
}

// stable insertion sort by coordinate 0 only (loss_vals[loss_vals[:, 0].argsort()], wfg.py:170)
TPE_HD void sort_by_first(double* v, int n, int M, double* tmp) {
  for (int i = 1; i < n; ++i) {
    for (int j = 0; j < M; ++j) tmp[j] = v[i * M + j];
    int p = i;
    while (p > 0 && tmp[0] < v[(p - 1) * M]) {
      for (int j = 0; j < M; ++j) v[p * M + j] = v[(p - 1) * M + j];
      --p;
    }
    for (int j = 0; j < M; ++j) v[p * M + j] = tmp[j];
  }
}

TPE_HD double hv_2d(const double* s, int n, const double* ref) {
  double acc = 0.0;
  for (int i = 0; i < n; ++i) {
    const double ex = TPE_SUB(ref[0], s[i * 2]);
    const double ey = TPE_SUB(i == 0 ? ref[1] : s[(i - 1) * 2 + 1], s[i * 2 + 1]);
    acc = TPE_ADD(acc, TPE_MUL(ex, ey));
  }
  return acc;
}

// wfg.py:16-38.  s sorted by x; order: n ints of scratch (ranks of the rows by y, stable).
TPE_HD double hv_3d(const double* s, int n, const double* ref, int* order) {
  for (int i = 0; i < n; ++i) order[i] = i;
  for (int i = 1; i < n; ++i) {  // stable argsort by y
    const int o = order[i];
    int p = i;
    while (p > 0 && s[o * 3 + 1] < s[order[p - 1] * 3 + 1]) { order[p] = order[p - 1]; --p; }
    order[p] = o;
  }
  // z[row = order[j]][col = j] = ref2 - s[order[j]][2]; cumulative max over rows then columns;
  // result = sum_i dx[i] * (sum_j zc[i][j] * dy[j])
  double total = 0.0;
  for (int i = 0; i < n; ++i) {
    const double dx = TPE_SUB(i + 1 < n ? s[(i + 1) * 3] : ref[0], s[i * 3]);
    double run = 0.0, inner = 0.0;
    for (int j = 0; j < n; ++j) {
      if (order[j] <= i) {
        const double z = TPE_SUB(ref[2], s[order[j] * 3 + 2]);
        run = z > run ? z : run;
      }
      const double yj = s[order[j] * 3 + 1];
      const double dy = TPE_SUB(j + 1 < n ? s[order[j + 1] * 3 + 1] : ref[1], yj);
      inner = TPE_ADD(inner, TPE_MUL(run, dy));
    }
    total = TPE_ADD(total, TPE_MUL(inner, dx));
  }
  return total;
}

// wfg.py:41-77.  s: [n, M] (lexsorted by the caller's construction).  The reference recursion
//   HV(S) = incl(last) + sum_i [ incl(i) - HV(front(limit(S_{>i}, i))) ]
// is run with an explicit frame stack in the scratch arena (no device call stack needed).
struct HvFrame {
  const double* s;
  double* lim;     // (n - 1) * M doubles + flag bytes for this level
  double sum, incl;
  int n, i;
};
TPE_HD double hv_nd(const double* s0, int n0, int M, const double* ref, double* arena) {
  HvFrame* fr = reinterpret_cast<HvFrame*>(arena);
  double* data = arena + ((size_t)(n0 + 2) * sizeof(HvFrame) + 7) / 8;
  int top = 0;
  fr[0].s = s0; fr[0].n = n0; fr[0].i = 0; fr[0].sum = 0.0; fr[0].incl = 0.0; fr[0].lim = data;
  double ret = 0.0;
  bool have_ret = false;
  while (top >= 0) {
    HvFrame& f = fr[top];
    const double* s = f.s;
    const int n = f.n;
    if (n == 1) {
      double out = 1.0;
      for (int j = 0; j < M; ++j) out = TPE_MUL(out, TPE_SUB(ref[j], s[j]));
      ret = out; have_ret = true; --top;
      continue;
    }
    if (n == 2) {
      double h1 = 1.0, h2 = 1.0, cap = 1.0;
      for (int j = 0; j < M; ++j) {
        const double a = s[j], b = s[M + j];
        h1 = TPE_MUL(h1, TPE_SUB(ref[j], a));
        h2 = TPE_MUL(h2, TPE_SUB(ref[j], b));
        cap = TPE_MUL(cap, TPE_SUB(ref[j], a > b ? a : b));
      }
      ret = TPE_SUB(TPE_ADD(h1, h2), cap); have_ret = true; --top;
      continue;
    }
    if (have_ret) {  // back from the child of index f.i
      f.sum = TPE_ADD(f.sum, TPE_SUB(f.incl, ret));
      ++f.i;
      have_ret = false;
    }
    if (f.i == n - 1) {
      double last = 1.0;
      for (int j = 0; j < M; ++j) last = TPE_MUL(last, TPE_SUB(ref[j], s[(n - 1) * M + j]));
      ret = TPE_ADD(last, f.sum); have_ret = true; --top;
      continue;
    }
    const int i = f.i;
    double incl = 1.0;
    for (int j = 0; j < M; ++j) incl = TPE_MUL(incl, TPE_SUB(ref[j], s[i * M + j]));
    f.incl = incl;
    double* lim = f.lim;
    int cnt = n - 1 - i;
    for (int r = 0; r < cnt; ++r)
      for (int j = 0; j < M; ++j) {
        const double a = s[i * M + j], b = s[(i + 1 + r) * M + j];
        lim[r * M + j] = a > b ? a : b;
      }
    uint8_t* flags = reinterpret_cast<uint8_t*>(lim + (size_t)(n - 1) * M);
    if (cnt > 3) {
      uint8_t* mask = flags;
      uint8_t* alive = flags + n;
      front_sorted(lim, cnt, M, mask, alive);
      int w = 0;
      for (int r = 0; r < cnt; ++r) {
        if (!mask[r]) continue;
        if (w != r)
          for (int j = 0; j < M; ++j) lim[w * M + j] = lim[r * M + j];
        ++w;
      }
      cnt = w;
    }
    HvFrame& c = fr[top + 1];
    c.s = lim; c.n = cnt; c.i = 0; c.sum = 0.0; c.incl = 0.0;
    c.lim = lim + (size_t)(n - 1) * M + (2 * n + 7) / 8 + 1;
    ++top;
  }
  return ret;
}

// arena doubles needed by hypervolume() for n points in M dims
TPE_HD size_t hv_arena_doubles(int n, int M) {
  return (size_t)n * M + (size_t)(n + 1) * (n + 2) / 2 * M + (size_t)(n + 2) * ((2 * n + 7) / 8 + 2) + 4 * n + 64 +
         ((size_t)(n + 2) * sizeof(HvFrame) + 7) / 8;
}

// compute_hypervolume (wfg.py:110-181).  v is copied into the arena (the input is not modified).
// Returns +inf when the reference point is not finite or the result is not finite.
#if defined(__CUDACC__)
__host__ __device__
#endif
inline double hypervolume(const double* v, int n, int M, const double* ref, bool assume_pareto, double* arena) {
  for (int j = 0; j < M; ++j)
    if (!isfinite(ref[j])) return INFINITY;
  if (n == 0) return 0.0;
  double* s = arena;
  double* tmp = s + (size_t)n * M;
  for (int i = 0; i < n * M; ++i) s[i] = v[i];
  int m = n;
  if (!assume_pareto) {
    m = unique_lexsort(s, n, M, tmp);
    uint8_t* mask = reinterpret_cast<uint8_t*>(tmp + M);
    uint8_t* alive = mask + n;
    front_sorted(s, m, M, mask, alive);
    int w = 0;
    for (int r = 0; r < m; ++r) {
      if (!mask[r]) continue;
      if (w != r)
        for (int j = 0; j < M; ++j) s[w * M + j] = s[r * M + j];
      ++w;
    }
    m = w;
  } else {
    sort_by_first(s, n, M, tmp);
  }
  double hv;
  if (M == 2) hv = hv_2d(s, m, ref);
  else if (M == 3) hv = hv_3d(s, m, ref, reinterpret_cast<int*>(tmp + M));
  else hv = hv_nd(s, m, M, ref, tmp + M + (size_t)(2 * n + 7) / 8 + 2);
  return isfinite(hv) ? hv : INFINITY;
}

}  // namespace tpe
