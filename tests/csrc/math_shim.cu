// TEST-ONLY host instantiation of optuna_b200/csrc/tpe_math.cuh (the product never links this).
// Lets `pytest -m "not gpu"` check the special-function logic against the oracle without a GPU.
#include "../../optuna_b200/csrc/tpe_math.cuh"
#include "../../optuna_b200/csrc/tpe_motpe.cuh"
#include <vector>

extern "C" {
#define MAP1(name, fn) void name(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = fn(x[i]); }
MAP1(shim_erf_np, tpe::erf_np)
MAP1(shim_erf_c, tpe::erf_c)
MAP1(shim_erfc_c, tpe::erfc_c)
MAP1(shim_ndtr_vec, tpe::ndtr_vec)
MAP1(shim_ndtr_single, tpe::ndtr_single)
MAP1(shim_log_ndtr, tpe::log_ndtr)
MAP1(shim_ndtri_exp, tpe::ndtri_exp)
void shim_log_gauss_mass(const double* a, const double* b, double* y, long n) {
  for (long i = 0; i < n; ++i) y[i] = tpe::log_gauss_mass(a[i], b[i]);
}
void shim_trunc_ppf(const double* q, const double* a, const double* b, double* y, long n) {
  for (long i = 0; i < n; ++i) y[i] = tpe::trunc_ppf(q[i], a[i], b[i]);
}
double shim_hypervolume(const double* v, int n, int m, const double* ref, int assume_pareto) {
  std::vector<double> arena(tpe::hv_arena_doubles(n, m) + 16);
  memset(arena.data(), 0xFF, arena.size() * 8);  // the device arena is not zeroed either
  return tpe::hypervolume(v, n, m, ref, assume_pareto != 0, arena.data());
}
double shim_pairwise(const double* x, long n) {
  return tpe::np_pairwise_sum(x, (int)n);
}
}
