// Truncated-normal special functions for the TPE kernels (fp64).
//
// Every function here is `__host__ __device__` so that tests/csrc/math_shim.cu can instantiate
// the host side and check the *logic* against the oracle on a machine without a GPU.  The
// product (libtpe_b200.so) only ever calls them from device code.
//
// What is evaluated (reference file:line for each):
//   erf_np        optuna/samplers/_tpe/_erf.py:112-142      msun rational forms as NumPy evaluates them
//   erf_c/erfc_c  FreeBSD msun s_erf.c (what libm math.erf/erfc run; _truncnorm.py:59-70 calls them)
//   ndtr_vec      _truncnorm.py:73-75
//   ndtr_single   _truncnorm.py:59-70
//   log_ndtr      _truncnorm.py:79-102
//   log_gauss_mass _truncnorm.py:113-149
//   ndtri_exp     _truncnorm.py:152-221 (per-element Newton stop instead of batch-global)
//   trunc_ppf     _truncnorm.py:224-266
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define TPE_HD __host__ __device__ __forceinline__
#else
#define TPE_HD inline
#endif

namespace tpe {

// Un-fused multiply / add: NumPy and libm round after every operation; nvcc would contract
// a*b+c into an FMA.  Keeping the reference's rounding sequence makes the rational
// approximations bit-identical to it (only exp/log differ by their own <=1-2 ulp).
#if defined(__CUDA_ARCH__)
#define TPE_MUL(a, b) __dmul_rn((a), (b))
#define TPE_ADD(a, b) __dadd_rn((a), (b))
#define TPE_SUB(a, b) __dsub_rn((a), (b))
#define TPE_DIV(a, b) __ddiv_rn((a), (b))
#else
#define TPE_MUL(a, b) ((a) * (b))
#define TPE_ADD(a, b) ((a) + (b))
#define TPE_SUB(a, b) ((a) - (b))
#define TPE_DIV(a, b) ((a) / (b))
#endif

constexpr double kSqrt2 = 1.4142135623730951;        // 2**0.5
constexpr double kInvSqrt2 = 0.7071067811865475;     // 1 / 2**0.5 (as Python rounds it)
constexpr double kLogSqrt2Pi = 0.9189385332046727;   // math.log(math.sqrt(2*math.pi))
constexpr double kLogisticC = 0.5513288954217921;    // math.sqrt(3) / math.pi
constexpr double kDblEps = 2.220446049250313e-16;

// ---- msun s_erf.c coefficient tables (lowest order first) -------------------------------------
constexpr double kErx = 8.45062911510467529297e-01;
constexpr double kEfx = 1.28379167095512586316e-01;
constexpr double kEfx8 = 1.02703333676410069053e+00;
#define TPE_PP {1.28379167095512558561e-01, -3.25042107247001499370e-01, -2.84817495755985104766e-02, \
                -5.77027029648944159157e-03, -2.37630166566501626084e-05}
#define TPE_QQ {1.0, 3.97917223959155352819e-01, 6.50222499887672944485e-02, 5.08130628187576562776e-03, \
                1.32494738004321644526e-04, -3.96022827877536812320e-06}
#define TPE_PA {-2.36211856075265944077e-03, 4.14856118683748331666e-01, -3.72207876035701323847e-01, \
                3.18346619901161753674e-01, -1.10894694282396677476e-01, 3.54783043256182359371e-02, \
                -2.16637559486879084300e-03}
#define TPE_QA {1.0, 1.06420880400844228286e-01, 5.40397917702171048937e-01, 7.18286544141962662868e-02, \
                1.26171219808761642112e-01, 1.36370839120290507362e-02, 1.19844998467991074170e-02}
#define TPE_RA {-9.86494403484714822705e-03, -6.93858572707181764372e-01, -1.05586262253232909814e+01, \
                -6.23753324503260060396e+01, -1.62396669462573470355e+02, -1.84605092906711035994e+02, \
                -8.12874355063065934246e+01, -9.81432934416914548592e+00}
#define TPE_SA {1.0, 1.96512716674392571292e+01, 1.37657754143519042600e+02, 4.34565877475229228821e+02, \
                6.45387271733267880336e+02, 4.29008140027567833386e+02, 1.08635005541779435134e+02, \
                6.57024977031928170135e+00, -6.04244152148580987438e-02}
#define TPE_RB {-9.86494292470009928597e-03, -7.99283237680523006574e-01, -1.77579549177547519889e+01, \
                -1.60636384855821916062e+02, -6.37566443368389627722e+02, -1.02509513161107724954e+03, \
                -4.83519191608651397019e+02}
#define TPE_SB {1.0, 3.03380607434824582924e+01, 3.25792512996573918826e+02, 1.53672958608443695994e+03, \
                3.19985821950859553908e+03, 2.55305040643316442583e+03, 4.74528541206955367215e+02, \
                -2.24409524465858183362e+01}

template <int N>
TPE_HD double horner(const double (&c)[N], double t) {
  double acc = c[N - 1];
#pragma unroll
  for (int i = N - 2; i >= 0; --i) acc = TPE_ADD(c[i], TPE_MUL(acc, t));
  return acc;
}

TPE_HD double rat_pp_qq(double z) {
  const double pp[] = TPE_PP; const double qq[] = TPE_QQ;
  return TPE_DIV(horner(pp, z), horner(qq, z));
}
TPE_HD double rat_pa_qa(double s) {
  const double pa[] = TPE_PA; const double qa[] = TPE_QA;
  return TPE_DIV(horner(pa, s), horner(qa, s));
}
TPE_HD double rat_ra_sa(double s) {
  const double ra[] = TPE_RA; const double sa[] = TPE_SA;
  return TPE_DIV(horner(ra, s), horner(sa, s));
}
TPE_HD double rat_rb_sb(double s) {
  const double rb[] = TPE_RB; const double sb[] = TPE_SB;
  return TPE_DIV(horner(rb, s), horner(sb, s));
}

// erf as the reference's NumPy path evaluates it (arrays >= 2000 elements; _erf.py:112-142).
TPE_HD double erf_np(double x) {
  if (x != x) return x;
  const double v = fabs(x);
  double r;
  if (v >= 6.0) {
    r = 1.0;
  } else if (v < 3.725290298461914e-09) {  // 2**-28
    r = TPE_MUL(TPE_ADD(1.0, kEfx), v);
  } else if (v < 0.84375) {
    const double z = TPE_MUL(v, v);
    r = TPE_MUL(v, TPE_ADD(1.0, rat_pp_qq(z)));
  } else if (v < 1.25) {
    r = TPE_ADD(kErx, rat_pa_qa(TPE_SUB(v, 1.0)));
  } else {
    const double z = TPE_MUL(v, v);
    const double s = TPE_DIV(1.0, z);
    const double q = (v < 2.857142857142857) ? rat_ra_sa(s) : rat_rb_sb(s);
    // 1 - exp(-z - 0.5625 + R/S) / x
    r = TPE_SUB(1.0, TPE_DIV(exp(TPE_ADD(TPE_SUB(-z, 0.5625), q)), v));
  }
  // np.sign(x) * r
  return (x > 0.0) ? r : ((x < 0.0) ? -r : TPE_MUL(x, r));
}

TPE_HD double clear_low_word(double x) {
  uint64_t u;
  memcpy(&u, &x, 8);
  u &= 0xFFFFFFFF00000000ull;
  double r;
  memcpy(&r, &u, 8);
  return r;
}

// msun erf / erfc in the C library's operation order (math.erf / math.erfc of the reference).
TPE_HD double erf_c(double x) {
  if (x != x) return x;
  const double v = fabs(x);
  if (isinf(x)) return x > 0 ? 1.0 : -1.0;
  if (v < 0.84375) {
    if (v < 3.725290298461914e-09) {
      if (v < 2.848094538889218e-306) return TPE_MUL(0.125, TPE_ADD(TPE_MUL(8.0, x), TPE_MUL(kEfx8, x)));
      return TPE_ADD(x, TPE_MUL(kEfx, x));
    }
    const double y = rat_pp_qq(TPE_MUL(x, x));
    return TPE_ADD(x, TPE_MUL(x, y));
  }
  if (v < 1.25) {
    const double pq = rat_pa_qa(TPE_SUB(v, 1.0));
    return x >= 0 ? TPE_ADD(kErx, pq) : TPE_SUB(-kErx, pq);
  }
  if (v >= 6.0) return x >= 0 ? 1.0 : -1.0;
  const double s = TPE_DIV(1.0, TPE_MUL(v, v));
  const double q = (v < 2.857142857142857) ? rat_ra_sa(s) : rat_rb_sb(s);
  const double z = clear_low_word(v);
  const double r = TPE_MUL(exp(TPE_SUB(TPE_MUL(-z, z), 0.5625)),
                           exp(TPE_ADD(TPE_MUL(TPE_SUB(z, v), TPE_ADD(z, v)), q)));
  return x >= 0 ? TPE_SUB(1.0, TPE_DIV(r, v)) : TPE_SUB(TPE_DIV(r, v), 1.0);
}

TPE_HD double erfc_c(double x) {
  if (x != x) return x;
  if (isinf(x)) return x > 0 ? 0.0 : 2.0;
  const double v = fabs(x);
  if (v < 0.84375) {
    if (v < 1.3877787807814457e-17) return TPE_SUB(1.0, x);  // 2**-56
    const double y = rat_pp_qq(TPE_MUL(x, x));
    if (x < 0.25) return TPE_SUB(1.0, TPE_ADD(x, TPE_MUL(x, y)));
    double r = TPE_MUL(x, y);
    r = TPE_ADD(r, TPE_SUB(x, 0.5));
    return TPE_SUB(0.5, r);
  }
  if (v < 1.25) {
    const double pq = rat_pa_qa(TPE_SUB(v, 1.0));
    if (x >= 0) return TPE_SUB(TPE_SUB(1.0, kErx), pq);
    return TPE_ADD(1.0, TPE_ADD(kErx, pq));
  }
  if (v < 28.0) {
    const double s = TPE_DIV(1.0, TPE_MUL(v, v));
    double q;
    if (v < 2.857142857142857) {
      q = rat_ra_sa(s);
    } else {
      if (x < 0 && v >= 6.0) return 2.0;
      q = rat_rb_sb(s);
    }
    const double z = clear_low_word(v);
    const double r = TPE_MUL(exp(TPE_SUB(TPE_MUL(-z, z), 0.5625)),
                             exp(TPE_ADD(TPE_MUL(TPE_SUB(z, v), TPE_ADD(z, v)), q)));
    return x > 0 ? TPE_DIV(r, v) : TPE_SUB(2.0, TPE_DIV(r, v));
  }
  return x > 0 ? 0.0 : 2.0;
}

TPE_HD double ndtr_vec(double t) {  // 0.5 + 0.5 * erf(t / 2**0.5)
  return TPE_ADD(0.5, TPE_MUL(0.5, erf_np(TPE_DIV(t, kSqrt2))));
}

TPE_HD double ndtr_single(double t) {
  const double u = TPE_DIV(t, kSqrt2);
  if (u < -kInvSqrt2) return TPE_MUL(0.5, erfc_c(-u));
  if (u < kInvSqrt2) return TPE_ADD(0.5, TPE_MUL(0.5, erf_c(u)));
  return TPE_SUB(1.0, TPE_MUL(0.5, erfc_c(u)));
}

TPE_HD double log_ndtr(double t) {
  if (t > 6.0) return -ndtr_single(-t);
  if (t > -20.0) return log(ndtr_single(t));
  // asymptotic series: -t^2/2 - ln(-t) - ln(2 pi)/2 + ln(sum_i (-1)^i (2i-1)!! / t^(2i))
  const double head = TPE_SUB(TPE_SUB(TPE_MUL(-0.5, TPE_MUL(t, t)), log(-t)), kLogSqrt2Pi);
  double prev = 0.0, total = 1.0, num = 1.0, den = 1.0;
  const double inv_t2 = TPE_DIV(1.0, TPE_MUL(t, t));
  double sgn = 1.0;
  int i = 0;
  while (fabs(prev - total) > kDblEps && i < 200) {
    ++i;
    prev = total;
    sgn = -sgn;
    den = TPE_MUL(den, inv_t2);
    num = TPE_MUL(num, (double)(2 * i - 1));
    total = TPE_ADD(total, TPE_MUL(TPE_MUL(sgn, num), den));
  }
  return TPE_ADD(head, log(total));
}

TPE_HD double log_diff(double lp, double lq) {  // lp + log1p(-exp(lq - lp))
  return TPE_ADD(lp, log1p(-exp(TPE_SUB(lq, lp))));
}

// ln(Phi(b) - Phi(a)); NaN inputs give NaN.
TPE_HD double log_gauss_mass(double a, double b) {
  if (b <= 0.0) return log_diff(log_ndtr(b), log_ndtr(a));
  if (a > 0.0) return log_diff(log_ndtr(-a), log_ndtr(-b));
  if (a != a || b != b) return a + b;
  return log1p(TPE_SUB(-ndtr_vec(a), ndtr_vec(-b)));
}

// Same quantity with the library erf() for the central case (used for the (K, P) normalisers of
// the estimator build, where a <= 0 < b and the mass is O(1): |difference| <= ~3e-16 absolute).
TPE_HD double log_gauss_mass_fast(double a, double b) {
  if (b <= 0.0 || a > 0.0 || a != a || b != b) return log_gauss_mass(a, b);
  const double pa = 0.5 + 0.5 * erf(a * 0.7071067811865476);
  const double pb = 0.5 + 0.5 * erf(-b * 0.7071067811865476);
  return log1p(-pa - pb);
}

TPE_HD double logaddexp(double p, double q) {  // numpy.logaddexp
  if (p == q) return p + 0.6931471805599453;
  const double d = p - q;
  if (d > 0) return p + log1p(exp(-d));
  if (d <= 0) return q + log1p(exp(d));
  return p + q;  // NaN
}

// x with log_ndtr(x) == y (y <= 0).  Newton iterations on f(x) = log_ndtr(x) - z.
TPE_HD double ndtri_exp(double y) {
  const bool flip = y > -1e-2;
  const double z = flip ? log(-expm1(y)) : y;
  double x;
  if (z < -5.0) x = -sqrt(TPE_MUL(-2.0, TPE_ADD(z, kLogSqrt2Pi)));
  else x = TPE_MUL(-kLogisticC, log(expm1(-z)));
  for (int it = 0; it < 100; ++it) {
    const double lphi = log_ndtr(x);
    const double lpdf = TPE_SUB(TPE_MUL(-0.5, TPE_MUL(x, x)), kLogSqrt2Pi);
    const double dx = TPE_MUL(TPE_SUB(lphi, z), exp(TPE_SUB(lphi, lpdf)));
    x = TPE_SUB(x, dx);
    if (fabs(dx) < TPE_MUL(1e-8, fabs(x))) break;
    if (!(dx == dx)) break;
  }
  return flip ? -x : x;
}

// Inverse CDF of the standard normal truncated to [a, b] at quantile q.
TPE_HD double trunc_ppf(double q, double a, double b) {
  if (a == b) return NAN;
  if (q == 0.0) return a;
  if (q == 1.0) return b;
  const double lm = log_gauss_mass(a, b);
  if (a < 0.0) return ndtri_exp(logaddexp(log_ndtr(a), TPE_ADD(log(q), lm)));
  return -ndtri_exp(logaddexp(log_ndtr(-b), TPE_ADD(log1p(-q), lm)));
}

// numpy's pairwise summation of n contiguous doubles (numpy/_core/src/umath/loops_utils.h.src).
#if defined(__CUDACC__)
__host__ __device__
#endif
inline double np_pairwise_sum(const double* a, int n) {
  if (n < 8) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r = TPE_ADD(r, a[i]);
    return r;
  }
  if (n <= 128) {
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] = TPE_ADD(r[j], a[i + j]);
    double res = TPE_ADD(TPE_ADD(TPE_ADD(r[0], r[1]), TPE_ADD(r[2], r[3])),
                         TPE_ADD(TPE_ADD(r[4], r[5]), TPE_ADD(r[6], r[7])));
    for (; i < n; ++i) res = TPE_ADD(res, a[i]);
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return TPE_ADD(np_pairwise_sum(a, n2), np_pairwise_sum(a + n2, n - n2));
}

}  // namespace tpe
