// Mixed search spaces with many candidates (multivariate TPE; BASELINE config 3 at n_ei_candidates in the thousands).
//
// A (candidate, kernel) cell sum of a mixed space has two kinds of terms (probability_distributions.py:154-223):
//   continuous columns          -((x - mu) / sigma)^2 / 2, one bandwidth per column (multivariate TPE)
//   tabulated discrete columns  T_j[row(x)][class(k)]      (k_disc_tables: one row per grid value of the candidate)
//   categorical columns         LW_j[class(k)][x]          (k_cat_tables)
// k_logpdf_pairs walks mu / sigma / class of a kernel with a row stride and fetches every table entry from global
// memory.  Here the estimator is laid out for the access pattern of the grid instead:
//   mxc [ceil(n_cont / 2)][kstride] double2   (mu - ctr) / (sigma sqrt 2) of two columns, kernel-minor: a warp of 32
//                                             consecutive kernels loads 512 contiguous bytes
//   mxd [ceil(n_disc+cat / 4)][kstride] ushort4   class indices of four columns, kernel-minor
// and a CTA keeps, for its CB candidates, the scaled coordinates and ALL their table rows in shared memory (config 3:
// 41 KB per candidate), so a cell costs one broadcast shared load + two fp64 instructions per pair of continuous
// columns and one shared-memory gather + one add per discrete / categorical column.  Thread = kernel; the prior
// kernel (its own bandwidths) is left to k_logpdf_prior_fix.  Only for candidates drawn by k_sample (on the grid, inside
// the support); points supplied through tpe_logpdf take k_logpdf_pairs.
#pragma once
#include "tpe_kernels.cuh"

namespace tpe {

struct MixCol {
  int32_t j;      // column (index into cols)
  int32_t kind;   // 0 continuous, 1 tabulated discrete, 2 categorical
  int32_t G;      // table entries per candidate (grid values / choices)
  int32_t off;    // offset (doubles) of the column's slice inside a candidate's block of shared memory
};

// one thread per observation kernel
__global__ void k_mixed_tables(const MixCol* __restrict__ mc, int ncont, int nd, const ColMeta* __restrict__ cols,
                               int32_t pc, const double* __restrict__ mu, const double* __restrict__ sigma,
                               const int32_t* __restrict__ cls, int64_t kobs, int64_t kstride,
                               double2* __restrict__ mxc, ushort4* __restrict__ mxd) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= kstride) return;
  const int ncp = (ncont + 1) >> 1, nd4 = (nd + 3) >> 2;
  for (int p = 0; p < ncp; ++p) {
    double v[2] = {0.0, 0.0};
    for (int h = 0; h < 2; ++h) {
      const int s = 2 * p + h;
      if (s < ncont && k < kobs) {
        const int j = mc[s].j;
        const ColMeta cm = cols[j];
        const double ctr = TPE_MUL(0.5, TPE_ADD(cm.klow, cm.khigh));
        const double inv = 1.0 / (sigma[j] * 1.4142135623730951);   // sigma of observation kernel 0 = of all of them
        v[h] = (mu[k * pc + j] - ctr) * inv;
      }
    }
    mxc[(int64_t)p * kstride + k] = make_double2(v[0], v[1]);
  }
  for (int g = 0; g < nd4; ++g) {
    unsigned short c[4] = {0, 0, 0, 0};
    for (int h = 0; h < 4; ++h) {
      const int q = 4 * g + h;
      if (q < nd && k < kobs) {
        const MixCol m = mc[ncont + q];
        c[h] = (unsigned short)(m.kind == 1 ? cls[k * pc + m.j] : (int)mu[k * pc + m.j]);
      }
    }
    mxd[(int64_t)g * kstride + k] = make_ushort4(c[0], c[1], c[2], c[3]);
  }
}

// grid = (ceil(Ct / CB), k-splits); dynamic shared memory: CB x (2 ncp + tabd + 1) doubles, then 4 nd4 ints
//   part [gridDim.y][ct_stride] (max, sum) per candidate
template <int CB>
__global__ void __launch_bounds__(512, 1)
k_logpdf_mixed(const double* __restrict__ S, int64_t Ct, const ColMeta* __restrict__ cols, int32_t pc,
               const MixCol* __restrict__ mc, int ncont, int nd, int tabd, const double* __restrict__ sigma,
               const double* __restrict__ cst, int64_t kobs, int64_t kstride, int64_t kps,
               const double2* __restrict__ mxc, const ushort4* __restrict__ mxd, const double* __restrict__ tab,
               const double* __restrict__ dtab, const uint8_t* __restrict__ oob, double skip,
               double2* __restrict__ part, int64_t ct_stride) {
  extern __shared__ __align__(16) double s_mix[];
  __shared__ double s_m[16][CB], s_s[16][CB];
  const int ncp = (ncont + 1) >> 1, nd4 = (nd + 3) >> 2;
  const int xw = 2 * ncp, cw = xw + tabd + 1;          // doubles per candidate: coordinates, tables, one zero
  int* s_off = reinterpret_cast<int*>(s_mix + (size_t)CB * cw);
  const int64_t ct0 = (int64_t)blockIdx.x * CB;
  for (int q = threadIdx.x; q < 4 * nd4; q += blockDim.x) s_off[q] = (q < nd) ? mc[ncont + q].off : tabd;
  for (int c = 0; c < CB; ++c) {
    const int64_t ct = ct0 + c;
    double* blk = s_mix + (size_t)c * cw;
    const bool on = ct < Ct;
    for (int s = threadIdx.x; s < xw; s += blockDim.x) {
      double v = 0.0;
      if (on && s < ncont) {
        const int j = mc[s].j;
        const ColMeta cm = cols[j];
        const double x = S[ct * pc + j];
        const double ctr = TPE_MUL(0.5, TPE_ADD(cm.klow, cm.khigh));
        v = ((cm.log ? log(x) : x) - ctr) * (1.0 / (sigma[j] * 1.4142135623730951));
      }
      blk[s] = v;
    }
    if (threadIdx.x == 0) blk[xw + tabd] = 0.0;
    for (int q = 0; q < nd; ++q) {
      const MixCol m = mc[ncont + q];
      double* dst = blk + xw + m.off;
      if (!on) {
        for (int g = threadIdx.x; g < m.G; g += blockDim.x) dst[g] = 0.0;
        continue;
      }
      const ColMeta cm = cols[m.j];
      const double x = S[ct * pc + m.j];
      if (m.kind == 1) {
        const int64_t row = (Ct < cm.grid) ? ct : (int64_t)rint(TPE_DIV(TPE_SUB(x, cm.low), cm.step));
        const double* src = dtab + cm.dtab_off + row * (cm.grid + 1);
        for (int g = threadIdx.x; g < m.G; g += blockDim.x) dst[g] = src[g];
      } else {
        const double* src = tab + cm.tab_off + (int64_t)(cm.nch + 1) * cm.nch + (int)x;
        for (int g = threadIdx.x; g < m.G; g += blockDim.x) dst[g] = src[(int64_t)g * cm.nch];
      }
    }
  }
  __syncthreads();
  double m[CB], s[CB];
#pragma unroll
  for (int c = 0; c < CB; ++c) {
    m[c] = -INFINITY;
    s[c] = 0.0;
  }
  const int64_t k0 = (int64_t)blockIdx.y * kps, k1 = (k0 + kps < kobs) ? k0 + kps : kobs;
  for (int64_t k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
    double q[CB], d[CB];
#pragma unroll
    for (int c = 0; c < CB; ++c) q[c] = d[c] = 0.0;
#pragma unroll 4
    for (int p = 0; p < ncp; ++p) {
      const double2 v = __ldg(mxc + (int64_t)p * kstride + k);
#pragma unroll
      for (int c = 0; c < CB; ++c) {
        const double2 x = *reinterpret_cast<const double2*>(s_mix + (size_t)c * cw + 2 * p);
        const double t0 = x.x - v.x, t1 = x.y - v.y;
        q[c] = fma(t0, t0, q[c]);
        q[c] = fma(t1, t1, q[c]);
      }
    }
#pragma unroll 2
    for (int g = 0; g < nd4; ++g) {
      const ushort4 cl = __ldg(mxd + (int64_t)g * kstride + k);
      const int o0 = s_off[4 * g] + cl.x, o1 = s_off[4 * g + 1] + cl.y, o2 = s_off[4 * g + 2] + cl.z,
                o3 = s_off[4 * g + 3] + cl.w;
#pragma unroll
      for (int c = 0; c < CB; ++c) {
        const double* t = s_mix + (size_t)c * cw + xw;
        d[c] += (t[o0] + t[o1]) + (t[o2] + t[o3]);
      }
    }
    const double ck = cst[k];
#pragma unroll
    for (int c = 0; c < CB; ++c) {
      const double L = (ck - q[c]) + d[c];
      const double dl = L - m[c];
      if (dl > 0.0) {            // (m = -inf: exp(-inf) = 0)
        s[c] = fma(s[c], exp(-dl), 1.0);
        m[c] = L;
      } else if (dl > -skip) {
        s[c] += exp(dl);
      } else if (L != L) {
        m[c] = L;
        s[c] = L;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CB; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double m2 = __shfl_xor_sync(0xffffffffu, m[c], o), s2 = __shfl_xor_sync(0xffffffffu, s[c], o);
      lse_merge(m2, s2, m[c], s[c]);
    }
    if ((threadIdx.x & 31) == 0) {
      s_m[threadIdx.x >> 5][c] = m[c];
      s_s[threadIdx.x >> 5][c] = s[c];
    }
  }
  __syncthreads();
  if (threadIdx.x < CB && ct0 + threadIdx.x < Ct) {
    const int c = threadIdx.x;
    double mm = s_m[0][c], ss = s_s[0][c];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) lse_merge(s_m[w][c], s_s[w][c], mm, ss);
    if (oob != nullptr && oob[ct0 + c]) {   // outside the support: every kernel gives -inf (k_logpdf_pairs, kind 2)
      mm = -INFINITY;
      ss = 0.0;
    }
    part[(int64_t)blockIdx.y * ct_stride + ct0 + c] = make_double2(mm, ss);
  }
}

}  // namespace tpe
