// Univariate TPE, all continuous columns of a trial stage by stage: ONE launch per stage over all columns
// (blockIdx.y = column) instead of one launch per stage per column.  Every kernel here is the single-column kernel's
// body (d_* in tpe_kernels.cuh / tpe_uni.cuh) applied to column y's slice of strided buffers, so the numbers are those
// of the per-column path bit for bit; the mixture weights do not depend on the column and are computed once.
#pragma once
#include "tpe_uni.cuh"

namespace tpe {

// strides (elements) of the per-column slices
struct UbDims {
  int64_t ks;      // kernels (mu, sigma, cst_part, cst, order, sorted tables, tile metadata x 128)
  int64_t cs;      // candidates (S, xT, oob, sorted candidates, part / fix / logl / logg): ct_stride
  int32_t pall;    // columns of the history matrix
};

// every column as the only one of its context (setup_columns with one column: slot and RNG rank 0)
__global__ void kb_cols(const ColMeta* __restrict__ in, int n, ColMeta* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  ColMeta cm = in[j];
  cm.slot = 0;
  cm.num_rank = 0;
  out[j] = cm;
}
__global__ void kb_mu(const double* __restrict__ X, UbDims d, const int64_t* __restrict__ rows, int64_t n,
                      const ColMeta* __restrict__ cols, double* __restrict__ mu) {
  d_mu(X, d.pall, rows, n, cols + blockIdx.y, 1, mu + blockIdx.y * d.ks);
}
__global__ void __launch_bounds__(1024, 1)
kb_sort_small(const double* __restrict__ mu, UbDims d, int m, int m2, int32_t* __restrict__ order) {
  d_sort_small(mu + blockIdx.y * d.ks, 1, 0, m, m2, order + blockIdx.y * d.ks);
}
// cooperative: grid = (G, columns); key / index scratch strided by 2 ks / 3 ks
__global__ void __launch_bounds__(512, 1)
kb_radix_sort(const double* __restrict__ mu, UbDims d, int n, uint64_t* __restrict__ keys, int32_t* __restrict__ idx,
              SortWork* __restrict__ wk, int32_t* __restrict__ order, const int* __restrict__ run_flag) {
  const int64_t y = blockIdx.y;
  uint64_t* ka = keys + y * 2 * d.ks;
  int32_t* ia = idx + y * 2 * d.ks;
  d_radix_sort_coop(mu + y * d.ks, 1, 0, n, ka, ka + d.ks, ia, ia + d.ks, wk + y, order + y * d.ks, run_flag);
}
__global__ void __launch_bounds__(256)
kb_order_update(const int* __restrict__ mode, const int32_t* __restrict__ old_order, UbDims d, int K_old, int K_new,
                const double* __restrict__ mu, int32_t* __restrict__ out, int* __restrict__ work) {
  d_order_update(mode, old_order + blockIdx.y * d.ks, K_old, K_new, mu + blockIdx.y * d.ks, out + blockIdx.y * d.ks,
                 work + 4 * blockIdx.y);
}
__global__ void kb_sigma_uni(const double* __restrict__ mu, const int32_t* __restrict__ order, UbDims d,
                             const ColMeta* __restrict__ cols, int64_t n, int magic_clip, int endpoints,
                             double* __restrict__ sigma) {
  d_sigma_uni(mu + blockIdx.y * d.ks, order + blockIdx.y * d.ks, cols + blockIdx.y, 1, 0, n, magic_clip, endpoints,
              sigma + blockIdx.y * d.ks);
}
// k_const for one continuous column, thread = kernel (the warp sum of k_const adds zeros to the one term)
__global__ void kb_const(const double* __restrict__ mu, const double* __restrict__ sigma, UbDims d,
                         const ColMeta* __restrict__ cols, int64_t K, double* __restrict__ cst_part) {
  const ColMeta cm = cols[blockIdx.y];
  const double* mu_y = mu + blockIdx.y * d.ks;
  const double* sg_y = sigma + blockIdx.y * d.ks;
  double* out = cst_part + blockIdx.y * d.ks;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < K; k += (int64_t)gridDim.x * blockDim.x) {
    const double m = mu_y[k], s = sg_y[k];
    const double a = TPE_DIV(TPE_SUB(cm.klow, m), s);
    const double b = TPE_DIV(TPE_SUB(cm.khigh, m), s);
    const double mass = log_gauss_mass_fast(a, b);
    double acc = 0.0;
    acc += kLogSqrt2Pi + mass + log(s);
    out[k] = -acc;
  }
}
// cst = cst_part + ln w (k_wfinal / k_weights_one: c = cst_part[k] + lw), -inf padding up to the stride
__global__ void kb_cst(const double* __restrict__ cst_part, const double* __restrict__ logw, UbDims d, int64_t K,
                       double* __restrict__ cst) {
  const double* in = cst_part + blockIdx.y * d.ks;
  double* out = cst + blockIdx.y * d.ks;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < d.ks; k += (int64_t)gridDim.x * blockDim.x)
    out[k] = (k < K) ? in[k] + logw[k] : -INFINITY;
}
__global__ void __launch_bounds__(kUniTile)
kb_uni_tables(const int32_t* __restrict__ order, const double* __restrict__ mu, const double* __restrict__ sigma,
              const double* __restrict__ cst, UbDims d, const ColMeta* __restrict__ cols, int64_t K,
              float4* __restrict__ s32, double2* __restrict__ smi, double* __restrict__ sc,
              UniTileMeta* __restrict__ meta, int fgt, int magic_clip, int32_t* __restrict__ bstart) {
  const int64_t y = blockIdx.y;
  d_uni_tables(order + y * d.ks, mu + y * d.ks, sigma + y * d.ks, cst + y * d.ks, cols + y, K, s32 + y * d.ks,
               smi + y * d.ks, sc + y * d.ks, meta + y * (d.ks / kUniTile), fgt, magic_clip,
               bstart + y * (kFgtMaxBoxes + 1));
}
__global__ void __launch_bounds__(128)
kb_fgt_coeff(const int32_t* __restrict__ order, const double* __restrict__ mu, const double* __restrict__ sigma,
             const double* __restrict__ cst, UbDims d, const ColMeta* __restrict__ cols, int64_t K, int magic_clip,
             const int32_t* __restrict__ bstart, double* __restrict__ coef, FgtBox* __restrict__ box) {
  const int64_t y = blockIdx.y;
  d_fgt_coeff(order + y * d.ks, mu + y * d.ks, sigma + y * d.ks, cst + y * d.ks, cols + y, K, magic_clip,
              bstart + y * (kFgtMaxBoxes + 1), coef + y * kFgtMaxBoxes * kFgtRow, box + y * kFgtMaxBoxes);
}
// U [columns][2 C]: the stretch of uniforms of column y
__global__ void kb_sample(const double* __restrict__ U, int32_t C, UbDims d, const ColMeta* __restrict__ cols,
                          const double* __restrict__ cdf, int64_t Kb, const double* __restrict__ mu,
                          const double* __restrict__ sigma, double* __restrict__ S, double* __restrict__ xT,
                          uint8_t* __restrict__ oob) {
  const int64_t y = blockIdx.y;
  d_sample(U + y * 2 * C, 1, C, cols + y, 1, 0, 1, cdf, Kb, mu + y * d.ks, sigma + y * d.ks, nullptr, S + y * d.cs,
           xT + y * d.cs, d.cs, oob + y * d.cs);
}
__global__ void __launch_bounds__(1024, 1)
kb_uni_sort_cands(const double* __restrict__ xT, int C, UbDims d, const ColMeta* __restrict__ cols,
                  double* __restrict__ xs, int32_t* __restrict__ cidx) {
  d_uni_sort_cands(xT + blockIdx.y * d.cs, C, cols + blockIdx.y, xs + blockIdx.y * d.cs, cidx + blockIdx.y * d.cs);
}
// part [columns][2 cs]: slice 0 the pairwise grid, slice 1 the fast Gauss transform
__global__ void __launch_bounds__(kUniWarps * 32)
kb_uni_grid(const float4* __restrict__ s32, const double2* __restrict__ smi, const double* __restrict__ sc,
            const UniTileMeta* __restrict__ meta, UbDims d, int64_t K, const double* __restrict__ xs,
            const int32_t* __restrict__ cidx, int C, double lse_skip, double2* __restrict__ part,
            const int32_t* __restrict__ tlist) {
  const int64_t y = blockIdx.y;
  d_uni_grid(s32 + y * d.ks, smi + y * d.ks, sc + y * d.ks, meta + y * (d.ks / kUniTile), K, xs + y * d.cs,
             cidx + y * d.cs, C, lse_skip, part + y * 2 * d.cs,
             tlist != nullptr ? tlist + y * (d.ks / kUniTile + 1) : nullptr);
}
__global__ void __launch_bounds__(256)
kb_uni_tile_list(const UniTileMeta* __restrict__ meta, UbDims d, int ntiles, int32_t* __restrict__ list) {
  d_uni_tile_list(meta + blockIdx.y * (d.ks / kUniTile), ntiles, list + blockIdx.y * (d.ks / kUniTile + 1));
}
// k_acq for at most two slices per estimator, thread = candidate: the merges lane 0 of k_acq's warp performs
__global__ void kb_acq2(const double2* __restrict__ part_l, int nsl, const double2* __restrict__ part_g, int nsg,
                        UbDims d, const uint8_t* __restrict__ oob, const double2* __restrict__ fix_l,
                        const double2* __restrict__ fix_g, int64_t Ct, double* __restrict__ logl,
                        double* __restrict__ logg) {
  const int64_t y = blockIdx.y;
  const int64_t ct = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (ct >= Ct) return;
  const int64_t o1 = y * d.cs + ct, o2 = y * 2 * d.cs + ct;
  double ml = -INFINITY, sl = 0.0, mg = -INFINITY, sg = 0.0;
  if (oob[o1]) {
    ml = fix_l[o1].x; sl = fix_l[o1].y;
    mg = fix_g[o1].x; sg = fix_g[o1].y;
  } else {
    for (int s = 0; s < nsl; ++s) {
      const double2 v = part_l[o2 + (int64_t)s * d.cs];
      lse_merge(v.x, v.y, ml, sl);
    }
    for (int s = 0; s < nsg; ++s) {
      const double2 v = part_g[o2 + (int64_t)s * d.cs];
      lse_merge(v.x, v.y, mg, sg);
    }
  }
  logl[o1] = (ml == -INFINITY) ? -INFINITY : log(sl) + ml;
  logg[o1] = (mg == -INFINITY) ? -INFINITY : log(sg) + mg;
}
__global__ void __launch_bounds__(256)
kb_fgt_eval(const double* __restrict__ coef, const FgtBox* __restrict__ box, const int32_t* __restrict__ bstart,
            const float4* __restrict__ s32, const double2* __restrict__ smi, const double* __restrict__ sc,
            const double* __restrict__ mu, const double* __restrict__ sigma, const double* __restrict__ cst, UbDims d,
            const ColMeta* __restrict__ cols, int64_t K, int magic_clip, const double* __restrict__ xT, int C,
            double2* __restrict__ part) {
  const int64_t y = blockIdx.y;
  d_fgt_eval(coef + y * kFgtMaxBoxes * kFgtRow, box + y * kFgtMaxBoxes, bstart + y * (kFgtMaxBoxes + 1),
             s32 + y * d.ks, smi + y * d.ks, sc + y * d.ks, mu + y * d.ks, sigma + y * d.ks, cst + y * d.ks, cols + y, K,
             magic_clip, xT + y * d.cs, C, part + y * 2 * d.cs + d.cs);
}
__global__ void kb_prior_fix(const double* __restrict__ S, int64_t Ct, UbDims d, const ColMeta* __restrict__ cols,
                             const double* __restrict__ mu, const double* __restrict__ sigma,
                             const double* __restrict__ cst, int64_t K, const uint8_t* __restrict__ oob,
                             double2* __restrict__ fix) {
  const int64_t y = blockIdx.y;
  d_logpdf_prior_fix(S + y * d.cs, Ct, cols + y, 1, mu + y * d.ks, sigma + y * d.ks, cst + y * d.ks, K, nullptr, nullptr,
                     oob + y * d.cs, fix + y * d.cs, 1);
}
__global__ void kb_acq(const double2* __restrict__ part_l, int nsl, const double2* __restrict__ part_g, int nsg,
                       UbDims d, const uint8_t* __restrict__ oob, const double2* __restrict__ fix_l,
                       const double2* __restrict__ fix_g, int64_t Ct, double* __restrict__ logl,
                       double* __restrict__ logg) {
  const int64_t y = blockIdx.y;
  d_acq(part_l + y * 2 * d.cs, nsl, part_g + y * 2 * d.cs, nsg, d.cs, oob + y * d.cs, fix_l + y * d.cs, fix_g + y * d.cs,
        Ct, logl + y * d.cs, logg + y * d.cs);
}
// grid = (1, columns): one ask per column
__global__ void __launch_bounds__(256)
kb_select(const double* __restrict__ logl, const double* __restrict__ logg, int32_t C, UbDims d,
          const double* __restrict__ S, double* __restrict__ out_x, double* __restrict__ out_acq,
          int64_t* __restrict__ out_best) {
  const int64_t y = blockIdx.y;
  d_select(logl + y * d.cs, logg + y * d.cs, C, S + y * d.cs, 1, out_x + y, out_acq + y, out_best + y);
}

}  // namespace tpe
