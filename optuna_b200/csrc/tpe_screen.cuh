// fp32-screened log-density kernel for wide continuous spaces (multivariate TPE, 17..32 columns).
//
// The exact kernel (k_logpdf_fast) is bound by the fp64 pipe at 2 instructions per
// (candidate, kernel, param) cell, yet only the ~5 % of kernels within `skip` of a candidate's running
// maximum contribute to its log-sum-exp at fp64 resolution.  This kernel therefore
//   1. SCREENS every (candidate, kernel) pair on the fp32 pipe with ONE FFMA per cell:
//        L~ = d_k + x.mu_k - |x|^2 / 2,   d_k = cst_k - |mu_k|^2 / 2      (centred, scaled coordinates)
//      and parks the pairs with L~ > running max - skip - margin, where `margin` bounds the fp32
//      evaluation error (dot product of 32 terms + input rounding), so the parked set is a superset of
//      the pairs the exact kernel would have kept;
//   2. RE-EVALUATES the parked pairs exactly in fp64 -- same arithmetic as k_logpdf_fast, from the
//      fp64 copy of the kernel tile and of the candidates held in shared memory -- and folds them into
//      the fp64 online log-sum-exp, warp-synchronously at tile boundaries (or when a buffer fills).
//   3. Terms between `near` = 34 and `skip` below the running max are smaller than e^-34 = 1.7e-15 of
//      the largest term: they are added from the fp32 value directly (relative error of such a term
//      <= 1e-3, i.e. < 2e-18 of the largest term each, < 2e-13 for 1e5 of them) -- no re-evaluation.
// Every term within `near` of the maximum enters the sum as an fp64 value identical to the exact
// kernel's; dropped terms obey the same truncation bound.  The fp64 pipe only sees the ~0.5 % near terms.
#pragma once
#include "tpe_common.cuh"

namespace tpe {

constexpr int kScrP = 32;        // columns (zero padded)
constexpr int kScrStride = 33;   // fp64 tile row stride in doubles: rows of different kernels fall in different banks
constexpr int kScrNT = 256;      // threads per CTA (one CTA per SM: 128 KB of fp64 candidates in shared memory)
constexpr int kScrRC = 2;        // candidates per lane
constexpr int kScrTK = 64;       // kernels per tile
constexpr int kScrST = 2;        // pipeline stages
constexpr int kScrCands = kScrNT * kScrRC;
constexpr int kScrBuf = 4;       // parked kernels per candidate between two folds
constexpr float kScrNear = 34.0f;  // terms further than this below the running max are taken from the fp32 value

struct ScreenSmem {
  double xs[kScrP][kScrCands];                       // fp64 candidates, [param][candidate]
  double t64[kScrST][kScrTK * kScrStride + 2];       // fp64 tile, padded rows
  double c64[kScrST][kScrTK];                        // exact per-kernel constants
  float t32[kScrST][kScrTK][kScrP];                  // fp32 tile
  float d32[kScrST][kScrTK];                         // screening constants
  uint64_t full[kScrST];
  int done[kScrST];   // warps that finished the tile currently in the stage
  double skip;
};

// candidates: centred / scaled coordinates in both precisions + the fp32 self term
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void k_screen_xprep(const double* __restrict__ xT, const double2* __restrict__ colprm, int64_t ct_stride,
                               double* __restrict__ x64, float* __restrict__ x32, float* __restrict__ e32,
                               float* __restrict__ gmax) {
  const int64_t ct = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (ct >= ct_stride) return;
  gmax[ct] = -INFINITY;
  double sq = 0.0;
  for (int p = 0; p < kScrP; ++p) {
    const double2 cp = colprm[p];
    const double v = (xT[(int64_t)p * ct_stride + ct] - cp.x) * cp.y;
    x64[(int64_t)p * ct_stride + ct] = v;
    x32[(int64_t)p * ct_stride + ct] = (float)v;
    sq = fma(v, v, sq);
  }
  e32[ct] = (float)(-0.5 * sq);
}
// kernels: fp32 table, padded fp64 table, screening constant d_k = cst_k - |mu_k|^2 / 2
__global__ void k_screen_tabprep(const double* __restrict__ tabc, const double* __restrict__ cst, int64_t Kf,
                                 float* __restrict__ tab32, double* __restrict__ tab64p, float* __restrict__ d32) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t k = warp; k < Kf; k += nwarps) {
    const double v = tabc[k * kScrP + lane];
    tab32[k * kScrP + lane] = (float)v;
    tab64p[k * kScrStride + lane] = v;
    if (lane == 0) tab64p[k * kScrStride + kScrP] = 0.0;
    const double sq = warp_sum(v * v);
    if (lane == 0) d32[k] = (float)(cst[k] - 0.5 * sq);
  }
}

__global__ void __launch_bounds__(kScrNT, 1)
k_logpdf_screen(const float* __restrict__ tab32, const double* __restrict__ tab64p, const double* __restrict__ cst,
                const float* __restrict__ d32, int64_t Kf, const double* __restrict__ x64, const float* __restrict__ x32,
                const float* __restrict__ e32, float* gmax, int64_t ct_stride, int64_t kps,
                double lse_skip, float margin, double2* __restrict__ part) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ScreenSmem& sm = *reinterpret_cast<ScreenSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t k0 = blockIdx.y * kps;
  const int64_t k1 = (k0 + kps < Kf) ? k0 + kps : Kf;
  const int ntiles = (k1 > k0) ? (int)((k1 - k0 + kScrTK - 1) / kScrTK) : 0;
  const int64_t cbase = (int64_t)blockIdx.x * kScrCands;
  // candidate r of this lane: local index cl[r] = warp * 64 + r * 32 + lane
  int cl[kScrRC];
#pragma unroll
  for (int r = 0; r < kScrRC; ++r) cl[r] = warp * (32 * kScrRC) + r * 32 + lane;

  if (tid == 0) {
    for (int s = 0; s < kScrST; ++s) {
      mbar_init(&sm.full[s], 1);
      sm.done[s] = 0;
    }
    mbar_fence_init();
    sm.skip = lse_skip;
  }
  __syncthreads();
  auto issue = [&](int t) {
    const int st = t % kScrST;
    const int64_t ks = k0 + (int64_t)t * kScrTK;
    const int tk = (int)((k1 - ks < kScrTK) ? (k1 - ks) : kScrTK);
    const uint32_t b64 = (uint32_t)(((size_t)tk * kScrStride * 8 + 15) & ~(size_t)15);
    const uint32_t b32 = (uint32_t)((size_t)tk * kScrP * 4);
    const uint32_t bc = (uint32_t)(((tk + 1) & ~1) * 8);
    const uint32_t bd = (uint32_t)(((tk + 3) & ~3) * 4);
    fence_proxy_async();
    mbar_expect_tx(&sm.full[st], b64 + b32 + bc + bd);
    bulk_g2s(sm.t64[st], tab64p + ks * kScrStride, b64, &sm.full[st]);
    bulk_g2s(sm.t32[st], tab32 + ks * kScrP, b32, &sm.full[st]);
    bulk_g2s(sm.c64[st], cst + ks, bc, &sm.full[st]);
    bulk_g2s(sm.d32[st], d32 + ks, bd, &sm.full[st]);
  };
  if (tid == 0)
    for (int t = 0; t < kScrST && t < ntiles; ++t) issue(t);

  // candidates: fp32 copy in registers, fp64 copy in shared memory
  float x[kScrRC][kScrP];
  float thr[kScrRC];
#pragma unroll
  for (int p = 0; p < kScrP; ++p)
#pragma unroll
    for (int r = 0; r < kScrRC; ++r) {
      x[r][p] = x32[(int64_t)p * ct_stride + cbase + cl[r]];
      sm.xs[p][cl[r]] = x64[(int64_t)p * ct_stride + cbase + cl[r]];
    }
  float ec[kScrRC], thr_hi[kScrRC], mref[kScrRC], sfar[kScrRC];
  double mx[kScrRC], sum[kScrRC], mglob[kScrRC];  // mglob: best exact log-term any k-split of this candidate has seen
  uint32_t buf[kScrRC];
  int cnt[kScrRC];
#pragma unroll
  for (int r = 0; r < kScrRC; ++r) {
    ec[r] = e32[cbase + cl[r]];
    mx[r] = -INFINITY;
    mglob[r] = -INFINITY;
    sum[r] = 0.0;
    sfar[r] = 0.f;
    buf[r] = 0;
    cnt[r] = 0;
    thr[r] = -INFINITY;     // below this the term is dropped
    thr_hi[r] = -INFINITY;  // above this the kernel is parked for the exact fp64 pass
    mref[r] = 0.f;          // running max in the screening frame: (float)(max) - e_c
  }
  __syncthreads();

  for (int t = 0; t < ntiles; ++t) {
    const int st = t % kScrST;
    mbar_wait(&sm.full[st], (uint32_t)((t / kScrST) & 1));
    const int64_t ks = k0 + (int64_t)t * kScrTK;
    const int tk = (int)((k1 - ks < kScrTK) ? (k1 - ks) : kScrTK);
    // shared reference maximum: read at the start of the tile, consumed at its end (latency hidden)
    float gpre[kScrRC];
#pragma unroll
    for (int r = 0; r < kScrRC; ++r) gpre[r] = __ldcg(&gmax[cbase + cl[r]]);
    // exact re-evaluation of the parked kernels of this tile; the whole warp folds together
    auto fold_parked = [&](bool tile_end) {
#pragma unroll
      for (int r = 0; r < kScrRC; ++r) {
        for (int e = 0; e < kScrBuf; ++e) {
          if (!__any_sync(0xffffffffu, cnt[r] > e)) break;
          const bool valid = cnt[r] > e;
          const int kk = valid ? (int)((buf[r] >> (8 * e)) & 0xffu) : 0;
          const double* row = sm.t64[st] + kk * kScrStride;
          double a0 = 0.0, a1 = 0.0;
#pragma unroll
          for (int p = 0; p < kScrP; p += 2) {
            const double d0 = sm.xs[p][cl[r]] - row[p];
            const double d1 = sm.xs[p + 1][cl[r]] - row[p + 1];
            a0 = fma(d0, d0, a0);
            a1 = fma(d1, d1, a1);
          }
          const double L = fma(-0.5, a0 + a1, sm.c64[st][kk]);
          const double d = L - mx[r];
          const double ex = exp(-fabs(d));
          const bool bigger = valid && (d > 0.0);
          const double grown = fma(sum[r], ex, 1.0);
          const double added = sum[r] + ex;
          sum[r] = bigger ? grown : ((valid && d == d) ? added : sum[r]);
          mx[r] = bigger ? L : mx[r];
        }
        cnt[r] = 0;
        buf[r] = 0;
        // Reference maximum for the thresholds: the best exact term seen by ANY k-split of this
        // candidate (shared through gmax), a valid lower bound of the final maximum.  Far terms are
        // accumulated relative to it.
        if (tile_end && mx[r] != -INFINITY) atomic_max_float(&gmax[cbase + cl[r]], __double2float_rd(mx[r]));
        const double g = (double)gpre[r];  // L2 value published by the other k-splits
        const double ref_new = g > mx[r] ? g : mx[r];
        if (ref_new != -INFINITY) {
          if (mglob[r] != -INFINITY && ref_new > mglob[r]) sfar[r] *= __expf((float)(mglob[r] - ref_new));
          mglob[r] = ref_new;
          mref[r] = __double2float_rd(ref_new) - ec[r];
          thr[r] = mref[r] - (float)sm.skip - margin;
          thr_hi[r] = mref[r] - kScrNear - margin;
        }
      }
    };
    for (int kk = 0; kk < tk; ++kk) {
      const float4* row = reinterpret_cast<const float4*>(&sm.t32[st][kk][0]);
      float a0[kScrRC], a1[kScrRC], a2[kScrRC], a3[kScrRC];
#pragma unroll
      for (int r = 0; r < kScrRC; ++r) a0[r] = a1[r] = a2[r] = a3[r] = 0.f;
#pragma unroll
      for (int q = 0; q < kScrP / 4; ++q) {
        const float4 v = row[q];
#pragma unroll
        for (int r = 0; r < kScrRC; ++r) {
          a0[r] = fmaf(x[r][4 * q], v.x, a0[r]);
          a1[r] = fmaf(x[r][4 * q + 1], v.y, a1[r]);
          a2[r] = fmaf(x[r][4 * q + 2], v.z, a2[r]);
          a3[r] = fmaf(x[r][4 * q + 3], v.w, a3[r]);
        }
      }
      const float dk = sm.d32[st][kk];
      float S[kScrRC];
      bool hit = false;
#pragma unroll
      for (int r = 0; r < kScrRC; ++r) {
        S[r] = (a0[r] + a1[r]) + (a2[r] + a3[r]) + dk;
        hit = hit || (S[r] > thr[r]);
      }
      if (!__any_sync(0xffffffffu, hit)) continue;  // the common case: nothing within `skip` of any maximum
      bool full = false;
#pragma unroll
      for (int r = 0; r < kScrRC; ++r) {
        if (S[r] > thr[r]) {
          if (S[r] > thr_hi[r]) {  // near the maximum: exact fp64 re-evaluation at the next fold
            buf[r] = (buf[r] << 8) | (uint32_t)kk;
            ++cnt[r];
          } else {                 // far: the fp32 value is accurate enough for a term < e^-34 of the maximum
            sfar[r] += __expf(S[r] - mref[r]);
          }
        }
        full = full || (cnt[r] == kScrBuf);
      }
      if (__any_sync(0xffffffffu, full)) fold_parked(false);
    }
    fold_parked(true);  // also publishes / adopts the shared reference maximum once per tile
    // No block barrier: the last warp to leave the tile refills its stage with tile t + ST.
    __syncwarp();
    if (lane == 0) {
      const int prev = atomicAdd(&sm.done[st], 1);
      if (prev == (kScrNT / 32) - 1) {
        sm.done[st] = 0;
        if (t + kScrST < ntiles) issue(t + kScrST);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kScrRC; ++r) {
    // partial in the reference frame: sum_exact * e^(mx - ref) + sum_far   (mx <= ref)
    double m = mglob[r], sacc = (double)sfar[r];
    if (mx[r] != -INFINITY) sacc += sum[r] * exp(mx[r] - mglob[r]);
    if (m == -INFINITY) sacc = 0.0;
    part[blockIdx.y * ct_stride + cbase + cl[r]] = make_double2(m, sacc);
  }
}

}  // namespace tpe
