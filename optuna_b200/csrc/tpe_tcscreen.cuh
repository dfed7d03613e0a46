// The multivariate g(x) grid, screened on the bf16 tensor-core path (all-continuous spaces, CONST tables).
//
// L(c, k) = cst_k - |x''_c - mu''_k|^2 / 2 = (cst_k - |mu''_k|^2 / 2) + (-|x''_c|^2 / 2) + <x''_c, mu''_k>.
// Only the terms within `skip` of a candidate's largest term enter its log-sum-exp: at config 2
// that is a few per cent of the C x K cells, yet k_logpdf_mma computes every cell sum in fp64.  Here the inner product
// is first taken in bf16 x bf16 -> fp32 (mma.sync.m16n8k16: exact products, fp32 accumulation), which places every cell
// within a rigorous distance delta of its true value:
//   |<x, mu>_bf16 - <x, mu>| <= sum_j |x_j| |mu_j| (2^-8 + 2^-16) + (P + 2) 2^-24 sum_j |x_j| |mu_j|
// (two roundings to bf16, 2^-9 relative each; fp32 accumulation), with |x''_j|, |mu''_j| <= rho = (range / 2) / sigma
// for observations and candidates inside the support (the caller checks) -- the host passes window = skip + 2 delta.
//   pass 1 (k_tcs_max): m~_c = max_k L~(c, k), so the true maximum lies in [m~_c - delta, m~_c + delta];
//   pass 2 (k_tcs_sum): the same products again; a cell with L~ >= m~_c - window is a survivor (every cell with
//          L >= max - skip is one), flagged in a bit mask; the survivors of a (128 x 128) tile are then evaluated EXACTLY
//          in fp64
//          and summed as e^(L - m~_c): four lanes own a candidate row for the whole kernel (a quarter of its x'' each, in
//          registers), walk the row's mask and take the fp64 inner product with the survivor's mu'' row from shared
//          memory (L = (cst_k - |mu''|^2 / 2) + (-|x''|^2 / 2) + <x'', mu''>, the expanded form k_logpdf_mma uses under
//          the same conditioning bound); every partial sum has one owner and a fixed order, no atomics on sums.
// Cells that are not survivors lie more than `skip` below the maximum: the truncation every grid kernel applies.
#pragma once
#include <cuda_bf16.h>

#include "tpe_uni.cuh"

namespace tpe {

constexpr int kTcsRows = 128;   // candidates per CTA
constexpr int kTcsTile = 128;   // kernels per tile
constexpr int kTcsNT = 512;

__device__ __forceinline__ int tcs_ordered(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float tcs_unordered(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// per estimator: bf16 copy of the CONST table (row stride PB + 8: conflict-free fragment loads) and
// ak = cst - |mu''|^2 / 2 in fp32; rows >= kobs are neutral (ak = -inf)
__global__ void k_tcs_tables(const double* __restrict__ tabc, const double* __restrict__ cst, int64_t kobs, int64_t kpad,
                             int pb, __nv_bfloat16* __restrict__ tabh, float* __restrict__ ak,
                             double* __restrict__ ak64) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= kpad) return;
  const int stride = pb + 8;
  double sq = 0.0;
  for (int j = 0; j < stride; ++j) {
    const double v = (k < kobs && j < pb) ? tabc[k * pb + j] : 0.0;
    sq = fma(v, v, sq);
    tabh[k * stride + j] = __double2bfloat16(v);
  }
  const double a = (k < kobs) ? cst[k] - 0.5 * sq : -INFINITY;
  ak[k] = __double2float_rn(a);
  ak64[k] = a;
}

// per ask: x'' = (x - ctr) / sigma in fp64 [ct_stride][pb], bc = -|x''|^2 / 2, gmax = -inf
__global__ void k_tcs_xprep(const double* __restrict__ xT, const double2* __restrict__ colprm, int64_t ct_stride, int pb,
                            double* __restrict__ x64, float* __restrict__ bc, double* __restrict__ bc64,
                            int* __restrict__ gmax) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= ct_stride) return;
  double sq = 0.0;
  for (int j = 0; j < pb; ++j) {
    const double2 p = colprm[j];
    const double v = (xT[(int64_t)j * ct_stride + c] - p.x) * p.y;
    x64[c * pb + j] = v;
    sq = fma(v, v, sq);
  }
  bc[c] = __double2float_rn(-0.5 * sq);
  bc64[c] = -0.5 * sq;
  gmax[c] = tcs_ordered(-INFINITY);
}

__device__ __forceinline__ void tcs_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void tcs_cp16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void tcs_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tcs_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int PB>
struct TcsSmem {
  static constexpr int HS = PB + 8;                    // bf16 row stride
  __nv_bfloat16 a[kTcsRows][HS];                       // candidates
  __nv_bfloat16 b[2][kTcsTile][HS];                    // kernel tiles (double buffered)
  float ak[2][kTcsTile];
  // k_tcs_sum only
  double b64[kTcsTile][PB];
  double c64[kTcsTile];                                // cst - |mu''|^2 / 2
  uint32_t mask[kTcsRows][4];
  double e64[64];
};

// issue the asynchronous copies of kernel tile `t` (bf16 rows + ak) into buffer `buf`
template <int PB>
__device__ __forceinline__ void tcs_load_tile(TcsSmem<PB>& sm, int buf, int64_t k0, const __nv_bfloat16* __restrict__ tabh,
                                              const float* __restrict__ ak) {
  constexpr int HS = PB + 8, CH = HS / 8;              // 16-byte chunks per row
  for (int i = threadIdx.x; i < kTcsTile * CH; i += kTcsNT) {
    const int r = i / CH, ch = i - r * CH;
    tcs_cp16(&sm.b[buf][r][ch * 8], tabh + (k0 + r) * HS + ch * 8);
  }
  if (threadIdx.x < kTcsTile / 4) tcs_cp16(&sm.ak[buf][threadIdx.x * 4], ak + k0 + threadIdx.x * 4);
}

// SUM == false: pass 1 (row maxima into gmax); SUM == true: pass 2 (survivor masks + exact sums into part)
//   grid = (ceil(Ct / 128), k-splits), 512 threads = 16 warps: warp (wm = w & 7, wn = w >> 3) owns rows 16 wm .. + 15 and
//   columns 64 wn .. + 63 of every 128 x 128 tile
template <int PB, bool SUM>
__global__ void __launch_bounds__(kTcsNT, (PB >= 64 ? 1 : 2))
k_tcs(const __nv_bfloat16* __restrict__ tabh, const float* __restrict__ ak, const double* __restrict__ tabc,
      const double* __restrict__ ak64, int64_t kpad, int64_t kps, const double* __restrict__ x64g,
      const float* __restrict__ bcg, const double* __restrict__ bc64g, int* __restrict__ gmax, int64_t Ct,
      int64_t ct_stride, float window, double2* __restrict__ part, unsigned long long* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char tcs_raw[];
  TcsSmem<PB>& sm = *reinterpret_cast<TcsSmem<PB>*>(tcs_raw);
  constexpr int KS = PB / 16;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp & 7, wn = warp >> 3;
  const int g = lane >> 2, tq = lane & 3;
  const int64_t c0 = (int64_t)blockIdx.x * kTcsRows;
  const int64_t k_begin = (int64_t)blockIdx.y * kps, k_end = (k_begin + kps < kpad) ? k_begin + kps : kpad;
  const int ntiles = (int)((k_end - k_begin + kTcsTile - 1) / kTcsTile);
  if (ntiles <= 0) return;
  tcs_load_tile<PB>(sm, 0, k_begin, tabh, ak);
  tcs_commit();
  // candidates: bf16 rows (and, for the sums, the fp64 rows)
  for (int i = threadIdx.x; i < kTcsRows * PB; i += kTcsNT) {
    const int r = i / PB, j = i - r * PB;
    const double v = x64g[(c0 + r) * PB + j];          // x64g has ct_stride (multiple of 1024) rows
    sm.a[r][j] = __double2bfloat16(v);
  }
  if (SUM && threadIdx.x < 64) sm.e64[threadIdx.x] = exp2((double)threadIdx.x * 0.015625);
  __syncthreads();
  const int row0 = wm * 16 + g, row1 = row0 + 8;
  const float bc0 = bcg[c0 + row0], bc1 = bcg[c0 + row1];
  float thr0 = 0.0f, thr1 = 0.0f, mx0 = -INFINITY, mx1 = -INFINITY;
  if (SUM) {   // survivors: acc + ak >= m~ - window - bc
    thr0 = tcs_unordered(gmax[c0 + row0]) - window - bc0;
    thr1 = tcs_unordered(gmax[c0 + row1]) - window - bc1;
  }
  // exact phase: the four lanes of a quad own candidate row 8 warp + (lane >> 2); lane i of the quad keeps dims
  // [i PB/4, (i + 1) PB/4) of x'' in registers, as PB/8 pairs in an order rotated by the quad's index (so that the
  // 128-bit loads of different quads fall into different banks)
  constexpr int NCH = PB / 8;
  const int qd = lane >> 2, qi = lane & 3;
  const int xrow = warp * 8 + qd;
  const unsigned qmask = 0xFu << (lane & ~3);
  double2 xs[NCH];
  double ref = 0.0, bcx = 0.0, ssum = 0.0, pend = 0.0;
  int npend = 0;
  if (SUM) {
#pragma unroll
    for (int sI = 0; sI < NCH; ++sI) {
      const int ch = (sI + qd) % NCH;
      xs[sI] = *reinterpret_cast<const double2*>(x64g + (c0 + xrow) * PB + qi * (PB / 4) + 2 * ch);
    }
    ref = (double)tcs_unordered(gmax[c0 + xrow]);
    bcx = bc64g[c0 + xrow];
  }
  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    const int64_t k0 = k_begin + (int64_t)t * kTcsTile;
    if (SUM) {   // fp64 copy of this tile (single buffer: the previous exact phase has finished)
      for (int i = threadIdx.x; i < kTcsTile * PB / 2; i += kTcsNT) {
        const int r = i / (PB / 2), ch = i - r * (PB / 2);
        tcs_cp16(&sm.b64[r][ch * 2], tabc + (k0 + r) * PB + ch * 2);
      }
      if (threadIdx.x < kTcsTile / 2) tcs_cp16(&sm.c64[threadIdx.x * 2], ak64 + k0 + threadIdx.x * 2);
      sm.mask[threadIdx.x >> 2][threadIdx.x & 3] = 0u;
    }
    if (t + 1 < ntiles) tcs_load_tile<PB>(sm, buf ^ 1, k0 + kTcsTile, tabh, ak);
    tcs_commit();
    // everything but the group just committed has landed (pass 2: tile t's bf16 rows; its fp64 rows are in the
    // group just committed together with tile t + 1 and are waited for before the exact phase)
    tcs_wait<1>();
    __syncthreads();
    uint32_t af[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int r = wm * 16 + g, kk = ks * 16 + tq * 2;
      af[ks][0] = *reinterpret_cast<const uint32_t*>(&sm.a[r][kk]);
      af[ks][1] = *reinterpret_cast<const uint32_t*>(&sm.a[r + 8][kk]);
      af[ks][2] = *reinterpret_cast<const uint32_t*>(&sm.a[r][kk + 8]);
      af[ks][3] = *reinterpret_cast<const uint32_t*>(&sm.a[r + 8][kk + 8]);
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int n0 = wn * 64 + nt * 8;
      float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int kk = ks * 16 + tq * 2;
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&sm.b[buf][n0 + g][kk]);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&sm.b[buf][n0 + g][kk + 8]);
        tcs_mma(acc, af[ks], b0, b1);
      }
      const int col = n0 + tq * 2;
      const float a0 = sm.ak[buf][col], a1 = sm.ak[buf][col + 1];
      const float v00 = acc[0] + a0, v01 = acc[1] + a1, v10 = acc[2] + a0, v11 = acc[3] + a1;
      if (!SUM) {
        mx0 = fmaxf(mx0, fmaxf(v00, v01));
        mx1 = fmaxf(mx1, fmaxf(v10, v11));
      } else {
        const uint32_t m0 = (v00 >= thr0 ? 1u : 0u) | (v01 >= thr0 ? 2u : 0u);
        const uint32_t m1 = (v10 >= thr1 ? 1u : 0u) | (v11 >= thr1 ? 2u : 0u);
        if (m0) atomicOr(&sm.mask[row0][col >> 5], m0 << (col & 31));
        if (m1) atomicOr(&sm.mask[row1][col >> 5], m1 << (col & 31));
      }
    }
    if (SUM) {
      tcs_wait<0>();
      __syncthreads();
      const uint32_t myword = sm.mask[xrow][qi];
#pragma unroll 1
      for (int ws = 0; ws < 4; ++ws) {
        uint32_t w = __shfl_sync(qmask, myword, (lane & ~3) | ws);
        while (w) {
          const int kl = ws * 32 + __ffs(w) - 1;
          w &= w - 1;
          const double* brow = &sm.b64[kl][qi * (PB / 4)];
          double g0 = 0.0, g1 = 0.0;
#pragma unroll
          for (int sI = 0; sI < NCH; ++sI) {
            const int ch = (sI + qd) % NCH;
            const double2 m = *reinterpret_cast<const double2*>(brow + 2 * ch);
            g0 = fma(xs[sI].x, m.x, g0);
            g1 = fma(xs[sI].y, m.y, g1);
          }
          double gsum = g0 + g1;
          gsum += __shfl_xor_sync(qmask, gsum, 1);
          gsum += __shfl_xor_sync(qmask, gsum, 2);
          const double L = (sm.c64[kl] + bcx) + gsum;
          if ((npend & 3) == qi) pend = L;          // lane i of the quad keeps every fourth survivor ...
          ++npend;
          if ((npend & 3) == 0) ssum += uni_exp(fmax(pend - ref, -700.0), sm.e64);   // ... and the four exps run together
        }
      }
    }
    __syncthreads();
  }
  if (!SUM) {
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    if (tq == 0) {
      if (mx0 > -INFINITY) atomicMax(&gmax[c0 + row0], tcs_ordered(mx0 + bc0));
      if (mx1 > -INFINITY) atomicMax(&gmax[c0 + row1], tcs_ordered(mx1 + bc1));
    }
  } else {
    if (stats != nullptr && qi == 0) atomicAdd(stats, (unsigned long long)npend);   // diagnostics: survivors
    if (qi < (npend & 3)) ssum += uni_exp(fmax(pend - ref, -700.0), sm.e64);   // the last, incomplete group of four
    const double s1 = ssum + __shfl_xor_sync(0xffffffffu, ssum, 1);
    const double s2 = s1 + __shfl_xor_sync(0xffffffffu, s1, 2);
    if (qi == 0 && c0 + xrow < Ct)
      part[(int64_t)blockIdx.y * ct_stride + c0 + xrow] =
          (s2 > 0.0 && ref > -INFINITY) ? make_double2(ref, s2) : make_double2(-INFINITY, 0.0);
  }
}

}  // namespace tpe
