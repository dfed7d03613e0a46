// Shared device helpers: mbarrier / TMA bulk-copy wrappers (sm_100a), block scans, key ordering.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tpe {

// ---- per-call column metadata (one entry per selected search-space column) ---------------------
enum ColClass : int32_t { COL_CONT = 0, COL_DISC = 1, COL_CAT = 2 };

struct ColMeta {
  int32_t cls;        // ColClass
  int32_t log;        // log-scaled
  int32_t nch;        // categorical: number of choices
  int32_t src;        // column in the history matrix
  int32_t slot;       // index among continuous columns (fast kernel), rank among cat / numeric otherwise
  int32_t num_rank;   // rank among numeric columns (RNG block order), -1 for categorical
  int32_t cat_rank;   // rank among categorical columns, -1 otherwise
  int32_t tab_off;    // categorical: offset (in doubles) of this column's (nch+1) x nch table
  int32_t dist_off;   // categorical distance table offset in ctx->cat_dist or -1
  int32_t grid;       // discrete column tabulated over its grid: number of grid values, else 0
  int64_t dtab_off;   // offset (doubles) of its grid x (grid + 1) table of cell masses
  double low, high, step;  // as given by the distribution (step = 0 when continuous)
  double klow, khigh;      // kernel-space support: (low - step/2, high + step/2), log applied
};

// ---- sm_100a async-copy primitives ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      " selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// TMA bulk (1-D) global -> shared copy; completion is signalled on `bar` (complete_tx::bytes).
// dst, src 16-byte aligned, bytes a multiple of 16.  SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- ordering of fp64 keys as unsigned integers ---------------------------------------------------
// Python's sorted() treats -0.0 == 0.0 (ties keep trial order), so fold -0.0 onto +0.0 first.
__device__ __forceinline__ uint64_t order_bits(double v) {
  if (v == 0.0) v = 0.0;
  const uint64_t u = static_cast<uint64_t>(__double_as_longlong(v));
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

// inverse of order_bits; the all-zero pattern (a freshly cleared slot) reads as -inf
__device__ __forceinline__ double from_order_bits(uint64_t b) {
  if (b == 0) return -INFINITY;
  const uint64_t u = (b >> 63) ? (b & 0x7fffffffffffffffull) : ~b;
  return __longlong_as_double(static_cast<long long>(u));
}

// ---- block-wide (1024 threads) ordered rank of flagged threads --------------------------------------
// Returns {exclusive rank of this thread among flagged threads, number flagged in the block}.
// s_warp: 32 ints of shared memory.  Two __syncthreads per call; blockDim.x must be 1024.
__device__ __forceinline__ int2 block_rank_1024(bool flag, int* s_warp) {
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpos = __popc(m & ((1u << lane) - 1u));
  __syncthreads();
  if (lane == 0) s_warp[warp] = __popc(m);
  __syncthreads();
  const int v = s_warp[lane];
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  const int wbase = __shfl_sync(0xffffffffu, incl - v, warp);
  return make_int2(wbase + wpos, total);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Online log-sum-exp accumulator: running max m and sum s of exp(L - m).
// Terms more than kLseSkip below the running max are dropped: each is < e^-kLseSkip of the
// largest term, so with K <= 1e9 kernels the relative change of the sum is < 6e-10 * e^-... see
// DESIGN.md "log-sum-exp truncation" (bound: K * e^-46 = 1e5 * 1.05e-20 ~ 1e-15).
constexpr double kLseSkip = 46.0;  // generic kernel; the fast kernel gets ln(K) + 30 from the host
__device__ __forceinline__ void lse_push(double L, double& m, double& s) {
  if (L > m) {
    s = s * exp(m - L) + 1.0;
    m = L;
  } else if (L - m > -kLseSkip) {
    s += exp(L - m);
  } else if (L != L) {
    m = L;  // NaN poisons the row (np.max propagates NaN)
    s = L;
  }
}
// merge two (m, s) accumulators
__device__ __forceinline__ void lse_merge(double m2, double s2, double& m, double& s) {
  if (m2 != m2 || m != m) {
    m = m + m2;
    s = m;
    return;
  }
  if (m2 == -INFINITY) return;
  if (m == -INFINITY) {
    m = m2;
    s = s2;
    return;
  }
  if (m2 > m) {
    s = s * exp(m - m2) + s2;
    m = m2;
  } else {
    s += s2 * exp(m2 - m);
  }
}

}  // namespace tpe
