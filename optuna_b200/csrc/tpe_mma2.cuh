// k_logpdf_mma2 -- round-2 revision of the fp64 tensor-core grid kernel (k_logpdf_mma, tpe_kernels.cuh): same
// tiling, same table layout, same two-tier log-sum-exp and the same accuracy bound; what changes is WHEN the
// log-sum-exp work is done.
//
//   SEED  The exact ("near") tier fired on ~90 % of the 128-value batches because every lane starts with
//         base = -inf and its running max converges slowly (round-1 ncu: 0.27 of 1.17 ms).  Now every CTA first
//         runs the DMMA chain over its FIRST tile for the maximum only (fp32 max of values rounded DOWN: a
//         rigorous lower bound of the true max), publishes it (atomicMax on the candidate's slot, shared by the
//         k-splits and the 4 lanes of a candidate) and starts the real pass with base = the best value any CTA
//         has seen: ~1100 kernels instead of none.  Cost: one extra tile per CTA (~1 %).
//   PIPE  Software pipelining inside the warp: the classification of the PREVIOUS step's values (cvt, compare,
//         ex2, add -- fp32 / SFU / ALU pipes) is interleaved with the DMMA chain of the current step, one value
//         per pair of k-steps, so the tensor pipe is fed while the log-sum-exp instructions issue; before, all
//         warps of an SM sub-partition ran their DMMA phases and their classification phases in lockstep.
#pragma once
#include "tpe_kernels.cuh"

namespace tpe {

// ------------------------------------------------------------------------------------------------------------
// Budgeted two-tier log-sum-exp (OPT bit 2).  LseTier sends a term to the cheap fp32 tier only when it lies more
// than ln K + 17.5 below the reference max -- a bound that must hold if ALL K terms sat just under that line.  Real
// mixtures are nowhere near that case: at config 2 the terms spread with a standard deviation of ~10 nats, 8 % of
// them fall inside those 29 nats and every 128-value batch of a warp carries ~10 exact folds (ncu, round 2: the
// exact path executes 57 % of the kernel's 563 M warp instructions).
// Here the guarantee is kept but spent adaptively: every lane owns a BUDGET of fp32-tier mass, B / (4 nsplit)
// relative to the reference max (4 lanes and nsplit k-splits share a candidate, B = 2e-7).  A term goes to the fp32
// tier iff its own value e^(L - base) still fits in what is left of the budget -- otherwise it is folded exactly.
// The fp32 tier is also made ~10x more accurate than LseTier's (fp64 range reduction: t = (L - base) log2 e,
// n = rint(t), f = t - n in fp64; 2^f by MUFU.EX2 on |f| <= 1/2; 2^n by an exponent add; the <= 8 terms of a step
// are summed in fp32 and the step sum goes to an fp64 accumulator): <= 1e-6 relative per term (3e-8 rounding of f,
// 2.4e-7 ex2.approx, 4.8e-7 fp32 adds), so the whole tier contributes <= 1e-6 * 2e-7 = 2e-13 of the sum, whatever
// the data.  Nothing is dropped: a vanishing term costs a vanishing share of the budget.  When `base` rises the
// consumed budget is rescaled with it (rounded up).
// At config 2 the line between the tiers moves from 29 to ~21 nats below the max: 5x fewer exact folds.
// ------------------------------------------------------------------------------------------------------------
struct LseBudget : LseTier {
  float consumed, budget;
  __device__ __forceinline__ void init_budget(float b) { init(); consumed = 0.0f; budget = b; }
  // LseTier::flush moves base; the consumed mass is relative to base
  __device__ __forceinline__ void flush_b() {
    const double old = base;
    flush();
    const float dd = __double2float_ru(old - base);                       // <= 0; -inf when nothing was folded before
    consumed = (old > -INFINITY) ? consumed * ex2_approx(dd * 1.44269504f) * 1.00001f : 0.0f;
  }
  __device__ __forceinline__ void sync_global_b(unsigned long long* slot) {
    if (m > gm) atomicMax(slot, static_cast<unsigned long long>(order_bits(m)));
    const double seen = from_order_bits(*reinterpret_cast<volatile unsigned long long*>(slot));
    gm = fmax(gm, seen);
    if (__any_sync(0xffffffffu, gm > base)) flush_b();
  }
  __device__ __forceinline__ void park_b(double L, bool near) {
    b3 = near ? b2 : b3;
    b2 = near ? b1 : b2;
    b1 = near ? b0 : b1;
    b0 = near ? L : b0;
    cnt += near ? 1 : 0;
    if (__any_sync(0xffffffffu, cnt == 4)) flush_b();
  }
  template <int N>
  __device__ __forceinline__ void push(const double (&L)[N]) {
    bool near[N];
    bool any = false;
    float run = 0.0f;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double df = L[i] - base;                                       // base = -inf -> +inf -> exact
      const double t = fmax(df * 1.4426950408889634074, -120.0);
      const double tn = t + 6755399441055744.0;                            // rint(t) in the low word
      const double f = t - (tn - 6755399441055744.0);                      // |f| <= 1/2
      const int n = __double2loint(tn);
      const float e = __int_as_float(__float_as_int(ex2_approx(__double2float_rn(f))) + (n << 23));
      const float next = run + e;
      const bool far = (df <= 0.0) && (consumed + next <= budget);         // NaN and df > 0 compare false: exact
      run = far ? next : run;
      near[i] = !far;
      any = any || near[i];
    }
    consumed += run;
    fsum += (double)run;
    if (__any_sync(0xffffffffu, any)) {
#pragma unroll
      for (int i = 0; i < N; ++i)
        if (__any_sync(0xffffffffu, near[i])) park_b(L[i], near[i]);
    }
  }
};

// ------------------------------------------------------------------------------------------------------------
// Lane-private parking lot in local memory (OPT bit 3).  ncu, round 2: only 1-2 % of the terms are "near" (exact
// tier) at config 2, yet the exact path executes 57 % of the kernel's warp instructions -- LseTier parks a near term
// in a 4-deep REGISTER buffer, which costs the whole warp a vote and an 8-instruction shift cascade per value slot
// that holds a near term in ANY lane (75-90 % of the 128-value batches), and a 5-exp flush whenever any lane's
// buffer fills.  Here a lane stores its near terms into a 16-deep buffer it indexes dynamically (the compiler
// places it in local memory: one predicated STL + IADD per value, no vote) and the warp folds when a lane reaches
// 12: ~10x fewer flushes, each near term is exponentiated exactly once.  Same classification, same bounds.
// ------------------------------------------------------------------------------------------------------------
struct LseLocal : LseTier {
  static constexpr int kDepth = 16, kFlushAt = 12;   // a step adds at most 2 * KG <= 4 terms per lane... see push
  double buf[kDepth];
  __device__ __forceinline__ void flush_l() {
    int mc = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mc = max(mc, __shfl_xor_sync(0xffffffffu, mc, o));
    double nm = m;
    for (int j = 0; j < mc; ++j) {
      const double v = (j < cnt) ? buf[j] : -INFINITY;
      nm = (v > nm) ? v : nm;
    }
    double t = s * exp_neg(m - nm);   // m = -inf: s = 0 and exp_neg(NaN) is finite -> 0
    for (int j = 0; j < mc; ++j) {
      const double v = (j < cnt) ? buf[j] : -INFINITY;
      t += exp_neg(v - nm);           // -inf -> 0
    }
    s = t;
    m = nm;
    cnt = 0;
    flush();                          // nothing parked in registers: folds the fp32 tier and moves base
  }
  __device__ __forceinline__ void sync_global_l(unsigned long long* slot) {
    // the maximum a lane could publish is the larger of its folded max and its parked terms; publishing the
    // folded one is enough for the others' thresholds (it is some kernel's L) and costs nothing
    if (m > gm) atomicMax(slot, static_cast<unsigned long long>(order_bits(m)));
    const double seen = from_order_bits(*reinterpret_cast<volatile unsigned long long*>(slot));
    gm = fmax(gm, seen);
    if (__any_sync(0xffffffffu, gm > base)) flush_l();
  }
  template <int N>
  __device__ __forceinline__ void push(const double (&L)[N], float skip, float tnear) {
    static_assert(kFlushAt + N <= kDepth + 1, "a step may not overflow the buffer");
    float add = 0.0f;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const float df = __double2float_rn(L[i] - base);  // base = -inf -> +inf -> near
      const bool near = df > -tnear;
      const bool far = !near && df > -skip;
      const float e = ex2_approx(df * 1.44269504f);
      add += far ? e : 0.0f;
      if (near) {
        buf[cnt] = L[i];
        ++cnt;
      }
    }
    ffar += add;
    if (__any_sync(0xffffffffu, cnt >= kFlushAt)) flush_l();
  }
};

template <int PB, int M, int KG, int NT, int TK, int ST, int MINB, int OPT>
__global__ void __launch_bounds__(NT, MINB)
k_logpdf_mma2(const double* __restrict__ tabm, const double* __restrict__ ckk, int64_t Kfp,
              const double2* __restrict__ colprm, const double* __restrict__ xT, int64_t ct_stride, int64_t kps,
              double lse_skip, double2* __restrict__ part, unsigned long long* __restrict__ gmax, double lse_near) {
  static_assert(PB % 8 == 0 && TK % (8 * KG) == 0, "bad tiling");
  constexpr bool SEED = (OPT & 1) != 0, PIPE = (OPT & 2) != 0, BUDGET = (OPT & 4) != 0, LBUF = (OPT & 8) != 0;
  static_assert(!(PIPE && BUDGET) && !(LBUF && (PIPE || BUDGET)), "pick one");
  constexpr int NI = PB / 4;        // k-steps of the mma chain
  constexpr int NI2 = NI / 2;       // pairs of k-steps (one LDS.128 each)
  constexpr int CW = 8 * M;         // candidates per warp
  constexpr int V = 2 * KG;         // values per lane, candidate group and step
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* tiles = reinterpret_cast<double*>(smem_raw);                             // ST * TK * PB
  double* csts = tiles + (size_t)ST * TK * PB;                                     // ST * TK
  uint64_t* full = reinterpret_cast<uint64_t*>(csts + (size_t)ST * TK);            // ST
  uint64_t* empty = full + ST;                                                     // ST
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  const int64_t k0 = blockIdx.y * kps;
  const int64_t k1 = (k0 + kps < Kfp) ? k0 + kps : Kfp;   // Kfp, kps: multiples of 8 * KG
  const int ntiles = (k1 > k0) ? (int)((k1 - k0 + TK - 1) / TK) : 0;
  const int64_t wbase = (int64_t)blockIdx.x * ((NT / 32) * CW) + (int64_t)(tid >> 5) * CW;

  if (tid == 0) {
    for (int s = 0; s < ST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NT / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](int t) {
    const int st = t % ST;
    const int64_t ks = k0 + (int64_t)t * TK;
    const int tk = (int)((k1 - ks < TK) ? (k1 - ks) : TK);
    const uint32_t b_tile = (uint32_t)((size_t)tk * PB * 8);
    const uint32_t b_cst = (uint32_t)(tk * 8);
    fence_proxy_async();
    mbar_expect_tx(&full[st], b_tile + b_cst);
    bulk_g2s(tiles + (size_t)st * TK * PB, tabm + ks * PB, b_tile, &full[st]);
    bulk_g2s(csts + (size_t)st * TK, ckk + ks, b_cst, &full[st]);
  };
  if (tid == 0) {
    for (int t = 0; t < ST - 1 && t < ntiles; ++t) issue(t);
  }

  // A fragments: a[m][i] = A[row g][col q] of k-step i = scaled coordinate 4 i + q of candidate 8 m + g
  double a[M][NI];
  double ha[M];
#pragma unroll
  for (int m = 0; m < M; ++m) ha[m] = 0.0;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int slot = 4 * i + q;
    const double2 cp = colprm[slot];
#pragma unroll
    for (int m = 0; m < M; ++m) {
      const int64_t ct = wbase + 8 * m + g;
      const double v = (xT[(int64_t)slot * ct_stride + ct] - cp.x) * cp.y;
      a[m][i] = v;
      ha[m] = fma(v, v, ha[m]);
    }
  }
#pragma unroll
  for (int m = 0; m < M; ++m) {
    ha[m] += __shfl_xor_sync(0xffffffffu, ha[m], 1);
    ha[m] += __shfl_xor_sync(0xffffffffu, ha[m], 2);
    ha[m] *= -0.5;
  }
  using Acc = std::conditional_t<BUDGET, LseBudget, std::conditional_t<LBUF, LseLocal, LseTier>>;
  Acc acc[M];
#pragma unroll
  for (int m = 0; m < M; ++m) {
    if constexpr (BUDGET) acc[m].init_budget((float)lse_near);   // the launcher passes the per-lane budget here
    else acc[m].init();
  }
  const float lim_skip = (float)lse_skip, lim_near = (float)lse_near;

  // one step = KG kernel groups x M candidate groups: the DMMA chains, optionally with `mid(i2)` between the
  // pairs of k-steps
  auto chain = [&](const double* tile, const double* ctile, int kg, double (&d0)[KG][M], double (&d1)[KG][M], auto&& mid) {
    const double2* fb = reinterpret_cast<const double2*>(tile + (size_t)kg * 8 * PB) + lane;
#pragma unroll
    for (int u = 0; u < KG; ++u) {
      const double2 cc = reinterpret_cast<const double2*>(ctile + (kg + u) * 8)[q];
#pragma unroll
      for (int m = 0; m < M; ++m) {
        d0[u][m] = cc.x;
        d1[u][m] = cc.y;
      }
    }
    double2 v[2][KG];
#pragma unroll
    for (int u = 0; u < KG; ++u) v[0][u] = fb[u * (4 * PB)];   // one kernel group = 8 * PB doubles
#pragma unroll
    for (int i2 = 0; i2 < NI2; ++i2) {
      if (i2 + 1 < NI2) {
#pragma unroll
        for (int u = 0; u < KG; ++u) v[(i2 + 1) & 1][u] = fb[u * (4 * PB) + (i2 + 1) * 32];
      }
#pragma unroll
      for (int u = 0; u < KG; ++u)
#pragma unroll
        for (int m = 0; m < M; ++m) dmma_8x8x4(d0[u][m], d1[u][m], a[m][2 * i2], v[i2 & 1][u].x);
      mid(i2);
#pragma unroll
      for (int u = 0; u < KG; ++u)
#pragma unroll
        for (int m = 0; m < M; ++m) dmma_8x8x4(d0[u][m], d1[u][m], a[m][2 * i2 + 1], v[i2 & 1][u].y);
    }
  };

  if constexpr (SEED) {
    if (ntiles > 0) {
      mbar_wait(&full[0], 0u);   // tile 0 stays in its stage: the main loop processes it again
      const int tk = (int)((k1 - k0 < TK) ? (k1 - k0) : TK);
      float mx[M];
#pragma unroll
      for (int m = 0; m < M; ++m) mx[m] = -INFINITY;
      for (int kg = 0; kg < tk / 8; kg += KG) {
        double d0[KG][M], d1[KG][M];
        chain(tiles, csts, kg, d0, d1, [](int) {});
#pragma unroll
        for (int u = 0; u < KG; ++u)
#pragma unroll
          for (int m = 0; m < M; ++m)   // rounded DOWN: the fp32 max stays <= some kernel's L
            mx[m] = fmaxf(mx[m], fmaxf(__double2float_rd(d0[u][m]), __double2float_rd(d1[u][m])));
      }
#pragma unroll
      for (int m = 0; m < M; ++m) {
        unsigned long long* slot = gmax + wbase + 8 * m + g;
        if (mx[m] > -INFINITY) atomicMax(slot, static_cast<unsigned long long>(order_bits((double)mx[m])));
      }
      __syncthreads();           // every warp of the CTA has published; other CTAs' values arrive as they come
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const double seen = from_order_bits(*reinterpret_cast<volatile unsigned long long*>(gmax + wbase + 8 * m + g));
        acc[m].gm = seen;
        acc[m].base = seen;      // m = -inf, s = 0: nothing folded yet, the first flush adopts `base`
      }
    }
  }

  // pending values of the previous step (PIPE)
  double pv[M][V] = {};
  bool have_prev = false;
  auto classify_one = [&](int m, int j, float& add, bool& nearflag) {
    const float df = __double2float_rn(pv[m][j] - acc[m].base);
    nearflag = df > -lim_near;
    const bool far = !nearflag && df > -lim_skip;
    const float e = ex2_approx(df * 1.44269504f);
    add += far ? e : 0.0f;
  };
  auto finish_prev = [&](float (&add)[M], bool (&nr)[M][V]) {
#pragma unroll
    for (int m = 0; m < M; ++m) {
      acc[m].ffar += add[m];
      bool any = false;
#pragma unroll
      for (int j = 0; j < V; ++j) any = any || nr[m][j];
      if (__any_sync(0xffffffffu, any)) {
#pragma unroll
        for (int j = 0; j < V; ++j)
          if (__any_sync(0xffffffffu, nr[m][j])) acc[m].park(pv[m][j], nr[m][j]);
      }
    }
  };

  for (int t = 0; t < ntiles; ++t) {
    const int st = t % ST;
    if (tid == 0 && t + ST - 1 < ntiles) {
      // the stage being refilled held tile t - 1: wait until every warp has released it
      if (t > 0) mbar_wait(&empty[(t - 1) % ST], (uint32_t)(((t - 1) / ST) & 1));
      issue(t + ST - 1);
    }
    mbar_wait(&full[st], (uint32_t)((t / ST) & 1));
    const int64_t ks = k0 + (int64_t)t * TK;
    const int tk = (int)((k1 - ks < TK) ? (k1 - ks) : TK);
    const double* tile = tiles + (size_t)st * TK * PB;
    const double* ctile = csts + (size_t)st * TK;
    for (int kg = 0; kg < tk / 8; kg += KG) {
      if (kg == 0 || (!SEED && t == 0)) {  // every tile (and, unseeded, every step of the CTA's first tile)
#pragma unroll
        for (int m = 0; m < M; ++m) {
          if constexpr (BUDGET) {
            acc[m].sync_global_b(gmax + wbase + 8 * m + g);
          } else if constexpr (LBUF) {
            acc[m].roll();
            acc[m].sync_global_l(gmax + wbase + 8 * m + g);
          } else {
            acc[m].roll();
            acc[m].sync_global(gmax + wbase + 8 * m + g);
          }
        }
      } else if (!BUDGET && (kg & 15) == 0) {  // long tiles (small PB): keep the fp32 runs at <= 32 terms
#pragma unroll
        for (int m = 0; m < M; ++m) acc[m].roll();
      }
      double d0[KG][M], d1[KG][M];
      if constexpr (PIPE) {
        float add[M];
        bool nr[M][V];
#pragma unroll
        for (int m = 0; m < M; ++m) {
          add[m] = 0.0f;
#pragma unroll
          for (int j = 0; j < V; ++j) nr[m][j] = false;
        }
        chain(tile, ctile, kg, d0, d1, [&](int i2) {
          if (have_prev) {
#pragma unroll
            for (int m = 0; m < M; ++m)
#pragma unroll
              for (int j = 0; j < V; ++j)
                if (j * NI2 / V == i2) classify_one(m, j, add[m], nr[m][j]);
          }
        });
        if (have_prev) finish_prev(add, nr);
#pragma unroll
        for (int m = 0; m < M; ++m)
#pragma unroll
          for (int u = 0; u < KG; ++u) {
            pv[m][2 * u] = d0[u][m];
            pv[m][2 * u + 1] = d1[u][m];
          }
        have_prev = true;
      } else {
        chain(tile, ctile, kg, d0, d1, [](int) {});
#pragma unroll
        for (int m = 0; m < M; ++m) {
          double vals[V];
#pragma unroll
          for (int u = 0; u < KG; ++u) {
            vals[2 * u] = d0[u][m];
            vals[2 * u + 1] = d1[u][m];
          }
          if constexpr (BUDGET) acc[m].template push<V>(vals);
          else if constexpr (LBUF) acc[m].template push<V>(vals, lim_skip, lim_near);
          else acc[m].template push_batch<V, true>(vals, lim_skip, lim_near);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);  // this warp is done with stage `st`
  }
  if constexpr (PIPE) {
    if (have_prev) {   // the last step's values
      float add[M];
      bool nr[M][V];
#pragma unroll
      for (int m = 0; m < M; ++m) {
        add[m] = 0.0f;
#pragma unroll
        for (int j = 0; j < V; ++j) classify_one(m, j, add[m], nr[m][j]);
      }
      finish_prev(add, nr);
    }
  }
#pragma unroll
  for (int m = 0; m < M; ++m) {
    if constexpr (LBUF) acc[m].flush_l();
    else acc[m].flush();
    double mm = acc[m].m, ss = acc[m].s;   // (LseBudget: the plain flush is enough at the end)
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      const double m2 = __shfl_xor_sync(0xffffffffu, mm, o), s2 = __shfl_xor_sync(0xffffffffu, ss, o);
      lse_merge(m2, s2, mm, ss);
    }
    if (q == 0) part[blockIdx.y * ct_stride + wbase + 8 * m + g] = make_double2(mm + ha[m], ss);
  }
}

}  // namespace tpe
