// k_logpdf_mma3 -- the fp64 tensor-core grid kernel with a WARP-COMPACTED exact tier.
//
// Same tiling, table layout, classification rule and accuracy bounds as k_logpdf_mma (tpe_kernels.cuh); what
// changes is who evaluates the exact ("near") terms.  ncu on k_logpdf_mma at config 2 (round 2,
// profiles/r2_ncu_mma_default.txt): only 1-2 % of the C x K terms are near, but a near term in ONE lane makes the
// whole warp run the parking cascade, and a lane whose 4-deep buffer fills makes the whole warp evaluate five fp64
// exponentials -- 57 % of the kernel's 563 M warp instructions, and fp64 instructions that compete with DMMA for
// the same pipe (fp64 pipe 11.7 % + DMMA 64.6 %).
// Here the lanes of a warp append their near terms (value, candidate row) to a QUEUE in shared memory (ballot +
// popc, no divergence); as soon as 32 entries are queued, every lane takes ONE of them: 32 useful exponentials per
// warp instruction instead of ~2.  The per-candidate sums live in shared memory relative to a fixed reference
// (fp64 does not care about the scale; the reference moves only if a term exceeds it by 300 nats), updated by a
// fixed-order segmented butterfly, so the result does not depend on timing any more than before (the thresholds
// still follow the maxima other CTAs publish).  The fp32 far tier is unchanged; its per-lane sum is re-based
// whenever the lane adopts a higher reference max.
#pragma once
#include "tpe_kernels.cuh"

namespace tpe {

constexpr int kQCap = 64;                                   // queue entries per warp
constexpr int kQWarpBytes = kQCap * 8 + kQCap + 8 * 8 * 3;  // qL, qg, sref, ssum, sbase

// e^x for -700 < x < 700 (LseTier::exp_neg without the clamp at 0): ~20 instructions
__device__ __forceinline__ double exp_any(double x) {
  const double xc = fmin(fmax(x, -700.0), 700.0);
  const double t = fma(xc, 1.4426950408889634074, 6755399441055744.0);
  const int n = __double2loint(t);
  const double nf = t - 6755399441055744.0;
  double r = fma(nf, -6.93147180369123816490e-01, xc);
  r = fma(nf, -1.90821492927058770002e-10, r);
  double p = 2.08767569878680989792e-09;
  p = fma(p, r, 2.50521083854417187751e-08);
  p = fma(p, r, 2.75573192239858906526e-07);
  p = fma(p, r, 2.75573192239858906526e-06);
  p = fma(p, r, 2.48015873015873015873e-05);
  p = fma(p, r, 1.98412698412698412698e-04);
  p = fma(p, r, 1.38888888888888888889e-03);
  p = fma(p, r, 8.33333333333333333333e-03);
  p = fma(p, r, 4.16666666666666666667e-02);
  p = fma(p, r, 1.66666666666666666667e-01);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double y = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
  return (x < -700.0) ? 0.0 : y;
}

template <int PB, int KG, int NT, int TK, int ST, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_logpdf_mma3(const double* __restrict__ tabm, const double* __restrict__ ckk, int64_t Kfp,
              const double2* __restrict__ colprm, const double* __restrict__ xT, int64_t ct_stride, int64_t kps,
              double lse_skip, double2* __restrict__ part, unsigned long long* __restrict__ gmax, double lse_near) {
  static_assert(PB % 8 == 0 && TK % (8 * KG) == 0, "bad tiling");
  constexpr int NI = PB / 4;        // k-steps of the mma chain
  constexpr int NI2 = NI / 2;
  constexpr int V = 2 * KG;         // values per lane and step
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* tiles = reinterpret_cast<double*>(smem_raw);                             // ST * TK * PB
  double* csts = tiles + (size_t)ST * TK * PB;                                     // ST * TK
  uint64_t* full = reinterpret_cast<uint64_t*>(csts + (size_t)ST * TK);            // ST
  uint64_t* empty = full + ST;                                                     // ST
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  // this warp's exact-tier state
  unsigned char* wsm = reinterpret_cast<unsigned char*>(empty + ST) + (size_t)warp * kQWarpBytes;
  double* qL = reinterpret_cast<double*>(wsm);                                     // [kQCap]
  double* sref = qL + kQCap;                                                       // [8] reference of the exact sums
  double* ssum = sref + 8;                                                         // [8] sum of e^(L - sref)
  unsigned long long* sbase = reinterpret_cast<unsigned long long*>(ssum + 8);     // [8] best max known (ordered bits)
  unsigned char* qg = reinterpret_cast<unsigned char*>(sbase + 8);                 // [kQCap]
  const int64_t k0 = blockIdx.y * kps;
  const int64_t k1 = (k0 + kps < Kfp) ? k0 + kps : Kfp;
  const int ntiles = (k1 > k0) ? (int)((k1 - k0 + TK - 1) / TK) : 0;
  const int64_t wbase = (int64_t)blockIdx.x * ((NT / 32) * 8) + (int64_t)warp * 8;

  if (tid == 0) {
    for (int s = 0; s < ST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NT / 32);
    }
    mbar_fence_init();
  }
  if (lane < 8) {
    sref[lane] = -INFINITY;
    ssum[lane] = 0.0;
    sbase[lane] = 0ull;   // from_order_bits(0) = -inf
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

  // A fragment: a[i] = scaled coordinate 4 i + q of candidate g of this warp
  double a[NI];
  double ha = 0.0;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int slot = 4 * i + q;
    const double2 cp = colprm[slot];
    const double v = (xT[(int64_t)slot * ct_stride + wbase + g] - cp.x) * cp.y;
    a[i] = v;
    ha = fma(v, v, ha);
  }
  ha += __shfl_xor_sync(0xffffffffu, ha, 1);
  ha += __shfl_xor_sync(0xffffffffu, ha, 2);
  ha *= -0.5;

  auto chain = [&](const double* tile, const double* ctile, int kg, double (&d0)[KG], double (&d1)[KG]) {
    const double2* fb = reinterpret_cast<const double2*>(tile + (size_t)kg * 8 * PB) + lane;
#pragma unroll
    for (int u = 0; u < KG; ++u) {
      const double2 cc = reinterpret_cast<const double2*>(ctile + (kg + u) * 8)[q];
      d0[u] = cc.x;
      d1[u] = cc.y;
    }
    double2 v[2][KG];
#pragma unroll
    for (int u = 0; u < KG; ++u) v[0][u] = fb[u * (4 * PB)];
#pragma unroll
    for (int i2 = 0; i2 < NI2; ++i2) {
      if (i2 + 1 < NI2) {
#pragma unroll
        for (int u = 0; u < KG; ++u) v[(i2 + 1) & 1][u] = fb[u * (4 * PB) + (i2 + 1) * 32];
      }
#pragma unroll
      for (int u = 0; u < KG; ++u) dmma_8x8x4(d0[u], d1[u], a[2 * i2], v[i2 & 1][u].x);
#pragma unroll
      for (int u = 0; u < KG; ++u) dmma_8x8x4(d0[u], d1[u], a[2 * i2 + 1], v[i2 & 1][u].y);
    }
  };

  unsigned long long* slot = gmax + wbase + g;
  // ---- seed: the maximum over this CTA's first tile (rounded down in fp32), shared through gmax ----
  if (ntiles > 0) {
    mbar_wait(&full[0], 0u);   // tile 0 stays in its stage: the main loop processes it again
    const int tk = (int)((k1 - k0 < TK) ? (k1 - k0) : TK);
    float mx = -INFINITY;
    for (int kg = 0; kg < tk / 8; kg += KG) {
      double d0[KG], d1[KG];
      chain(tiles, csts, kg, d0, d1);
#pragma unroll
      for (int u = 0; u < KG; ++u) mx = fmaxf(mx, fmaxf(__double2float_rd(d0[u]), __double2float_rd(d1[u])));
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    if (q == 0 && mx > -INFINITY) atomicMax(slot, static_cast<unsigned long long>(order_bits((double)mx)));
    __syncthreads();           // every warp of the CTA has published; other CTAs' values arrive as they come
    if (q == 0) sbase[g] = *reinterpret_cast<volatile unsigned long long*>(slot);
    __syncwarp();
  }

  double base = from_order_bits(sbase[g]);   // reference of the classification and of the far sum (a lower bound of the max)
  double fsum = 0.0;                         // fp32-tier mass relative to `base`
  float ffar = 0.0f;
  int qcount = 0;                            // uniform over the warp
  const float lim_skip = (float)lse_skip, lim_near = (float)lse_near;

  auto rebase = [&](double nb) {             // nb >= base; the fp32-tier sum follows the reference
    if (nb > base) {
      const float sc = ex2_approx(__double2float_ru(base - nb) * 1.44269504f);   // base = -inf: 0 (and the sum is 0)
      fsum = (fsum + (double)ffar) * (double)sc;
      ffar = 0.0f;
      base = nb;
    }
  };
  // the 32 oldest queue entries (or all of them at the end): one per lane
  auto drain = [&](int n) {
    const bool have = lane < n;
    const double L = have ? qL[lane] : -INFINITY;
    const int eg = have ? (int)qg[lane] : -1;
    // references not set yet: the first queued term of that candidate
    unsigned todo = __ballot_sync(0xffffffffu, have && sref[eg < 0 ? 0 : eg] == -INFINITY);
    while (todo) {
      const int src = __ffs(todo) - 1;
      const int gg = __shfl_sync(0xffffffffu, eg, src);
      const double Ls = __shfl_sync(0xffffffffu, L, src);
      if (lane == 0) sref[gg] = Ls;
      todo &= ~__ballot_sync(0xffffffffu, eg == gg);
    }
    __syncwarp();
    double x = have ? L - sref[eg < 0 ? 0 : eg] : -INFINITY;
    if (__any_sync(0xffffffffu, x > 300.0)) {   // a reference 300 nats too low: move it up (sums follow)
      for (int gg = 0; gg < 8; ++gg) {
        double mxl = (eg == gg) ? L : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double t = __shfl_xor_sync(0xffffffffu, mxl, o);
          mxl = t > mxl ? t : mxl;
        }
        if (lane == 0 && mxl > sref[gg] + 300.0) {
          ssum[gg] *= exp_any(sref[gg] - mxl);
          sref[gg] = mxl;
        }
      }
      __syncwarp();
      x = have ? L - sref[eg < 0 ? 0 : eg] : -INFINITY;
    }
    const double e = exp_any(x);
    if (have && L > from_order_bits(sbase[eg])) atomicMax(&sbase[eg], static_cast<unsigned long long>(order_bits(L)));
    unsigned groups = __ballot_sync(0xffffffffu, have);
    // fixed-order segmented sum: candidates in ascending order, each by the same butterfly
    for (int gg = 0; gg < 8; ++gg) {
      if (!__any_sync(0xffffffffu, eg == gg)) continue;
      double v = (eg == gg) ? e : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) ssum[gg] += v;
    }
    (void)groups;
    // shift the rest of the queue down
    const int rest = qcount - n;
    double tL = 0.0;
    unsigned char tg = 0;
    if (lane < rest) { tL = qL[n + lane]; tg = qg[n + lane]; }
    __syncwarp();
    if (lane < rest) { qL[lane] = tL; qg[lane] = tg; }
    qcount = rest;
    __syncwarp();
  };

  for (int t = 0; t < ntiles; ++t) {
    const int st = t % ST;
    if (tid == 0 && t + ST - 1 < ntiles) {
      if (t > 0) mbar_wait(&empty[(t - 1) % ST], (uint32_t)(((t - 1) / ST) & 1));
      issue(t + ST - 1);
    }
    mbar_wait(&full[st], (uint32_t)((t / ST) & 1));
    // tile boundary: publish / adopt the best max of this candidate, re-base the fp32 tier
    {
      if (q == 0) {
        const unsigned long long mine = sbase[g];
        unsigned long long seen = *reinterpret_cast<volatile unsigned long long*>(slot);
        if (mine > seen) { atomicMax(slot, mine); seen = mine; }
        sbase[g] = seen;
      }
      __syncwarp();
      const double nb = from_order_bits(sbase[g]);
      fsum += (double)ffar;   // keeps the fp32 runs short (<= one tile)
      ffar = 0.0f;
      rebase(nb);
    }
    const int64_t ks = k0 + (int64_t)t * TK;
    const int tk = (int)((k1 - ks < TK) ? (k1 - ks) : TK);
    const double* tile = tiles + (size_t)st * TK * PB;
    const double* ctile = csts + (size_t)st * TK;
    for (int kg = 0; kg < tk / 8; kg += KG) {
      if (kg != 0 && (kg & 15) == 0) {   // long tiles (small PB): fp32 runs of <= 32 terms
        fsum += (double)ffar;
        ffar = 0.0f;
      }
      double d0[KG], d1[KG];
      chain(tile, ctile, kg, d0, d1);
      double vals[V];
#pragma unroll
      for (int u = 0; u < KG; ++u) {
        vals[2 * u] = d0[u];
        vals[2 * u + 1] = d1[u];
      }
      bool near[V];
      float add = 0.0f;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float df = __double2float_rn(vals[i] - base);  // base = -inf -> +inf -> near
        near[i] = df > -lim_near;
        const bool far = !near[i] && df > -lim_skip;
        const float e = ex2_approx(df * 1.44269504f);
        add += far ? e : 0.0f;
      }
      ffar += add;
      bool drained = false;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const unsigned mask = __ballot_sync(0xffffffffu, near[i]);
        if (mask) {
          if (near[i]) {
            const int pos = qcount + __popc(mask & ((1u << lane) - 1u));
            qL[pos] = vals[i];
            qg[pos] = (unsigned char)g;
          }
          qcount += __popc(mask);
          __syncwarp();
          if (qcount >= 32) { drain(32); drained = true; }
        }
      }
      if (drained) {
        // adopt a higher reference only when it pays (a stale base is still a valid lower bound)
        const double nb = from_order_bits(sbase[g]);
        if (__any_sync(0xffffffffu, nb > base + 1.0)) rebase(nb > base + 1.0 ? nb : base);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
  }
  if (qcount > 0) drain(qcount);
  __syncwarp();
  // ---- per candidate: exact part (sref, ssum) + the four lanes' fp32-tier sums (relative to their base) ----
  {
    const double r = sref[g];
    const double mxk = from_order_bits(sbase[g]);           // the largest term any lane / CTA has seen
    const double ref = (r > -INFINITY) ? r : base;          // no exact term in this slice: the far sums' own reference
    double far_part = (fsum + (double)ffar);
    far_part = (far_part > 0.0) ? far_part * exp_any(base - ref) : 0.0;
    far_part += __shfl_xor_sync(0xffffffffu, far_part, 1);
    far_part += __shfl_xor_sync(0xffffffffu, far_part, 2);
    double ss = ((r > -INFINITY) ? ssum[g] : 0.0) + far_part;
    double mm = ref;
    if (!(ref > -INFINITY)) { mm = -INFINITY; ss = 0.0; }
    else if (mxk > -INFINITY && mxk != ref) {               // report relative to the max (keeps s ~ O(1) for the merge)
      ss *= exp_any(ref - mxk);
      mm = mxk;
    }
    if (q == 0) part[blockIdx.y * ct_stride + wbase + g] = make_double2(mm + ha, ss);
  }
}

}  // namespace tpe
