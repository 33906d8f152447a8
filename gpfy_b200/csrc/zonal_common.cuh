// Shared pieces of the zonal kernels: t = Vhat xhat on the FP64 tensor cores and the monic Gegenbauer
// recurrence evaluated on the accumulator fragments.
//
// Why tensor cores for a K = dim (9..36) contraction: with one point per lane every row of Vhat has to be
// broadcast from shared memory to all 32 lanes (6 LDS.128 = 3 KB of register-file traffic per row and warp), and
// at 128 B/clk that broadcast, not the FP64 pipe, bounded the kernels (measured: 9..13 k cycles per 32-point tile
// for ~4 k cycles of FP64 work).  As DMMA operands a row tile of Vhat costs 3 LDS.64 per lane and the point
// coordinates live in registers as B fragments for the whole tile.
//
// m8n8k4 fragment layout (lane = 4 g + t4):  A[g][t4],  B[t4][g],  C[g][2 t4 + {0, 1}].
//   T[row0 + g][8 nt + 2 t4 + j] = sum_ks sum_t4  Vhat[row0 + g][4 ks + t4] * Xhat[8 nt + g'][4 ks + t4]
#pragma once
#include "common.cuh"

namespace gpfy {

// Monic three-term recurrence of the Gegenbauer family with parameter lam:
//   p_0 = 1, p_1 = t, p_k = t p_{k-1} - beta[k] p_{k-2},   C_k^lam(t) = lead[k] p_k(t)
// (same polynomials as gegenbauer.py:47-60; two FP64 operations per step instead of three).
struct MonicTab {
  double beta[GPFY_MAX_DEGREES + 2];
  double lead[GPFY_MAX_DEGREES + 2];
  // d/dt C_l^lam = 2 sum_{j = l-1, l-3, ... >= 0} (j + lam) C_j^lam = sum_j dco[j] p_j   (equivalent to 2 lam C_{l-1}^{lam+1},
  // gegenbauer.py:73-81, but on the polynomials the lam-recurrence passes through anyway; all coefficients positive)
  double dco[GPFY_MAX_DEGREES + 2];
};

static inline MonicTab make_monic(double lam) {
  MonicTab m;
  m.lead[0] = 1.0;
  m.lead[1] = 2.0 * lam;
  m.beta[0] = m.beta[1] = 0.0;
  for (int k = 2; k < GPFY_MAX_DEGREES + 2; ++k) {
    const double d1 = 2.0 * (k + lam - 1.0) / k, d2 = (k + 2.0 * lam - 2.0) / k;
    m.lead[k] = d1 * m.lead[k - 1];
    m.beta[k] = m.lead[k] != 0.0 ? d2 * m.lead[k - 2] / m.lead[k] : 0.0;
  }
  for (int k = 0; k < GPFY_MAX_DEGREES + 2; ++k) m.dco[k] = 2.0 * (k + lam) * m.lead[k];
  return m;
}

#ifdef __CUDACC__
// All lanes evaluate degree l >= 1 (warp-uniform): p = p_l, p1 = p_{l-1}, p2 = p_{l-2}.
// The loop advances two degrees per trip in place, (p_{2j}, p_{2j+1}) -> (p_{2j+2}, p_{2j+3}), so no register
// is ever copied; the last one or two steps are written out to keep p_{l-2}.
template <int NV>
__device__ __forceinline__ void monic_uniform(const double* beta, int l, const double (&t)[NV], double (&p)[NV], double (&p1)[NV],
                                              double (&p2)[NV]) {
  if (l == 1) {
#pragma unroll
    for (int u = 0; u < NV; ++u) { p2[u] = 0.0; p1[u] = 1.0; p[u] = t[u]; }
    return;
  }
  double a[NV], b[NV];
#pragma unroll
  for (int u = 0; u < NV; ++u) { a[u] = 1.0; b[u] = t[u]; }
  const int nd = (l - 2 - (l & 1)) >> 1;
  double b0 = beta[2], b1 = beta[3];   // coefficients are fetched one trip ahead of their use
#pragma unroll 1
  for (int j = 0; j < nd; ++j) {
    const double nb0 = beta[2 * j + 4], nb1 = beta[2 * j + 5];
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      a[u] = fma(t[u], b[u], -b0 * a[u]);
      b[u] = fma(t[u], a[u], -b1 * b[u]);
    }
    b0 = nb0;
    b1 = nb1;
  }
  if (l & 1) {   // after nd trips (b0, b1) = (beta[l - 1], beta[l])
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      p2[u] = b[u];
      p1[u] = fma(t[u], b[u], -b0 * a[u]);
      p[u] = fma(t[u], p1[u], -b1 * b[u]);
    }
  } else {       // (b0, b1) = (beta[l], beta[l + 1])
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      p2[u] = a[u];
      p1[u] = b[u];
      p[u] = fma(t[u], b[u], -b0 * a[u]);
    }
  }
}

// monic_uniform plus the derivative zp = d/dt C_l^lam(t) = sum over j = l-1, l-3, ... of dco[j] p_j (see MonicTab::dco).
template <int NV>
__device__ __forceinline__ void monic_uniform_d(const double* beta, const double* dco, int l, const double (&t)[NV], double (&p)[NV],
                                                double (&zp)[NV]) {
  if (l == 1) {
#pragma unroll
    for (int u = 0; u < NV; ++u) { p[u] = t[u]; zp[u] = dco[0]; }
    return;
  }
  const bool odd = l & 1;
  double a[NV], b[NV];
  {
    const double c0 = dco[0], c1 = dco[1];
#pragma unroll
    for (int u = 0; u < NV; ++u) { a[u] = 1.0; b[u] = t[u]; zp[u] = odd ? c0 : c1 * t[u]; }
  }
  const int nd = (l - 2 - (l & 1)) >> 1;
  double b0 = beta[2], b1 = beta[3];
#pragma unroll 1
  for (int j = 0; j < nd; ++j) {
    const double nb0 = beta[2 * j + 4], nb1 = beta[2 * j + 5];
    const double dc = dco[2 * j + 2 + (odd ? 0 : 1)];   // odd l collects the even polynomials (a), even l the odd ones (b)
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      a[u] = fma(t[u], b[u], -b0 * a[u]);
      b[u] = fma(t[u], a[u], -b1 * b[u]);
      zp[u] = fma(dc, odd ? a[u] : b[u], zp[u]);
    }
    b0 = nb0;
    b1 = nb1;
  }
  if (odd) {   // (b0, b1) = (beta[l - 1], beta[l]):  p_{l-1} joins the derivative sum
    const double dc = dco[l - 1];
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const double p1 = fma(t[u], b[u], -b0 * a[u]);
      p[u] = fma(t[u], p1, -b1 * b[u]);
      zp[u] = fma(dc, p1, zp[u]);
    }
  } else {     // (b0, b1) = (beta[l], beta[l + 1]);  b = p_{l-1} is already in the sum
#pragma unroll
    for (int u = 0; u < NV; ++u) p[u] = fma(t[u], b[u], -b0 * a[u]);
  }
}

// monic_mixed plus the derivative (per-lane degree l; l <= 0 gives zp = 0)
template <int NV>
__device__ __forceinline__ void monic_mixed_d(const double* beta, const double* dco, int l, int lmax, const double (&t)[NV],
                                              double (&p)[NV], double (&zp)[NV]) {
  double p1[NV];
  const double z0 = l >= 1 ? ((l & 1) ? dco[0] : 0.0) : 0.0, z1 = (l >= 2 && !(l & 1)) ? dco[1] : 0.0;
#pragma unroll
  for (int u = 0; u < NV; ++u) { p1[u] = 1.0; p[u] = t[u]; zp[u] = fma(z1, t[u], z0); }
#pragma unroll 1
  for (int k = 2; k <= lmax; ++k) {
    const double b = beta[k];
    const bool on = k <= l;
    const double dc = (k < l && !((l - 1 - k) & 1)) ? dco[k] : 0.0;
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const double pn = fma(t[u], p[u], -b * p1[u]);
      p1[u] = on ? p[u] : p1[u];
      p[u] = on ? pn : p[u];
      zp[u] = fma(dc, pn, zp[u]);
    }
  }
}

// Lanes hold rows of different degrees (a row tile that straddles a degree boundary): run to the largest degree
// of the tile and let every lane stop at its own.  For l < 1 the results are unused (callers special-case them).
template <int NV>
__device__ __forceinline__ void monic_mixed(const double* beta, int l, int lmax, const double (&t)[NV], double (&p)[NV],
                                            double (&p1)[NV], double (&p2)[NV]) {
#pragma unroll
  for (int u = 0; u < NV; ++u) { p2[u] = 0.0; p1[u] = 1.0; p[u] = t[u]; }
#pragma unroll 1
  for (int k = 2; k <= lmax; ++k) {
    const double b = beta[k];
    const bool on = k <= l;
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const double pn = fma(t[u], p[u], -b * p1[u]);
      p2[u] = on ? p1[u] : p2[u];
      p1[u] = on ? p[u] : p1[u];
      p[u] = on ? pn : p[u];
    }
  }
}
#endif

}  // namespace gpfy
