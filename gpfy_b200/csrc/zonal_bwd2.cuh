// K_B2  fused backward through the zonal map for P = 1, Mp <= 288, ambient dim <= 16 (the headline shapes).
//
// One persistent CTA per SM walks 32-point tiles of the chunk.  Per tile:
//   load   : the [Mp x 32] slice of C = W2 Z, the tile's xhat rows and per-point weights arrive in shared
//            memory through a double-buffered cp.async + mbarrier pipeline (issued one tile ahead)
//   phase 1: lane = point, warps take batches of 4 consecutive rows: t = Vhat xhat, the (alpha + 1)
//            recurrence gives BOTH d/dt C_l^alpha = 2 alpha C_{l-1}^{alpha+1} (gegenbauer.py:73-81) and, through
//            C_l^alpha = alpha / (l + alpha) (C_l^{alpha+1} - C_{l-2}^{alpha+1}), the zonal value z itself, so
//            psi = z . C (variational.py:326-327 folded) needs no extra pass;  dt = dz * zp overwrites C in place
//   phase 2: FP64 tensor cores on the shared-memory dT tile:
//            dVhat[rows, dim] += dT[rows, 32] Xhat[32, dim]      (VJP of spherical_harmonics.py:308; per-warp
//                                                                 accumulators persist over all tiles of the CTA)
//            dxhat[dim, 32]    = Vhat^T[dim, rows] dT[rows, 32]   (split over 6 row groups, summed in the tail)
//   tail   : one warp (rotating): likelihood terms (Gaussian, fused) and du -> dw / db partial sums
// dT never goes to HBM, the separate dVhat and likelihood kernels disappear, and the GEMM needs no psi epilogue.
#pragma once
#include "svgp_kernels.cuh"

namespace gpfy {

constexpr int ZB2_THREADS = 384, ZB2_WARPS = 12, ZB2_LD = 32, ZB2_XLD = 36, ZB2_KG = 6, ZB2_SUMS = 40;
// C / dT tile: [Mp][32] with the column XOR-swizzled by the row so that both the per-row accesses of phase 1
// and the 8 x 4 / 4 x 8 tensor-core fragment reads of phase 2 hit each 8-byte bank at most twice
__host__ __device__ __forceinline__ int zb2_swz(int row, int col) { return row * ZB2_LD + (col ^ ((row & 3) << 2)); }

struct ZonalBwd2Args {
  const double* C;       // [Mp][ldz]
  int64_t ldz;
  const double* wp;      // [ldz]  p r
  const double* cq;      // [ldz]  2 q r^2
  const double* dr;      // [ldz]  direct d/dr terms (non-fused likelihood only)
  const double* XH;      // [npts][DIMP]
  const double* R;       // [npts]
  const double* mz;      // [ldz]  mu2 . z       (fused likelihood)
  const double* Y;       // [npts]               (fused likelihood)
  const double* vptr;    // [1] kernel variance  (fused likelihood)
  const double* s2ptr;   // [1] noise variance   (fused likelihood)
  int order;
  const double* vhat_p;  // [Mp][DIMP], mu2 [Mp] right behind it
  int Mp, M, d;
  int64_t npts;
  double* partials;      // [gridDim][npart]
  int npart;
  double* dVpart;        // [gridDim][Mp][CT * 8]
  DegreeTable deg;
  RecCoef rc;
  double alpha;
};

static inline int zb2_smem_bytes(int Mp, int dimp) {
  const int hdr = 64 + 32 * dimp;  // wp | cq | xhat rows
  int dbl = Mp * dimp + Mp + 2 * Mp * ZB2_LD + 2 * hdr + 16 * ZB2_XLD + ZB2_KG * 16 * 32 + 2 * ZB2_WARPS * 32 + ZB2_WARPS * ZB2_SUMS +
            GPFY_MAX_DEGREES + 8;
  return dbl * 8 + Mp * 4 + 64;
}

// (z, zp) = (C_l^alpha(t), d/dt C_l^alpha(t)) from the alpha + 1 recurrence; l < 0 marks a padding row
__device__ __forceinline__ void geg_pair(const RecCoef& rc, const double* zr, int l, double t, double& z, double& zp) {
  if (l <= 0) {
    z = (l == 0) ? 1.0 : 0.0;
    zp = 0.0;
    return;
  }
  double Dm2 = 0.0, Dm1 = 1.0, D = rc.two_alpha_p1 * t;
  for (int k = 2; k <= l; ++k) {
    const double Dn = fma(rc.d1[k] * t, D, -rc.d2[k] * Dm1);
    Dm2 = Dm1;
    Dm1 = D;
    D = Dn;
  }
  zp = rc.two_alpha * Dm1;
  z = zr[l] * (D - Dm2);
}

template <int DIMP, bool FUSE_LIK>
__global__ void __launch_bounds__(ZB2_THREADS, 1) zonal_bwd2_kernel(ZonalBwd2Args a) {
  constexpr int CT = (DIMP + 7) / 8;
  constexpr int HDR = 64 + 32 * DIMP;
  extern __shared__ __align__(16) double smem[];
  __shared__ uint64_t bar_stage, bar_full[2];
  const int Mp = a.Mp;
  double* vs = smem;                          // [Mp][DIMP]
  double* mu2s = vs + Mp * DIMP;              // [Mp]
  double* cbuf = mu2s + Mp;                   // [2][Mp][36]
  double* hdr = cbuf + 2 * Mp * ZB2_LD;       // [2][wp(32) | cq(32) | xhat(32 x DIMP)]
  double* xsT = hdr + 2 * HDR;                // [16][XLD]  xhat^T (zero rows above DIMP)
  double* red = xsT + 16 * ZB2_XLD;            // [6][16][32]
  double* psir = red + ZB2_KG * 16 * 32;      // [2][12][32]
  double* sums = psir + 2 * ZB2_WARPS * 32;   // [12][40]
  double* zr = sums + ZB2_WARPS * ZB2_SUMS;         // [MAX_DEGREES + 8]
  int* rowdeg = reinterpret_cast<int*>(zr + GPFY_MAX_DEGREES + 8);  // [Mp]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int64_t npts_e = (a.npts + 1) & ~(int64_t)1;
  const int tiles = (int)((a.npts + 31) / 32);
  const int n_my = ((int)blockIdx.x < tiles) ? (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  // ---- one-time setup ----
  for (int i = tid; i < Mp; i += ZB2_THREADS) {
    int l = -1;
    if (i < a.M) {
      int o = 0;
      l = 0;
      while (i >= o + a.deg.m[l]) { o += a.deg.m[l]; ++l; }
    }
    rowdeg[i] = l;
  }
  for (int i = tid; i < GPFY_MAX_DEGREES + 8; i += ZB2_THREADS) zr[i] = a.alpha / ((double)i + a.alpha + (a.alpha == 0.0 && i == 0 ? 1.0 : 0.0));
  for (int i = tid; i < 16 * ZB2_XLD; i += ZB2_THREADS) xsT[i] = 0.0;
  for (int i = tid; i < ZB2_WARPS * ZB2_SUMS; i += ZB2_THREADS) sums[i] = 0.0;
  if (tid == 0) {
    mbar_init(&bar_full[0], ZB2_THREADS);
    mbar_init(&bar_full[1], ZB2_THREADS);
  }
  stage_params_tma(vs, a.vhat_p, (unsigned)((Mp * DIMP + Mp) * 8), &bar_stage);  // includes a __syncthreads

  // producer: this thread's share of a tile (Mp x 16 chunks of 16 bytes) + header chunks
  auto issue_tile = [&](int it) {
    const int buf = it & 1;
    const int64_t p0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * 32;
    double* cb = cbuf + buf * Mp * ZB2_LD;
    for (int q = tid; q < Mp * 16; q += ZB2_THREADS) {
      const int row = q >> 4, cc = (q & 15) * 2;
      const bool ok = (p0 + cc) < npts_e;
      cp_async16(cb + zb2_swz(row, cc), a.C + (int64_t)row * a.ldz + (ok ? p0 + cc : 0), ok);
    }
    double* hb = hdr + buf * HDR;
    if (tid < 32) {
      const int which = tid >> 4, cc = (tid & 15) * 2;
      const bool ok = (p0 + cc) < npts_e;
      cp_async16(hb + which * 32 + cc, (which ? a.cq : a.wp) + (ok ? p0 + cc : 0), ok);
    } else if (tid < 32 + 16 * DIMP) {
      const int q = tid - 32;  // 16-byte chunk of the contiguous [32][DIMP] block of xhat rows
      const int64_t e = p0 * DIMP + q * 2;
      const bool ok = e < a.npts * DIMP;
      cp_async16(hb + 64 + q * 2, a.XH + (ok ? e : 0), ok);
    }
    cp_async_mbar_arrive(&bar_full[buf]);
  };
  if (n_my > 0) issue_tile(0);
  if (n_my > 1) issue_tile(1);

  double accv[3][CT][2];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < CT; ++c) accv[r][c][0] = accv[r][c][1] = 0.0;

  LikParams lp;
  lp.kind = GPFY_LIK_GAUSSIAN;
  lp.sigma2 = FUSE_LIK ? a.s2ptr[0] : 1.0;
  const double kv = FUSE_LIK ? a.vptr[0] : 1.0;

  for (int it = 0; it < n_my; ++it) {
    const int buf = it & 1;
    const int64_t p0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * 32;
    const int64_t n = p0 + lane;
    const bool valid = n < a.npts;
    const int tw = it % ZB2_WARPS;  // tail warp of this tile
    double* cb = cbuf + buf * Mp * ZB2_LD;
    const double* hb = hdr + buf * HDR;
    // the tail warp fetches its per-point scalars early
    double t_r = 1.0, t_mz = 0.0, t_y = 0.0, t_dr = 0.0;
    if (warp == tw && valid) {
      t_r = a.R[n];
      if (FUSE_LIK) { t_mz = a.mz[n]; t_y = a.Y[n]; }
      else t_dr = a.dr[n];
    }
    mbar_wait(&bar_full[buf], (it >> 1) & 1);

    // ---- phase 1 ----
    double xh[DIMP];
#pragma unroll
    for (int c = 0; c < DIMP; c += 2) {
      const double2 v2 = *reinterpret_cast<const double2*>(hb + 64 + lane * DIMP + c);
      xh[c] = v2.x; xh[c + 1] = v2.y;
    }
    if (warp == 0) {
#pragma unroll
      for (int c = 0; c < DIMP; ++c) xsT[c * ZB2_XLD + lane] = xh[c];
    }
    const double cpw = hb[lane], cqw = hb[32 + lane];
    double psi = 0.0;
    for (int b = warp; b < Mp / 4; b += ZB2_WARPS) {
      const int r0 = 4 * b;
      double cval[4], tt[4], z[4], zp[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) cval[u] = cb[zb2_swz(r0 + u, lane)];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double* v = vs + (r0 + u) * DIMP;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int c = 0; c < DIMP; c += 2) {
          const double2 vv = *reinterpret_cast<const double2*>(v + c);
          s0 = fma(vv.x, xh[c], s0);
          s1 = fma(vv.y, xh[c + 1], s1);
        }
        tt[u] = s0 + s1;
      }
      const int l0 = rowdeg[r0], l3 = rowdeg[r0 + 3];
      if (l0 == l3 && l0 >= 1) {  // four rows of one degree: interleaved recurrences
        double Dm2[4], Dm1[4], D[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { Dm2[u] = 0.0; Dm1[u] = 1.0; D[u] = a.rc.two_alpha_p1 * tt[u]; }
        for (int k = 2; k <= l0; ++k) {
          const double d1 = a.rc.d1[k], d2 = a.rc.d2[k];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const double Dn = fma(d1 * tt[u], D[u], -d2 * Dm1[u]);
            Dm2[u] = Dm1[u];
            Dm1[u] = D[u];
            D[u] = Dn;
          }
        }
        const double zrl = zr[l0];
#pragma unroll
        for (int u = 0; u < 4; ++u) { zp[u] = a.rc.two_alpha * Dm1[u]; z[u] = zrl * (D[u] - Dm2[u]); }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) geg_pair(a.rc, zr, rowdeg[r0 + u], tt[u], z[u], zp[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double dz = fma(cqw, cval[u], cpw * mu2s[r0 + u]);
        cb[zb2_swz(r0 + u, lane)] = dz * zp[u];
        psi = fma(z[u], cval[u], psi);
      }
    }
    psir[(buf * ZB2_WARPS + warp) * 32 + lane] = psi;
    __syncthreads();  // A: dT tile, psi parts and xhat^T complete

    // ---- phase 2a: dVhat += dT Xhat ----
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
      double af[3], bf[CT];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int row = (warp + ZB2_WARPS * r) * 8 + g;
        af[r] = row < Mp ? cb[zb2_swz(row, 4 * k4 + t4)] : 0.0;
      }
#pragma unroll
      for (int c = 0; c < CT; ++c) bf[c] = xsT[(8 * c + g) * ZB2_XLD + 4 * k4 + t4];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < CT; ++c) dmma884(accv[r][c][0], accv[r][c][1], af[r], bf[c]);
    }
    // ---- phase 2b: dxhat = Vhat^T dT, row groups kg, dim tile mt ----
    {
      const int kg = warp % ZB2_KG, mt = warp / ZB2_KG;
      if (8 * mt < DIMP) {
        const int nk4 = Mp / 4, per = (nk4 + ZB2_KG - 1) / ZB2_KG;
        const int k_lo = kg * per, k_hi = min(nk4, k_lo + per);
        double acc2[4][2];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) acc2[nt][0] = acc2[nt][1] = 0.0;
        const int dimi = 8 * mt + g;
        for (int k4 = k_lo; k4 < k_hi; ++k4) {
          const int row = 4 * k4 + t4;
          const double av = dimi < DIMP ? vs[row * DIMP + dimi] : 0.0;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) dmma884(acc2[nt][0], acc2[nt][1], av, cb[zb2_swz(row, 8 * nt + g)]);
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          *reinterpret_cast<double2*>(red + (kg * 16 + dimi) * 32 + 8 * nt + 2 * t4) = make_double2(acc2[nt][0], acc2[nt][1]);
      }
    }
    __syncthreads();  // B: red complete, C/dT buffer free
    if (it + 2 < n_my) issue_tile(it + 2);

    // ---- tail: one warp, lane = point ----
    if (warp == tw) {
      double part[4 + DIMP];  // DB, DW[0..d), then E, DV, DS2 at the end
#pragma unroll
      for (int k = 0; k < 4 + DIMP; ++k) part[k] = 0.0;
      if (valid) {
        double dxh[DIMP];
#pragma unroll
        for (int c = 0; c < DIMP; ++c) {
          double s = 0.0;
#pragma unroll
          for (int kg = 0; kg < ZB2_KG; ++kg) s += red[(kg * 16 + c) * 32 + lane];
          dxh[c] = s;
        }
        const double r = t_r;
        double drv = t_dr;
        if (FUSE_LIK) {
          double ps = 0.0;
#pragma unroll
          for (int w = 0; w < ZB2_WARPS; ++w) ps += psir[(buf * ZB2_WARPS + w) * 32 + lane];
          double kd = kv, dkd = 0.0;
          {
            double rp = 1.0;
            for (int k = 0; k < 2 * a.order - 1; ++k) rp *= r;
            if (a.order > 0) { kd = kv * rp * r; dkd = 2.0 * a.order * kv * rp; }
          }
          const double fmu = r * t_mz, fvar = kd + r * r * ps;
          double e, pp, qq, ds2;
          lik_terms(lp, fmu, fvar, t_y, e, pp, qq, ds2);
          part[1 + DIMP] = e;
          part[2 + DIMP] = qq * (kd / kv);
          part[3 + DIMP] = ds2;
          drv = pp * t_mz + qq * (dkd + 2.0 * r * ps);
        }
        double dot = 0.0;
#pragma unroll
        for (int c = 0; c < DIMP; ++c) dot = fma(xh[c], dxh[c], dot);
#pragma unroll
        for (int c = 0; c < DIMP; ++c) {
          const double du = (dxh[c] - xh[c] * dot) / r + drv * xh[c];
          const double u = xh[c] * r;
          if (c < a.d) part[1 + c] = du * u;
          else if (c == a.d) part[0] = du;
        }
      }
      double* sw = sums + warp * ZB2_SUMS;
#pragma unroll
      for (int k = 0; k < 4 + DIMP; ++k) {
        const double s = warp_sum(part[k]);
        if (lane == 0) sw[k] += s;
      }
    }
  }

  // ---- write-out (accumulating over chunks) ----
  __syncthreads();
  if (tid < a.npart) {
    // sums layout per warp: [0] DB, [1 .. 1 + d) DW, [1 + DIMP ..] E, DV, DS2
    int k;
    if (tid == PART_E) k = 1 + DIMP;
    else if (tid == PART_DV) k = 2 + DIMP;
    else if (tid == PART_DS2) k = 3 + DIMP;
    else if (tid == PART_DB) k = 0;
    else k = 1 + (tid - PART_DW);
    double s = 0.0;
    for (int w = 0; w < ZB2_WARPS; ++w) s += sums[w * ZB2_SUMS + k];
    if (FUSE_LIK || (tid != PART_E && tid != PART_DV && tid != PART_DS2)) a.partials[(int64_t)blockIdx.x * a.npart + tid] += s;
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int row = (warp + ZB2_WARPS * r) * 8 + g;
    if (row < Mp) {
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        double2* dst = reinterpret_cast<double2*>(a.dVpart + ((int64_t)blockIdx.x * Mp + row) * (CT * 8) + c * 8 + 2 * t4);
        double2 o = *dst;
        o.x += accv[r][c][0];
        o.y += accv[r][c][1];
        *dst = o;
      }
    }
  }
}

}  // namespace gpfy
