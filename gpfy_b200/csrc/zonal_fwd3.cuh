// K_A3  zonal forward without CTA barriers: the same arithmetic per 8-row tile as zonal_fwd2_kernel (to_sphere spherical.py:214-249,
// t = Vhat xhat spherical_harmonics.py:308 on the FP64 tensor cores, monic Gegenbauer recurrence gegenbauer.py:47-60, mu2 . z
// variational.py:305 folded, Gaussian per-point weights likelihoods.py:210-218), but ONE WARP carries a whole 32-point tile through all
// row tiles.  mu2 . z is then a sum inside the warp (registers, three shuffles at the end) instead of a sum over the CTA's warps, so the
// per-tile CTA barrier, the helper warp and the double-buffered partial sums of zonal_fwd2_kernel disappear: warps drift freely and the
// FP64 pipe always has sixteen independent instruction streams per SM to pick from (features3.cuh is the same idea for Phi).
#pragma once
#include "zonal_fwd2.cuh"

namespace gpfy {

constexpr int ZF3_WARPS = 8;

static inline int zf3_smem_bytes(int Mp, int dimp, int P, bool with_mz, int pmax) {
  int dbl = Mp * dimp + (with_mz ? P * Mp : 0) + ZF3_WARPS * (32 * dimp + pmax * 32 + 32) + 3 * GPFY_MAX_DEGREES + 10;
  return dbl * 8 + Mp * 4 + 16;
}

template <int DIMP, int PMAX, bool ZPOUT>
__global__ void __launch_bounds__(32 * ZF3_WARPS, 2) zonal_fwd3_kernel(const __grid_constant__ ZonalFwd2Args args) {
  constexpr int KS = DIMP / 4;
  constexpr int NTHR = 32 * ZF3_WARPS;
  const ZonalArgs& a = args.z;
  extern __shared__ __align__(16) double smem[];
  __shared__ uint64_t bar;
  const int Mp = a.Mp;
  double* vs = smem;                                   // [Mp][DIMP]
  double* mu2s = vs + Mp * DIMP;                       // [P][Mp] (only when a.mz)
  double* xsw = mu2s + (a.mz ? a.P * Mp : 0);          // [warps][32][DIMP] the tile's coordinates
  double* mzw = xsw + ZF3_WARPS * 32 * DIMP;           // [warps][PMAX][32] mu2 . z of the tile's points
  double* rsw = mzw + ZF3_WARPS * PMAX * 32;           // [warps][32] radii
  double* leads = rsw + ZF3_WARPS * 32;                // [MAX_DEGREES + 2]
  double* betas = leads + GPFY_MAX_DEGREES + 2;        // [MAX_DEGREES + 6]
  double* dcos = betas + GPFY_MAX_DEGREES + 6;         // [MAX_DEGREES + 2]
  int* rowdeg = reinterpret_cast<int*>(dcos + GPFY_MAX_DEGREES + 2);  // [Mp]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int64_t npts_e = (a.npts + 1) & ~(int64_t)1;
  const int64_t tiles = (npts_e + 31) / 32;

  for (int i = tid; i < Mp; i += NTHR) {
    int l = -1;
    if (i < a.M) {
      int o = 0;
      l = 0;
      while (i >= o + a.deg.m[l]) { o += a.deg.m[l]; ++l; }
    }
    rowdeg[i] = l;
  }
  for (int i = tid; i < GPFY_MAX_DEGREES + 2; i += NTHR) { leads[i] = args.mt.lead[i]; dcos[i] = args.mt.dco[i]; }
  for (int i = tid; i < GPFY_MAX_DEGREES + 6; i += NTHR) betas[i] = i < GPFY_MAX_DEGREES + 2 ? args.mt.beta[i] : 0.0;
  stage_params_tma(vs, a.vhat_p, (unsigned)((Mp * DIMP + (a.mz ? a.P * Mp : 0)) * 8), &bar);
  __syncthreads();   // the only CTA barrier: parameters staged

  double* x = xsw + warp * 32 * DIMP;
  double* mzs = mzw + warp * PMAX * 32;
  double* rs = rsw + warp * 32;
  const int ntile = (ZPOUT || a.Zh) ? Mp / 8 : (a.M + 7) / 8;   // the derivative panel / float planes cover the padding rows too (zeros)
  const int64_t stride = (int64_t)gridDim.x * ZF3_WARPS;

  for (int64_t tile = (int64_t)blockIdx.x * ZF3_WARPS + warp; tile < tiles; tile += stride) {
    const int64_t p0 = tile * 32;
    // ---- lane = point: onto the sphere; coordinates to the warp's tile, r and xhat to HBM ----
    {
      const int64_t n = p0 + lane;
      double xh[DIMP];
#pragma unroll
      for (int c = 0; c < DIMP; ++c) xh[c] = 0.0;
      double r = 1.0;
      if (n < a.npts) {
        r = load_point<DIMP>(a.X, a.xcols, a.apply_to_sphere, a.sw, a.sb, a.dim, n, xh, a.proj);
        if (a.R) a.R[n] = r;
        if (a.XH) {
#pragma unroll
          for (int c = 0; c < DIMP; c += 2) *reinterpret_cast<double2*>(a.XH + n * DIMP + c) = make_double2(xh[c], xh[c + 1]);
        }
      }
#pragma unroll
      for (int c = 0; c < DIMP; c += 2) *reinterpret_cast<double2*>(x + lane * DIMP + c) = make_double2(xh[c], xh[c + 1]);
      rs[lane] = r;
    }
    __syncwarp();
    double xb[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) xb[ks][nt] = x[(8 * nt + g) * DIMP + 4 * ks + t4];

    double mzacc[PMAX][8];
#pragma unroll
    for (int p = 0; p < PMAX; ++p)
#pragma unroll
      for (int u = 0; u < 8; ++u) mzacc[p][u] = 0.0;

#pragma unroll 1
    for (int rt = 0; rt < ntile; ++rt) {
      const int row = rt * 8 + g;
      double t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = 0.0;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const double av = vs[row * DIMP + 4 * ks + t4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma884(t[2 * nt], t[2 * nt + 1], av, xb[ks][nt]);
      }
      const int l = rowdeg[row];
      const int lo = rowdeg[rt * 8], hi = rowdeg[rt * 8 + 7];
      double p[8], z[8];
      if constexpr (ZPOUT) {
        double zp[8];
        if (lo == hi && lo >= 1) {
          monic_uniform_d<8>(betas, dcos, lo, t, p, zp);
          const double ld = leads[lo];
#pragma unroll
          for (int u = 0; u < 8; ++u) z[u] = ld * p[u];
        } else {
          const int lmax = __reduce_max_sync(0xffffffffu, l);
          monic_mixed_d<8>(betas, dcos, l, lmax, t, p, zp);
          const double ld = l >= 1 ? leads[l] : 0.0;
#pragma unroll
          for (int u = 0; u < 8; ++u) z[u] = l >= 1 ? ld * p[u] : (l == 0 ? 1.0 : 0.0);
        }
        double* zrow = a.ZPb + ((p0 >> 5) * (int64_t)Mp + row) * 32;
        const int swz = (row & 3) << 2;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          *reinterpret_cast<double2*>(zrow + ((8 * nt + 2 * t4) ^ swz)) = make_double2(zp[2 * nt], zp[2 * nt + 1]);
      } else {
        double p1[8], p2[8];
        if (lo == hi && lo >= 1) {
          monic_uniform<8>(betas, lo, t, p, p1, p2);
          const double ld = leads[lo];
#pragma unroll
          for (int u = 0; u < 8; ++u) z[u] = ld * p[u];
        } else {
          const int lmax = __reduce_max_sync(0xffffffffu, l);
          monic_mixed<8>(betas, l, lmax, t, p, p1, p2);
          const double ld = l >= 1 ? leads[l] : 0.0;
#pragma unroll
          for (int u = 0; u < 8; ++u) z[u] = l >= 1 ? ld * p[u] : (l == 0 ? 1.0 : 0.0);
        }
      }
      if (a.Zh) {   // the same values as two TF32 planes (hi + lo), tile-contiguous [n / 32][Mp][32], for the tcgen05 contractions;
                    // padding rows (z = 0) and the columns past the end of the chunk are written as zeros
        float* zhrow = a.Zh + ((p0 >> 5) * (int64_t)Mp + row) * 32;
        float* zlrow = a.Zl + ((p0 >> 5) * (int64_t)Mp + row) * 32;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int64_t n0 = p0 + 8 * nt + 2 * t4;
          float h0, l0, h1, l1;
          tf32_split(row < a.M ? z[2 * nt] : 0.0, h0, l0);
          tf32_split(row < a.M ? z[2 * nt + 1] : 0.0, h1, l1);
          if (n0 + 1 >= a.npts) { h1 = 0.f; l1 = 0.f; }
          if (n0 >= a.npts) { h0 = 0.f; l0 = 0.f; }
          *reinterpret_cast<float2*>(zhrow + 8 * nt + 2 * t4) = make_float2(h0, h1);
          *reinterpret_cast<float2*>(zlrow + 8 * nt + 2 * t4) = make_float2(l0, l1);
        }
      }
      if (row < a.M) {
        if (!(args.dbg & 1) && a.Z)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int64_t n0 = p0 + 8 * nt + 2 * t4;
          double* dst = a.Z + (int64_t)row * a.ldz + n0;
          if (n0 + 1 < a.npts) *reinterpret_cast<double2*>(dst) = make_double2(z[2 * nt], z[2 * nt + 1]);
          else if (n0 < a.npts) {
            dst[0] = z[2 * nt];
            if (n0 + 1 < npts_e) dst[1] = 0.0;  // keep the even-padded tail column finite
          }
        }
        if (a.mz) {
#pragma unroll
          for (int pp = 0; pp < PMAX; ++pp)
            if (pp < a.P) {
              const double m2 = mu2s[pp * Mp + row];
#pragma unroll
              for (int u = 0; u < 8; ++u) mzacc[pp][u] = fma(m2, z[u], mzacc[pp][u]);
            }
        }
      }
    }
    if (a.mz) {
      // mz_p[n] = mu2_p . z_n: over the 8 row lanes by shuffles (fixed order), then lane = point finishes the tile's per-point outputs
#pragma unroll
      for (int pp = 0; pp < PMAX; ++pp)
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          double v = mzacc[pp][u];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (g == 0) mzs[pp * 32 + 8 * (u >> 1) + 2 * t4 + (u & 1)] = v;
        }
      __syncwarp();
      const int64_t n = p0 + lane;
      const bool valid = n < a.npts;
      const double r = rs[lane];
#pragma unroll
      for (int pp = 0; pp < PMAX; ++pp)
        if (pp < a.P) {
          const double s = mzs[pp * 32 + lane];
          if (valid) {
            a.mz[(int64_t)pp * a.ldz + n] = s;
            if (a.Yg) {
              const double inv = 1.0 / a.s2g[0];
              const double pw = (a.Yg[n] - r * s) * inv, qw = -0.5 * inv;
              a.wp[n] = pw * r;
              a.wq[n] = qw * r * r;
              a.cq[n] = 2.0 * qw * r * r;
            }
          }
        }
      if (!valid && n < npts_e && a.Yg) { a.wp[n] = 0.0; a.wq[n] = 0.0; a.cq[n] = 0.0; }
    }
    __syncwarp();   // the warp's tile and sums are consumed before the next tile's overwrite them
  }
}

}  // namespace gpfy
