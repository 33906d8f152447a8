// Device kernels of the SVGP hot path (see DESIGN.md for the algorithm and the reference lines each
// step restates).  Notation: Z = un-whitened zonal values C_l^alpha(Vhat xhat) [Mp, ld] (feature-major,
// points contiguous, like the reference's [M, N] arrays); C = W2 Z; dT = cotangent of t = Vhat xhat.
#pragma once
#include "common.cuh"
#include "gemm_i8.cuh"
#include "gh20.h"

namespace gpfy {

struct DegreeTable {
  int F;
  int m[GPFY_MAX_DEGREES];
};

// ------------------------------------------------------------------------------------------------
// K_A  zonal forward: to_sphere (spherical.py:214-249) + V xhat (spherical_harmonics.py:308) +
// Gegenbauer recurrence in registers (gegenbauer.py:47-60).  One point per thread, Vhat staged in
// shared memory by a TMA bulk copy and read as warp-wide broadcasts.
// ------------------------------------------------------------------------------------------------
struct ZonalArgs {
  const double* X;      // [n, xcols] row-major; xcols = d (raw) or dim (already on the sphere)
  int xcols;
  int apply_to_sphere;  // 1: scale by sw, append sb, normalise;  0: X rows are unit vectors of length dim
  int proj;             // > 0: sw is a [xcols, proj] projection matrix W and u = x W (spherical.py:225-226)
  const double* sw;     // [d] sqrt(weight_variances) (already broadcast in scalar mode)
  const double* sb;     // [1] sqrt(bias_variance)
  const double* vhat_p; // [Mp, DIMP] zero padded
  int Mp, M, dim;
  int64_t npts;
  double* Z;            // [Mp(+pad), ldz]
  int64_t ldz;
  double* XH;           // [npts, DIMP] unit vectors (zero padded), may be null
  double* R;            // [npts] radii, may be null
  double* mz;           // [P][ldz] mu2_p . z_n, may be null (then mu2 is not staged)
  int P;
  // Gaussian likelihood, P = 1 (fused path): the per-point weights only need fmu = r mu2 . z, so they are
  // produced here (likelihoods.py:210-218: p = (y - fmu) / s2, q = -1 / (2 s2)); null otherwise
  const double* Yg;     // [npts]
  const double* s2g;    // [1]
  double* wq;           // [ldz] q r^2
  double* wp;           // [ldz] p r
  double* cq;           // [ldz] 2 q r^2
  double* ZPb;          // [tiles][Mp][32] blocked + swizzled derivative panel d/dt C_l^alpha(t) for zonal_bwd4, or null
  float* Zh;            // hi / lo TF32 planes of Z, tile-contiguous [n / 32][Mp][32], for the tcgen05 contractions (GPFY_PREC_TF32X3), or null
  float* Zl;
  DegreeTable deg;
  RecCoef rc;
};

template <int DIMP>
__device__ __forceinline__ double load_point(const double* __restrict__ X, int xcols, int apply, const double* sw,
                                             const double* sb, int dim, int64_t n, double (&xh)[DIMP], int proj = 0) {
  const double* x = X + n * (int64_t)xcols;
  double r = 1.0;
  if (apply) {
    double ss = 0.0;
    if (proj > 0) {                           // u = x W, W [xcols, proj]   spherical.py:225-226
#pragma unroll
      for (int c = 0; c < DIMP; ++c) xh[c] = 0.0;
#pragma unroll 1
      for (int j = 0; j < xcols; ++j) {
        const double xj = x[j];
        const double* wr = sw + j * proj;
#pragma unroll
        for (int c = 0; c < DIMP; ++c)
          if (c < proj) xh[c] = fma(xj, wr[c], xh[c]);
      }
#pragma unroll
      for (int c = 0; c < DIMP; ++c) {
        if (c == proj) xh[c] = sb[0];         // bias = sqrt(b)           spherical.py:245-247
        ss += xh[c] * xh[c];
      }
    } else {
#pragma unroll
      for (int c = 0; c < DIMP; ++c) {
        double u = 0.0;
        if (c < xcols) u = x[c] * sw[c];       // X * sqrt(w)              spherical.py:228
        else if (c == xcols) u = sb[0];         // bias = sqrt(b)           spherical.py:245-247
        xh[c] = u;
        ss += u * u;
      }
    }
    r = sqrt(ss);                             // spherical.py:248
    const double rinv = 1.0 / r;              // spherical.py:249 (one reciprocal: <= 1 ulp from the division)
#pragma unroll
    for (int c = 0; c < DIMP; ++c) xh[c] = xh[c] * rinv;
  } else {
#pragma unroll
    for (int c = 0; c < DIMP; ++c) xh[c] = (c < dim) ? x[c] : 0.0;
  }
  return r;
}

// Block = 8 warps x 32 points: lane = point, warp w takes rows j = w, w + 8, ... of every degree, so
// per-thread dependent work is M / 8 rows and every global access is a 256-byte coalesced row segment.
template <int DIMP, int PMAX>
__global__ void __launch_bounds__(256) zonal_fwd_kernel(ZonalArgs a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ uint64_t bar;
  __shared__ double red[8][PMAX][32];
  double* vs = smem;                     // [Mp][DIMP]
  double* mu2s = smem + a.Mp * DIMP;     // [P][Mp] (only when a.mz)
  stage_params_tma(vs, a.vhat_p, (unsigned)((a.Mp * DIMP + (a.mz ? a.P * a.Mp : 0)) * 8), &bar);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * 32 + lane;
  const bool valid = n < a.npts;
  const bool padcol = !valid && n < ((a.npts + 1) & ~(int64_t)1);  // keep the even-padded tail column finite
  double xh[DIMP];
  double r = 1.0;
  if (valid) {
    r = load_point<DIMP>(a.X, a.xcols, a.apply_to_sphere, a.sw, a.sb, a.dim, n, xh, a.proj);
    if (warp == 0) {
      if (a.R) a.R[n] = r;
      if (a.XH) {
#pragma unroll
        for (int c = 0; c < DIMP; c += 2) *reinterpret_cast<double2*>(a.XH + n * DIMP + c) = make_double2(xh[c], xh[c + 1]);
      }
    }
  }
  double mz[PMAX];
#pragma unroll
  for (int p = 0; p < PMAX; ++p) mz[p] = 0.0;
  int base = 0;
  for (int l = 0; l < a.deg.F; ++l) {
    const int ml = a.deg.m[l];
    if (valid) {
      GPFY_DEGREE_SWITCH(l, {
        for (int j = warp; j < ml; j += 8) {
          const int i = base + j;
          const double* v = vs + i * DIMP;
          double t0 = 0.0, t1 = 0.0;
#pragma unroll
          for (int c = 0; c < DIMP; c += 2) {
            double2 vv = *reinterpret_cast<const double2*>(v + c);
            t0 = fma(vv.x, xh[c], t0);
            t1 = fma(vv.y, xh[c + 1], t1);
          }
          const double z = geg_eval_d<LS>(a.rc, l, t0 + t1);
          a.Z[(int64_t)i * a.ldz + n] = z;
          if (a.mz) {
#pragma unroll
            for (int p = 0; p < PMAX; ++p)
              if (p < a.P) mz[p] = fma(mu2s[p * a.Mp + i], z, mz[p]);
          }
        }
      });
    } else if (padcol) {
      for (int j = warp; j < ml; j += 8) a.Z[(int64_t)(base + j) * a.ldz + n] = 0.0;
    }
    base += ml;
  }
  if (a.mz) {  // mz_p[n] = mu2_p . z_n  (variational.py:305 folded), fixed-order sum over the 8 warps
#pragma unroll
    for (int p = 0; p < PMAX; ++p) red[warp][p][lane] = mz[p];
    __syncthreads();
    if (warp == 0 && valid) {
#pragma unroll
      for (int p = 0; p < PMAX; ++p)
        if (p < a.P) {
          double s = 0.0;
#pragma unroll
          for (int w = 0; w < 8; ++w) s += red[w][p][lane];
          a.mz[(int64_t)p * a.ldz + n] = s;
          if (a.Yg) {
            const double inv = 1.0 / a.s2g[0];
            const double pw = (a.Yg[n] - r * s) * inv, qw = -0.5 * inv;
            a.wp[n] = pw * r;
            a.wq[n] = qw * r * r;
            a.cq[n] = 2.0 * qw * r * r;
          }
        }
    }
    if (warp == 0 && padcol && a.Yg) { a.wp[n] = 0.0; a.wq[n] = 0.0; a.cq[n] = 0.0; }
  }
}

// ------------------------------------------------------------------------------------------------
// K_F  per-degree whitening for the explicit feature matrix (spherical_harmonics.py:316, :362, :385):
//   out[off_l + i, n] = scale_l * r_n * sum_k Linv_l[i][k] z[off_l + k, n]
// One point per thread; the zonal values of a degree are held in registers 32 at a time and the
// rows of Linv_l are read as warp-uniform (broadcast) loads.  Linv is the dense block-diagonal
// [Mp, Mp] matrix, so entries right of the block are exact zeros.
// ------------------------------------------------------------------------------------------------
struct WhitenArgs {
  const double* Z;
  int64_t ldz;
  const double* Linv;   // [Mp (+slack)][Mp]
  int Mp;
  const double* scale;  // [F] or null
  const double* R;      // [npts] or null
  int64_t npts;
  double* out;          // [M][ldo], already offset to this chunk's first column
  int64_t ldo;
  DegreeTable deg;
};

__global__ void __launch_bounds__(128) whiten_kernel(WhitenArgs a) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.npts) return;
  const double r = a.R ? a.R[n] : 1.0;
  int off = 0;
  for (int l = 0; l < a.deg.F; ++l) {
    const int m = a.deg.m[l];
    const double sc = (a.scale ? a.scale[l] : 1.0) * r;
    for (int kc = 0; kc < m; kc += 32) {
      double z[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) z[k] = (kc + k < m) ? a.Z[(int64_t)(off + kc + k) * a.ldz + n] : 0.0;
      for (int i = kc; i < m; ++i) {
        const double* Li = a.Linv + (int64_t)(off + i) * a.Mp + off + kc;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          s0 = fma(__ldg(Li + k), z[k], s0);
          s1 = fma(__ldg(Li + k + 1), z[k + 1], s1);
        }
        double* dst = a.out + (int64_t)(off + i) * a.ldo + n;
        const double val = sc * (s0 + s1);
        *dst = (kc == 0) ? val : *dst + val;
      }
    }
    off += m;
  }
}

// Elementwise Gegenbauer + derivative (gegenbauer.py:63-85)
__global__ void gegenbauer_kernel(RecCoef rc, int level, const double* __restrict__ x, int64_t n, double* out,
                                  double* dout) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double t = x[i];
  out[i] = geg_eval(rc, level, t);
  if (dout) dout[i] = geg_deriv(rc, level, t);
}

// ------------------------------------------------------------------------------------------------
// Likelihood terms: e, p = de/dfmu, q = de/dfvar for one (point, output)
// ------------------------------------------------------------------------------------------------
struct LikParams {
  int kind;          // GPFY_LIK_*
  double sigma2;     // Gaussian noise variance
};
__constant__ double c_gh_x[GPFY_GH_N] = GPFY_GH_X_INIT;
__constant__ double c_gh_w[GPFY_GH_N] = GPFY_GH_W_INIT;

__device__ __forceinline__ void lik_terms(const LikParams& lp, double fmu, double fvar, double y, double& e, double& p,
                                          double& q, double& ds2) {
  if (lp.kind == GPFY_LIK_GAUSSIAN) {
    // likelihoods.py:210-218
    const double inv = 1.0 / lp.sigma2;
    const double res = y - fmu;
    const double s = res * res + fvar;
    e = -0.5 * 1.8378770664093454835606594728112 - 0.5 * log(lp.sigma2) - 0.5 * s * inv;
    p = res * inv;
    q = -0.5 * inv;
    ds2 = -0.5 * inv + 0.5 * s * inv * inv;
  } else {
    // likelihoods.py:254-292 + quadrature.py:30-53: 20-node Gauss-Hermite of the Bernoulli log-density
    // with probs = Phi(f) (1 - 2e-3) + 1e-3
    const double sd = sqrt(fvar);
    const double s2sd = 1.4142135623730951 * sd;
    const double inv_sqrt_pi = 0.56418958354775628694807945156077;
    const double inv_sqrt_2pi = 0.39894228040143267793994605993438;
    double ee = 0.0, pp = 0.0, qq = 0.0;
#pragma unroll 1
    for (int i = 0; i < GPFY_GH_N; ++i) {
      const double xi = c_gh_x[i], wi = c_gh_w[i] * inv_sqrt_pi;
      const double f = fma(s2sd, xi, fmu);
      const double pr = 0.5 * (1.0 + erf(f * 0.70710678118654752440)) * (1.0 - 2e-3) + 1e-3;
      const double l0 = log1p(-pr), l1 = log(pr);
      ee += wi * ((1.0 - y) * l0 + y * l1);
      const double dl = (y / pr - (1.0 - y) / (1.0 - pr)) * (1.0 - 2e-3) * inv_sqrt_2pi * exp(-0.5 * f * f);
      pp += wi * dl;
      qq += wi * dl * xi;
    }
    e = ee;
    p = pp;
    q = qq / (1.4142135623730951 * sd);
    ds2 = 0.0;
  }
}

// ------------------------------------------------------------------------------------------------
// K_C  per-point likelihood + backward through the zonal map.
//   pass 1: psi_p = z . C_p,  fmu_p = r mu2_p . z          (variational.py:305,326-327 folded)
//   fvar_p = v r^(2 order) + r^2 psi_p                       (model.py:136, spherical.py:400)
//   e, p, q from the likelihood;  wq = q r^2, wp = p r (weights of the SYRK / mean reductions)
//   pass 2: dz = sum_p r p mu2_p + 2 q r^2 C_p ;  dt = dz * 2 alpha C_{l-1}^{alpha+1}(t)  (gegenbauer.py:73-81)
//           dT overwrites C_0;  dxhat += Vhat^T dt
//   du = (dxhat - xhat (xhat . dxhat)) / r + dr xhat  -> partial sums for d w, d b
// One point per thread; per-block partial sums written to `partials` (deterministic two-stage sum).
// ------------------------------------------------------------------------------------------------
constexpr int PART_E = 0, PART_DV = 1, PART_DS2 = 2, PART_DB = 3, PART_DW = 4;  // then d entries

struct LikPointArgs {
  const double* psi_part; // [P][npsi][ldz] partial sums of z . C_p from the GEMM epilogue
  int npsi;
  const double* mz;     // [P][ldz]
  int64_t ldz;
  const double* R;
  const double* Y;      // [npts][P]
  const double* vptr;   // [1] kernel variance
  const double* s2ptr;  // [1] noise variance (Gaussian)
  int P, order, lik_kind;
  int64_t npts;
  double* wq;           // [P][ldz]  q r^2   (SYRK weights)
  double* wp;           // [P][ldz]  p r     (mean-gradient weights)
  double* cq;           // [P][ldz]  2 q r^2 (coefficient of C in dz)
  double* dr;           // [ldz]     direct d/dr terms
  double* partials;     // [npts / 32][npart]: one row per 32 points (PART_E, PART_DV, PART_DS2 written here)
  int npart;
  int accumulate;
  // psi_n = z_n . C_n formed HERE from the two panels instead of by the contraction's epilogue (P = 1; used with the int8
  // contraction, whose epilogue warps pay ~40 cycles per FP64 instruction while the tensor pipe is saturated): Zp [Mp][ldz] and the
  // blocked, swizzled C (GemmArgs::c_blocked).  Null = read psi_part.
  const double* Zp = nullptr;
  const double* Cb = nullptr;
  int Mp = 0, M = 0;
};

// One point per thread: psi, fmu, fvar, the likelihood terms and everything the backward pass needs
// per point (model.py:136, spherical.py:400, likelihoods.py:210-218 / :254-278).
__global__ void __launch_bounds__(256) lik_point_kernel(LikPointArgs a) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = n < a.npts;
  double pe = 0.0, pdv = 0.0, pds2 = 0.0;
  if (!valid && n < ((a.npts + 1) & ~(int64_t)1)) {
    for (int p = 0; p < a.P; ++p) { a.wq[p * a.ldz + n] = 0.0; a.wp[p * a.ldz + n] = 0.0; }
  }
  if (valid) {
    LikParams lp;
    lp.kind = a.lik_kind;
    lp.sigma2 = a.lik_kind == GPFY_LIK_GAUSSIAN ? a.s2ptr[0] : 1.0;
    const double v = a.vptr[0];
    const double r = a.R[n];
    double kd = v, dkd = 0.0;   // K_diag = v r^(2 order) and its r-derivative
    {
      double rp = 1.0;
      for (int k = 0; k < 2 * a.order - 1; ++k) rp *= r;
      if (a.order > 0) { kd = v * rp * r; dkd = 2.0 * a.order * v * rp; }
    }
    const double r2 = r * r;
    double dr = 0.0;
    for (int p = 0; p < a.P; ++p) {
      double psi = 0.0;
      if (a.Cb) {   // both reads are coalesced over the warp's 32 points: a Z row and the matching 256-byte row of the C tile
        const double* zc = a.Zp + n;
        const double* cc = a.Cb + (n >> 5) * (int64_t)a.Mp * 32;
        const int ln = (int)(n & 31);
        double ps[4] = {0.0, 0.0, 0.0, 0.0};
        int row = 0;
        for (; row + 4 <= a.M; row += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) ps[u] = fma(zc[(int64_t)(row + u) * a.ldz], cc[(row + u) * 32 + (ln ^ (u << 2))], ps[u]);
        }
        for (; row < a.M; ++row) ps[0] = fma(zc[(int64_t)row * a.ldz], cc[row * 32 + (ln ^ ((row & 3) << 2))], ps[0]);
        psi = (ps[0] + ps[1]) + (ps[2] + ps[3]);
      } else
      for (int k = 0; k < a.npsi; ++k) psi += a.psi_part[((int64_t)p * a.npsi + k) * a.ldz + n];
      const double mzp = a.mz[(int64_t)p * a.ldz + n];
      const double fmu = r * mzp;
      const double fvar = kd + r2 * psi;
      double e, pp, qq, ds2;
      lik_terms(lp, fmu, fvar, a.Y[n * a.P + p], e, pp, qq, ds2);
      pe += e;
      pds2 += ds2;
      pdv += qq * (kd / v);
      a.wq[p * a.ldz + n] = qq * r2;
      a.wp[p * a.ldz + n] = pp * r;
      a.cq[p * a.ldz + n] = 2.0 * qq * r2;
      dr += pp * mzp + qq * (dkd + 2.0 * r * psi);
    }
    a.dr[n] = dr;
  }
  pe = warp_sum(pe); pdv = warp_sum(pdv); pds2 = warp_sum(pds2);
  if ((threadIdx.x & 31) == 0 && (n >> 5) < ((a.npts + 31) >> 5)) {
    double* dst = a.partials + (n >> 5) * a.npart;
    if (a.accumulate) { dst[PART_E] += pe; dst[PART_DV] += pdv; dst[PART_DS2] += pds2; }
    else { dst[PART_E] = pe; dst[PART_DV] = pdv; dst[PART_DS2] = pds2; }
  }
}

struct ZonalBwdArgs {
  const double* C;      // [P][Mp][ldz]
  double* dT;           // [Mp][ldz] (separate buffer: lets the loads of C run ahead of the stores)
  int64_t ldz;
  int64_t c_stride;     // Mp * ldz
  const double* wp;     // [P][ldz]  p r
  const double* cq;     // [P][ldz]  2 q r^2
  const double* dr;     // [ldz]
  const double* XH;     // [npts][DIMP]
  const double* R;
  const double* vhat_p; // [Mp][DIMP], mu2 [P][Mp] right behind it
  int Mp, M, d, P;      // d = sphere dim (index of the bias coordinate)
  const double* X;      // raw rows [npts][xcols] and W [xcols][proj]: only read for a projection kernel (proj > 0)
  const double* W;
  int xcols, proj;
  int64_t npts;
  double* partials;     // [npts / 32][npart]; PART_DB and PART_DW.. written here
  int npart;
  int accumulate;
  DegreeTable deg;
  RecCoef rc;
};

// Backward through the zonal map.  Block = 8 warps x 32 points (lane = point, warp w owns rows
// w, w + 8, ... of every degree):
//   dz = sum_p (r p) mu2_p + (2 q r^2) C_p ;  dt = dz * 2 alpha C_{l-1}^{alpha+1}(t)   (gegenbauer.py:73-81)
//   dT overwrites C_0;  dxhat += Vhat^T dt;  du = (dxhat - xhat (xhat . dxhat)) / r + dr xhat  (VJP of spherical.py:248-249)
template <int DIMP, int PMAX>
// xh + dxh alone are 2 DIMP doubles per thread: beyond DIMP = 16 the kernel runs one CTA per SM (up to 255 registers);
// under the 128-register cap of two CTAs it spilled 3-5 KB per thread (d = 32: 17.7 ms of a 30.8 ms step).
__global__ void __launch_bounds__(256, (DIMP > 16 ? 1 : 2)) zonal_bwd_kernel(ZonalBwdArgs a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ uint64_t bar;
  double* vs = smem;                      // [Mp][DIMP]
  double* mu2s = smem + a.Mp * DIMP;      // [P][Mp]
  double* red = mu2s + a.P * a.Mp;        // [8][DIMP][32]
  stage_params_tma(vs, a.vhat_p, (unsigned)((a.Mp * DIMP + a.P * a.Mp) * 8), &bar);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * 32 + lane;
  const bool valid = n < a.npts;
  double xh[DIMP], dxh[DIMP];
#pragma unroll
  for (int c = 0; c < DIMP; ++c) { xh[c] = 0.0; dxh[c] = 0.0; }
  if (valid) {
#pragma unroll
    for (int c = 0; c < DIMP; c += 2) {
      double2 t2 = *reinterpret_cast<const double2*>(a.XH + n * DIMP + c);
      xh[c] = t2.x; xh[c + 1] = t2.y;
    }
    double cp_[PMAX], cq_[PMAX];
#pragma unroll
    for (int p = 0; p < PMAX; ++p) {
      cp_[p] = (p < a.P) ? a.wp[p * a.ldz + n] : 0.0;
      cq_[p] = (p < a.P) ? a.cq[p * a.ldz + n] : 0.0;
    }
    const double* __restrict__ Cg = a.C;
    double* __restrict__ dTg = a.dT;
    int base = 0;
    for (int l = 0; l < a.deg.F; ++l) {
      const int ml = a.deg.m[l];
      GPFY_DEGREE_SWITCH(l, {
        // rows j, j + 8, j + 16, j + 24 of this warp are batched: 4 independent loads in flight
        for (int j0 = warp; j0 < ml; j0 += 32) {
          double cv[4][PMAX];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = j0 + 8 * u;
#pragma unroll
            for (int p = 0; p < PMAX; ++p)
              cv[u][p] = (j < ml && p < a.P) ? __ldg(Cg + p * a.c_stride + (int64_t)(base + j) * a.ldz + n) : 0.0;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = j0 + 8 * u;
            if (j < ml) {
              const int i = base + j;
              const double* vv = vs + i * DIMP;
              double t0 = 0.0, t1 = 0.0;
#pragma unroll
              for (int c = 0; c < DIMP; c += 2) {
                double2 v2 = *reinterpret_cast<const double2*>(vv + c);
                t0 = fma(v2.x, xh[c], t0);
                t1 = fma(v2.y, xh[c + 1], t1);
              }
              const double zp = geg_deriv_d<LS>(a.rc, l, t0 + t1);
              double dz = 0.0;
#pragma unroll
              for (int p = 0; p < PMAX; ++p)
                if (p < a.P) dz += cp_[p] * mu2s[p * a.Mp + i] + cq_[p] * cv[u][p];
              const double dt = dz * zp;
              dTg[(int64_t)i * a.ldz + n] = dt;
#pragma unroll
              for (int c = 0; c < DIMP; c += 2) {
                double2 v2 = *reinterpret_cast<const double2*>(vv + c);
                dxh[c] = fma(v2.x, dt, dxh[c]);
                dxh[c + 1] = fma(v2.y, dt, dxh[c + 1]);
              }
            }
          }
        }
      });
      base += ml;
    }
  }
#pragma unroll
  for (int c = 0; c < DIMP; ++c) red[(warp * DIMP + c) * 32 + lane] = dxh[c];
  __syncthreads();
  if (warp == 0) {
    double part[1 + DIMP];
#pragma unroll
    for (int k = 0; k < 1 + DIMP; ++k) part[k] = 0.0;
    if (valid) {
      const double r = a.R[n], dr = a.dr[n];
#pragma unroll
      for (int c = 0; c < DIMP; ++c) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[(w * DIMP + c) * 32 + lane];
        dxh[c] = s;
      }
      double dot = 0.0;
#pragma unroll
      for (int c = 0; c < DIMP; ++c) dot = fma(xh[c], dxh[c], dot);
#pragma unroll
      for (int c = 0; c < DIMP; ++c) {
        const double du = (dxh[c] - xh[c] * dot) / r + dr * xh[c];
        const double u = xh[c] * r;
        if (c < a.d) part[1 + c] = a.proj > 0 ? du : du * u;   // d w_j = sum du_j x_j / (2 sqrt w_j) = sum du_j u_j / (2 w_j)
        else if (c == a.d) part[0] = du;     // d b   = sum du_dim / (2 sqrt b)
      }
    }
    if (a.proj > 0) {
      // d W[j][k] = sum_n x_nj du_nk (VJP of X @ W, spherical.py:225-226); part[1 + k] holds du_k here.  The du
      // go through shared memory so that the (j, k) loops stay rolled (this is the rarely used path).
      __syncwarp();
#pragma unroll
      for (int c = 0; c < DIMP; ++c) red[c * 32 + lane] = c < a.d ? part[1 + c] : 0.0;
      __syncwarp();
      {
        const double s0 = warp_sum(part[0]);
        if (lane == 0) {
          double* dst = a.partials + (int64_t)blockIdx.x * a.npart + PART_DB;
          *dst = a.accumulate ? *dst + s0 : s0;
        }
      }
#pragma unroll 1
      for (int j = 0; j < a.xcols; ++j) {
        const double xj = valid ? a.X[n * a.xcols + j] : 0.0;
#pragma unroll 1
        for (int k = 0; k < a.proj; ++k) {
          const double sv = warp_sum(xj * red[k * 32 + lane]);
          if (lane == 0) {
            double* dst = a.partials + (int64_t)blockIdx.x * a.npart + PART_DW + j * a.proj + k;
            *dst = a.accumulate ? *dst + sv : sv;
          }
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 1 + DIMP; ++k) {
        const double s = warp_sum(part[k]);
        if (lane == 0 && k < 1 + a.d) {
          double* dst = a.partials + (int64_t)blockIdx.x * a.npart + (k == 0 ? PART_DB : PART_DW + k - 1);
          *dst = a.accumulate ? *dst + s : s;
        }
      }
    }
  }
}

// W <- lower(W + W^T) with the diagonal kept: z^T W z is unchanged for every z, the upper triangle becomes zero
__global__ void fold_lower_kernel(double* W, int n) {
  const int j = blockIdx.x * 32 + threadIdx.x;
  for (int i = blockIdx.y * 32 + threadIdx.y; i < n && i < (blockIdx.y + 1) * 32; i += 8) {
    if (j < i && j < n) {
      const double lo = W[(int64_t)i * n + j], up = W[(int64_t)j * n + i];
      W[(int64_t)i * n + j] = lo + up;
      W[(int64_t)j * n + i] = 0.0;
    }
  }
}

// predict_diag epilogue: fmu = r mz, fvar = v r^(2 order) + r^2 psi  [n, P]  (model.py:220-225)
struct PredictArgs {
  const double* psi_part;  // [P][npsi][ldz]
  int npsi;
  const double* mz;        // [P][ldz]
  int64_t ldz;
  const double* R;
  const double* vptr;
  int P, order;
  int64_t npts;
  double* fmu;         // [npts][P]
  double* fvar;
  // psi_n = z_n . C_n formed here from the two panels (int8 contraction, P = 1; see LikPointArgs): Zp [Mp][ldz], blocked C.  Null = psi_part.
  const double* Zp = nullptr;
  const double* Cb = nullptr;
  int Mp = 0, M = 0;
};

__global__ void __launch_bounds__(256) predict_kernel(PredictArgs a) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.npts) return;
  const double r = a.R[n];
  double kd = a.vptr[0];
  for (int k = 0; k < 2 * a.order; ++k) kd *= r;
  for (int p = 0; p < a.P; ++p) {
    double psi = 0.0;
    if (a.Cb) {
      const double* zc = a.Zp + n;
      const double* cc = a.Cb + (n >> 5) * (int64_t)a.Mp * 32;
      const int ln = (int)(n & 31);
      double ps[4] = {0.0, 0.0, 0.0, 0.0};
      int row = 0;
      for (; row + 4 <= a.M; row += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) ps[u] = fma(zc[(int64_t)(row + u) * a.ldz], cc[(row + u) * 32 + (ln ^ (u << 2))], ps[u]);
      }
      for (; row < a.M; ++row) ps[0] = fma(zc[(int64_t)row * a.ldz], cc[row * 32 + (ln ^ ((row & 3) << 2))], ps[0]);
      psi = (ps[0] + ps[1]) + (ps[2] + ps[3]);
    } else
    for (int k = 0; k < a.npsi; ++k) psi += a.psi_part[((int64_t)p * a.npsi + k) * a.ldz + n];
    a.fmu[n * a.P + p] = r * a.mz[(int64_t)p * a.ldz + n];
    a.fvar[n * a.P + p] = kd + r * r * psi;
  }
}

// to_sphere only (spherical.py:214-249); sqrt of the variances is taken once per block in shared memory
template <int DIMP>
__global__ void to_sphere_kernel(const double* X, int d, int dim, const double* w, int scalar_w, const double* b, int64_t n,
                                 double* xs, double* r, int proj) {
  __shared__ double ssw[64];
  __shared__ double ssb;
  if (proj == 0 && threadIdx.x < d) ssw[threadIdx.x] = sqrt(w[scalar_w ? 0 : threadIdx.x]);
  if (threadIdx.x == 0) ssb = sqrt(b[0]);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double xh[DIMP];
  const double rr = load_point<DIMP>(X, d, 1, proj > 0 ? w : ssw, &ssb, dim, i, xh, proj);
  r[i] = rr;
#pragma unroll
  for (int c = 0; c < DIMP; ++c)
    if (c < dim) xs[i * dim + c] = xh[c];
}

// expected log-likelihood per point, summed over outputs (likelihoods.py:188-220, :254-278)
__global__ void var_exp_kernel(int lik_kind, const double* s2ptr, int64_t n, int P, const double* fmu, const double* fvar,
                               const double* Y, double* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  LikParams lp;
  lp.kind = lik_kind;
  lp.sigma2 = lik_kind == GPFY_LIK_GAUSSIAN ? s2ptr[0] : 1.0;
  double s = 0.0;
  for (int p = 0; p < P; ++p) {
    double e, pp, qq, ds2;
    lik_terms(lp, fmu[i * P + p], fvar[i * P + p], Y[i * P + p], e, pp, qq, ds2);
    s += e;
  }
  out[i] = s;
}

// copy a chunk of A [M, N] (ld = N, any parity) into the padded panel [Mp(+pad)][ldz]
__global__ void copy_panel_kernel(const double* A, int64_t ldA, int M, int64_t npts, double* Zp, int64_t ldz) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (n < ((npts + 1) & ~(int64_t)1)) Zp[(int64_t)i * ldz + n] = (n < npts) ? A[(int64_t)i * ldA + n] : 0.0;
}

// mean[n, p] = sum_i mu[i, p] A[i, n];  var[n, p] = sum_k psi_part
__global__ void project_finish_kernel(const double* Zp, int64_t ldz, int M, int P, const double* mu, const double* psi_part,
                                      int npsi, int64_t npts, double* mean, double* var) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= npts) return;
  for (int p = 0; p < P; ++p) {
    double s = 0.0;
    for (int i = 0; i < M; ++i) s = fma(__ldg(mu + i * P + p), Zp[(int64_t)i * ldz + n], s);
    mean[n * P + p] = s;
    double v = 0.0;
    for (int k = 0; k < npsi; ++k) v += psi_part[((int64_t)p * npsi + k) * ldz + n];
    var[n * P + p] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// K_D  weighted SYRK on the FP64 tensor cores:  G (lower 96x96 blocks) += Z diag(wq) Z^T  and, on the
// diagonal blocks, bz += Z wp through one extra MMA column.  Split over points: CTA (blk, ks) owns a
// contiguous range of the chunk and accumulates into its own slot of Gpart (no atomics, deterministic).
// ------------------------------------------------------------------------------------------------
constexpr int SYRK_B = 96, SYRK_BK = 16, SYRK_STAGES = 3, SYRK_THREADS = 256;  // 768 Z chunks per operand per stage: 3 per thread
constexpr int SYRK_ZS = SYRK_BK + 4;
constexpr int SYRK_STAGE_DBL = 2 * SYRK_B * SYRK_ZS + 2 * SYRK_BK;
constexpr int SYRK_SMEM = SYRK_STAGES * SYRK_STAGE_DBL * 8;

// How the point range of a chunk is split among the CTAs of each 96 x 96 block.  A diagonal block costs a CTA 12 tensor-
// core tiles per warp and k-step, a full off-diagonal one 18, an off-diagonal block of a partial last block row (Mp = 352
// or 256: 64 of 96 rows) 12: every block TYPE gets its own number of splits, proportional to that cost, so all CTAs carry
// about the same work whatever the mix of types (with one uniform split count 4 diagonal + 6 off-diagonal blocks left a
// fifth of the SMs with two full off-diagonal CTAs: +22 % on the SYRK at 13 degrees).
struct SyrkSplit {
  int nb;        // row blocks
  int last_rows; // valid rows of the last block row (96 when Mp is a multiple of 96)
  int ks[4];     // splits of a block of type 0 diagonal, 1 off-diagonal, 2 off-diagonal in a partial last row, 3 partial diagonal
  __host__ __device__ int type(int bi, int bj) const {
    const bool part = bi == nb - 1 && last_rows < 96;
    return bi == bj ? (part ? 3 : 0) : (part ? 2 : 1);
  }
  __host__ __device__ int bslot(int bi) const {   // first bz slot of diagonal block bi (= its first G slot)
    int o = 0;
    for (int b = 0; b < bi; ++b) o += ks[type(b, b)];
    return o;
  }
  __host__ __device__ int gslot(int bi, int bj) const {   // first G slot of block (bi, bj): diagonal blocks, then (1,0), (2,0), (2,1), ...
    if (bi == bj) return bslot(bi);
    int o = bslot(nb);
    for (int i = 1; i < bi; ++i) o += i * ks[type(i, 0)];
    return o + bj * ks[type(bi, 0)];
  }
  __host__ __device__ int total() const { return gslot(nb, 0); }
};

struct SyrkArgs {
  const double* Z;
  int64_t ldz;
  int Mp;
  int64_t npts;
  const double* wq;   // [npts]
  const double* wp;   // [npts]
  double* Gpart;      // [split.total()][96*96]
  double* bzpart;     // [split.bslot(nb)][96]
  SyrkSplit split;
  int64_t chunk;      // rows of a full chunk: block type t covers it in split.ks[t] ranges of ceil(chunk / ks[t]) rows (multiple of 16)
  int accumulate;
  // GPFY_PREC_F64_I8 on the fused path: the diagonal-block CTAs, through whose shared memory every element of the panel passes
  // exactly once, also cut it into the int8 digit planes of the contraction (gemm_i8.cuh) - the separate slicing pass (a full
  // read of the float64 panel, HBM-bound) disappears.  Null = no slicing.
  int8_t* zs_out = nullptr;
  const double* kscale = nullptr;   // [Mp] 2^48 / s_k
  int* counter = nullptr;           // work-unit counter, zero at launch
  int ks_eff = 1;                   // point ranges per block used by THIS launch (<= split.ks[.]: short slabs use fewer, their other slots stay zero)
};

// One k-step (16 points) of the slicing: thread = one point x 16 features of the stage; the six 16-byte digit vectors of 16
// consecutive points are contiguous in the digit-plane image (256 bytes per plane).
__device__ __forceinline__ void syrk_slice_step(const double* zs, const double* ksc, int c, int pt, int8_t* dst) {
  Digits16 d;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    uint64_t t[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int f = 16 * c + 4 * w + j;
      t[j] = gi8_digits(zs[f * (SYRK_BK + 4) + pt], ksc[f]);
    }
    gi8_pack4(t, d, w);
  }
#pragma unroll
  for (int s = 0; s < GI8_S; ++s) *reinterpret_cast<uint4*>(dst + s * 4096) = make_uint4(d.w[s][0], d.w[s][1], d.w[s][2], d.w[s][3]);
}

// Diagonal blocks only need the lower triangle.  In 8 x 8 tensor-core tiles that is tile row r with columns
// 0..r (78 of 144 tiles); the 12 tile rows are dealt to the 8 warps as 11 | 10 | 9 | 8 | 7+0 | 6+1 | 5+2 | 4+3, i.e.
// 12, 11, 10, 9, 9, 9, 9, 9 tiles, which also evens out the four SM sub-partitions (warp w runs on w % 4).
// An off-diagonal block keeps the 4 x 2 grid of 24 x 48 warp tiles (18 tiles per warp).  Diagonal CTAs come first
// in the grid so that the two CTAs sharing an SM are one of each kind.
template <int RA, int RB>
__device__ __forceinline__ void syrk_diag_step(const double* zs, const double* wqs, double (&acc)[12][2], int g, int t4) {
  // one k4 step: rows of tile row RA (and RB < RA) against the columns 0..RA.  The weight goes onto the A fragments (one or two
  // multiplies per step) instead of every B fragment (RA + 1 of them: as many FP64-pipe multiplies as tensor-core tiles)
#pragma unroll
  for (int k4 = 0; k4 < SYRK_BK / 4; ++k4) {
    const double w = wqs[k4 * 4 + t4];
    const double afaw = zs[(RA * 8 + g) * SYRK_ZS + k4 * 4 + t4] * w;
    double afbw = 0.0;
    if constexpr (RB >= 0) afbw = zs[(RB * 8 + g) * SYRK_ZS + k4 * 4 + t4] * w;
#pragma unroll
    for (int c = 0; c <= RA; ++c) {
      const double bf = zs[(c * 8 + g) * SYRK_ZS + k4 * 4 + t4];
      dmma884(acc[c][0], acc[c][1], afaw, bf);
      if constexpr (RB >= 0) {
        if (c <= RB) dmma884(acc[RA + 1 + c][0], acc[RA + 1 + c][1], afbw, bf);
      }
    }
  }
}

__global__ void __launch_bounds__(SYRK_THREADS, 2) syrk_dmma_kernel(SyrkArgs a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ uint64_t bar_full[SYRK_STAGES], bar_empty[SYRK_STAGES];
  __shared__ double ksc_s[SYRK_B];
  __shared__ int unit_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp & 3, wn = warp >> 2;   // off-diagonal: 4 x 2 warps, warp tile 24 x 48
  const int g = lane >> 2, t4 = lane & 3;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < SYRK_STAGES; ++s) {
      mbar_init(&bar_full[s], SYRK_THREADS);
      mbar_init(&bar_empty[s], SYRK_THREADS / 32);
    }
  }
  // Persistent CTAs pull work units (block, point range) from a counter, heaviest block type first: full off-diagonal (18 tiles per
  // warp and k4 step), partial off-diagonal (12), diagonal (12 .. 9).  With one unit per CTA, as before, Mp = 352 / 256 left a fifth of
  // the SMs with two full off-diagonal CTAs (288 tiles per step against an average of 202).  Every unit owns its slot of the split-K
  // accumulators, so the result does not depend on which CTA ran it.  gbase = k-steps this CTA has pushed through the ring so far.
  int gbase = 0;
  const int nblocks = a.split.nb * (a.split.nb + 1) / 2, nunits = nblocks * a.ks_eff;
  // counter == null: one unit per CTA in the fixed order "diagonal blocks first", which at Mp = 288 (three full block rows, 2 x 147
  // CTAs) gives every SM exactly one diagonal and one off-diagonal CTA - measured 8 % faster there than the dynamic order, which
  // lets two off-diagonal CTAs share an SM.
  for (int round = 0;; ++round) {
  __syncthreads();                               // the previous unit is finished by every warp (ring drained, ksc_s / unit_s free)
  if (tid == 0) unit_s = a.counter ? atomicAdd(a.counter, 1) : (round == 0 ? (int)blockIdx.x : nunits);
  __syncthreads();
  const int unit = unit_s;
  if (unit >= nunits) break;
  int bi = 0, bj = 0;
  const int ks = unit % a.ks_eff;
  {
    const int nb = a.split.nb;
    int rank = unit / a.ks_eff;
    bool found = false;
    const int order_dyn[4] = {1, 2, 0, 3};       // block types by falling cost
    const int order_fix[4] = {0, 3, 1, 2};       // diagonal blocks first
    for (int t = 0; t < 4 && !found; ++t)
      for (int i = 0; i < nb && !found; ++i)
        for (int j = 0; j <= i && !found; ++j)
          if (a.split.type(i, j) == (a.counter ? order_dyn[t] : order_fix[t])) {
            if (rank == 0) { bi = i; bj = j; found = true; } else --rank;
          }
  }
  const bool diag = bi == bj;
  const int64_t range = ((a.chunk + a.ks_eff - 1) / a.ks_eff + SYRK_BK - 1) / SYRK_BK * SYRK_BK;
  const int64_t p0 = (int64_t)ks * range;
  int64_t p1 = p0 + range;
  if (p1 > a.npts) p1 = a.npts;
  const int kt = p1 > p0 ? (int)((p1 - p0 + SYRK_BK - 1) / SYRK_BK) : 0;
  if (a.zs_out && diag && tid < SYRK_B) ksc_s[tid] = (bi * SYRK_B + tid) < a.Mp ? a.kscale[bi * SYRK_B + tid] : 0.0;
  __syncthreads();

  // producer: 6 chunks of 16 bytes per thread and stage (Zi rows r, r+32, r+64; same for Zj) + weights;
  // a diagonal block reads its single operand from the Zi half only
  const int z_r = tid >> 3, z_kc = (tid & 7) * 2;
  const double* pz[6];
  bool okz[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int row = (i < 3 ? bi : bj) * SYRK_B + z_r + (i % 3) * 32;
    okz[i] = row < a.Mp && !(diag && i >= 3);
    pz[i] = a.Z + (int64_t)(okz[i] ? row : 0) * a.ldz + p0 + z_kc;
  }
  const double* pw = ((tid >> 3) & 1 ? a.wp : a.wq) + p0 + z_kc;   // threads 0..15
  int pit = 0;
  auto produce = [&]() {
    if (pit < kt) {
      const int gi = gbase + pit, st = gi % SYRK_STAGES;
      if (gi >= SYRK_STAGES) mbar_wait(&bar_empty[st], ((gi / SYRK_STAGES) - 1) & 1);
      double* zi = smem + st * SYRK_STAGE_DBL;
      const bool inr = (p0 + (int64_t)pit * SYRK_BK + z_kc) < p1;
#pragma unroll
      for (int i = 0; i < 6; ++i)
        if (!(diag && i >= 3)) cp_async16(zi + (i / 3) * SYRK_B * SYRK_ZS + (z_r + (i % 3) * 32) * SYRK_ZS + z_kc, pz[i], okz[i] && inr);
      if (tid < 16) cp_async16(zi + 2 * SYRK_B * SYRK_ZS + ((tid >> 3) & 1) * SYRK_BK + z_kc, pw, inr);
      cp_async_mbar_arrive(&bar_full[st]);
      ++pit;
#pragma unroll
      for (int i = 0; i < 6; ++i) pz[i] += SYRK_BK;
      pw += SYRK_BK;
    }
  };
#pragma unroll
  for (int s = 0; s < SYRK_STAGES - 1; ++s) produce();

  double* gp = a.Gpart + (int64_t)(a.split.gslot(bi, bj) + ks) * (SYRK_B * SYRK_B);
  if (!diag) {
    double acc[3][6][2];
    bool ract[3];   // sub-tiles whose rows lie beyond Mp (a partial last block row) are skipped by every warp alike
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      ract[r] = (bi * SYRK_B + r * 32 + wm * 8) < a.Mp;
#pragma unroll
      for (int c = 0; c < 6; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;
    }
    for (int kb = 0; kb < kt; ++kb) {
      const int st = (gbase + kb) % SYRK_STAGES;
      mbar_wait(&bar_full[st], ((gbase + kb) / SYRK_STAGES) & 1);
      const double* base = smem + st * SYRK_STAGE_DBL;
      const double* zi = base + (wm * 8 + g) * SYRK_ZS + t4;   // sub-tile r = rows r * 32 + wm * 8 (interleaved, see gemm.cu)
      const double* zj = base + SYRK_B * SYRK_ZS + (wn * 48 + g) * SYRK_ZS + t4;
      const double* wqs = base + 2 * SYRK_B * SYRK_ZS;
#pragma unroll
      for (int k4 = 0; k4 < SYRK_BK / 4; ++k4) {
        double af[3], bf[6];
        const double w = wqs[k4 * 4 + t4];
#pragma unroll
        for (int r = 0; r < 3; ++r) af[r] = zi[r * 32 * SYRK_ZS + k4 * 4] * w;   // the weight on the three A fragments, not the six B
#pragma unroll
        for (int c = 0; c < 6; ++c) bf[c] = zj[c * 8 * SYRK_ZS + k4 * 4];
#pragma unroll
        for (int r = 0; r < 3; ++r)
          if (ract[r]) {
#pragma unroll
            for (int c = 0; c < 6; ++c) dmma884(acc[r][c][0], acc[r][c][1], af[r], bf[c]);
          }
        if (k4 == 0) produce();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_empty[st]);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int row = r * 32 + wm * 8 + g;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const int col = wn * 48 + c * 8 + 2 * t4;
        double2* dst = reinterpret_cast<double2*>(gp + row * SYRK_B + col);
        double2 v = make_double2(acc[r][c][0], acc[r][c][1]);
        if (a.accumulate) { double2 o = *dst; v.x += o.x; v.y += o.y; }
        *dst = v;
      }
    }
    gbase += kt;
    continue;
  }

  // ---- diagonal block ----
  double acc[12][2];
#pragma unroll
  for (int c = 0; c < 12; ++c) acc[c][0] = acc[c][1] = 0.0;
  // bz = Z wp on the FP64 vector pipe: thread = one point of the k-step x 6 of the block's 96 rows (16 x 16 threads), summed over the
  // 16 point lanes once at the end.  As an extra tensor-core column (one m8n8k4 tile per tile row with seven of its eight columns
  // unused) it cost 12 of a diagonal CTA's 90 tiles per step; here it is 6 fused multiply-adds per thread.
  double bzacc[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) bzacc[i] = 0.0;
  const int bz_pt = tid & 15, bz_r0 = (tid >> 4) * 6;
  for (int kb = 0; kb < kt; ++kb) {
    const int st = (gbase + kb) % SYRK_STAGES;
    mbar_wait(&bar_full[st], ((gbase + kb) / SYRK_STAGES) & 1);
    const double* zs = smem + st * SYRK_STAGE_DBL;
    const double* wqs = zs + 2 * SYRK_B * SYRK_ZS;
    produce();
    if (a.zs_out && warp >= 5) {   // warps 5..7 (the lightest tensor-core loads): 96 (point, 16-feature chunk) items per k-step
      const int item = (warp - 5) * 32 + lane, pt = item & 15, c = item >> 4;
      const int64_t n = p0 + (int64_t)kb * SYRK_BK + pt;
      const int krow = bi * (SYRK_B / 16) + c;                 // 16-feature chunk of the panel
      if (n < p1 && krow * 16 < a.Mp)
        syrk_slice_step(zs, ksc_s, c, pt,
                        a.zs_out + ((n >> 7) * (int64_t)(a.Mp / 32) + (krow >> 1)) * GI8_A_STAGE + (krow & 1) * 2048 + (int)(n & 127) * 16);
    }
    {
      const double wpv = wqs[SYRK_BK + bz_pt];
#pragma unroll
      for (int i = 0; i < 6; ++i) bzacc[i] = fma(zs[(bz_r0 + i) * SYRK_ZS + bz_pt], wpv, bzacc[i]);
    }
    switch (warp) {
      case 0: syrk_diag_step<11, -1>(zs, wqs, acc, g, t4); break;
      case 1: syrk_diag_step<10, -1>(zs, wqs, acc, g, t4); break;
      case 2: syrk_diag_step<9, -1>(zs, wqs, acc, g, t4); break;
      case 3: syrk_diag_step<8, -1>(zs, wqs, acc, g, t4); break;
      case 4: syrk_diag_step<7, 0>(zs, wqs, acc, g, t4); break;
      case 5: syrk_diag_step<6, 1>(zs, wqs, acc, g, t4); break;
      case 6: syrk_diag_step<5, 2>(zs, wqs, acc, g, t4); break;
      default: syrk_diag_step<4, 3>(zs, wqs, acc, g, t4); break;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar_empty[st]);
  }
  const int ra = warp < 4 ? 11 - warp : 11 - warp, rb = warp < 4 ? -1 : warp - 4;
  double* bzp = a.bzpart + (int64_t)(a.split.bslot(bi) + ks) * SYRK_B;
#pragma unroll
  for (int c = 0; c < 12; ++c) {
    int row, col;
    if (c <= ra) { row = ra * 8 + g; col = c * 8 + 2 * t4; }
    else if (rb >= 0 && c - ra - 1 <= rb) { row = rb * 8 + g; col = (c - ra - 1) * 8 + 2 * t4; }
    else continue;
    double2* dst = reinterpret_cast<double2*>(gp + row * SYRK_B + col);
    double2 v = make_double2(acc[c][0], acc[c][1]);
    if (a.accumulate) { double2 o = *dst; v.x += o.x; v.y += o.y; }
    *dst = v;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {   // over the 16 point lanes (a half warp), fixed order
    double v = bzacc[i];
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    if (bz_pt == 0) {
      double* d0 = bzp + bz_r0 + i;
      *d0 = a.accumulate ? *d0 + v : v;
    }
  }
  gbase += kt;
  }   // next unit
}

// ------------------------------------------------------------------------------------------------
// K_E  dVhat = dT Xhat^T  (VJP of spherical_harmonics.py:308) on the FP64 tensor cores.  A warp owns 32
// feature rows and a range of points; A fragments (dT) and B fragments (Xhat) are read straight from
// global memory (32-byte segments per row), 8 accumulator tiles per warp.  dVpart[ks][Mp][16].
// ------------------------------------------------------------------------------------------------
struct DvhatArgs {
  const double* dT;
  int64_t ldz;
  const double* XH;   // [npts][dimp]
  int dimp;
  int Mp;
  int64_t npts;
  double* dVpart;     // [KS][Mp][CT * 8], CT = ceil(dimp / 8)
  int KS;
  int64_t range;      // points per split (multiple of 16)
  int accumulate;
};

template <int CT>
__global__ void __launch_bounds__(128) dvhat_dmma_kernel(DvhatArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int nrg = a.Mp / 32;
  const int wid = blockIdx.x * 4 + warp;
  const int rg = wid % nrg, ks = wid / nrg;
  if (ks >= a.KS) return;
  const int row0 = rg * 32;
  const int64_t p0 = (int64_t)ks * a.range;
  int64_t p1 = p0 + a.range;
  if (p1 > a.npts) p1 = a.npts;
  double acc[4][CT][2];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < CT; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;
  constexpr int KU = CT <= 2 ? 4 : 2;  // k4 steps batched per iteration (loads in flight)
  for (int64_t n0 = p0; n0 < p1; n0 += 4 * KU) {
    double af[KU][4], bf[KU][CT];
#pragma unroll
    for (int k4 = 0; k4 < KU; ++k4) {
      const int64_t n = n0 + k4 * 4 + t4;
      const bool ok = n < p1;
#pragma unroll
      for (int r = 0; r < 4; ++r) af[k4][r] = ok ? a.dT[(int64_t)(row0 + r * 8 + g) * a.ldz + n] : 0.0;
#pragma unroll
      for (int c = 0; c < CT; ++c) bf[k4][c] = (ok && (c * 8 + g) < a.dimp) ? a.XH[n * a.dimp + c * 8 + g] : 0.0;
    }
#pragma unroll
    for (int k4 = 0; k4 < KU; ++k4)
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < CT; ++c) dmma884(acc[r][c][0], acc[r][c][1], af[k4][r], bf[k4][c]);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < CT; ++c) {
      double2* dst = reinterpret_cast<double2*>(a.dVpart + ((int64_t)ks * a.Mp + row0 + r * 8 + g) * (CT * 8) + c * 8 + 2 * t4);
      double2 v = make_double2(acc[r][c][0], acc[r][c][1]);
      if (a.accumulate) { double2 o = *dst; v.x += o.x; v.y += o.y; }
      *dst = v;
    }
}

}  // namespace gpfy
