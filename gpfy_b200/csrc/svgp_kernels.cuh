// Device kernels of the SVGP hot path (see DESIGN.md for the algorithm and the reference lines each
// step restates).  Notation: Z = un-whitened zonal values C_l^alpha(Vhat xhat) [Mp, ld] (feature-major,
// points contiguous, like the reference's [M, N] arrays); C = W2 Z; dT = cotangent of t = Vhat xhat.
#pragma once
#include "common.cuh"
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
  const double* sw;     // [d] sqrt(weight_variances) (already broadcast in scalar mode)
  const double* sb;     // [1] sqrt(bias_variance)
  const double* vhat_p; // [Mp, DIMP] zero padded
  int Mp, M, dim;
  int64_t npts;
  double* Z;            // [Mp(+pad), ldz]
  int64_t ldz;
  double* XH;           // [npts, DIMP] unit vectors (zero padded), may be null
  double* R;            // [npts] radii, may be null
  DegreeTable deg;
  RecCoef rc;
};

template <int DIMP>
__device__ __forceinline__ double load_point(const double* __restrict__ X, int xcols, int apply, const double* sw,
                                             const double* sb, int dim, int64_t n, double (&xh)[DIMP]) {
  const double* x = X + n * (int64_t)xcols;
  double r = 1.0;
  if (apply) {
    double ss = 0.0;
#pragma unroll
    for (int c = 0; c < DIMP; ++c) {
      double u = 0.0;
      if (c < xcols) u = x[c] * sw[c];       // X * sqrt(w)              spherical.py:228
      else if (c == xcols) u = sb[0];         // bias = sqrt(b)           spherical.py:245-247
      xh[c] = u;
      ss += u * u;
    }
    r = sqrt(ss);                             // spherical.py:248
#pragma unroll
    for (int c = 0; c < DIMP; ++c) xh[c] = xh[c] / r;  // spherical.py:249
  } else {
#pragma unroll
    for (int c = 0; c < DIMP; ++c) xh[c] = (c < dim) ? x[c] : 0.0;
  }
  return r;
}

template <int DIMP>
__global__ void __launch_bounds__(128) zonal_fwd_kernel(ZonalArgs a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ uint64_t bar;
  double* vs = smem;  // [Mp][DIMP]
  stage_params_tma(vs, a.vhat_p, (unsigned)(a.Mp * DIMP * 8), &bar);

  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.npts) {
    // keep the (even-padded) tail column finite: the GEMM / SYRK read 16-byte pairs of points
    if (n < ((a.npts + 1) & ~(int64_t)1))
      for (int i = 0; i < a.Mp; ++i) a.Z[(int64_t)i * a.ldz + n] = 0.0;
    return;
  }
  double xh[DIMP];
  double r = load_point<DIMP>(a.X, a.xcols, a.apply_to_sphere, a.sw, a.sb, a.dim, n, xh);
  if (a.R) a.R[n] = r;
  if (a.XH) {
#pragma unroll
    for (int c = 0; c < DIMP; c += 2) *reinterpret_cast<double2*>(a.XH + n * DIMP + c) = make_double2(xh[c], xh[c + 1]);
  }
  int i = 0;
  for (int l = 0; l < a.deg.F; ++l) {
    const int ml = a.deg.m[l];
    for (int j = 0; j < ml; ++j, ++i) {
      const double* v = vs + i * DIMP;
      double t = 0.0;
#pragma unroll
      for (int c = 0; c < DIMP; c += 2) {
        double2 vv = *reinterpret_cast<const double2*>(v + c);
        t = fma(vv.x, xh[c], t);
        t = fma(vv.y, xh[c + 1], t);
      }
      a.Z[(int64_t)i * a.ldz + n] = geg_eval(a.rc, l, t);
    }
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

struct LikBwdArgs {
  const double* Z;
  double* C;            // [P][Mp][ldz]; C[0] is overwritten with dT
  int64_t ldz;
  int64_t c_stride;     // Mp * ldz
  const double* XH;     // [npts][DIMP]
  const double* R;
  const double* Y;      // [npts][P]
  const double* vhat_p; // [Mp][DIMP]
  const double* mu2;    // [P][Mp]
  const double* vptr;   // [1] kernel variance
  const double* s2ptr;  // [1] noise variance (Gaussian)
  int Mp, M, d, P, order;
  int64_t npts;
  double* wq;           // [P][ldz]
  double* wp;           // [P][ldz]
  double* partials;     // [gridDim.x][npart]
  int npart;
  int accumulate;       // add to existing partials
  int lik_kind;
  DegreeTable deg;
  RecCoef rc;
};

template <int DIMP, int PMAX>
__global__ void __launch_bounds__(128) lik_bwd_kernel(LikBwdArgs a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ uint64_t bar;
  __shared__ double red[4][PART_DW + DIMP];
  double* vs = smem;                      // [Mp][DIMP]
  double* mu2s = smem + a.Mp * DIMP;      // [P][Mp]
  // one TMA transaction covers both panels: the host lays mu2 right behind vhat_p
  stage_params_tma(vs, a.vhat_p, (unsigned)((a.Mp * DIMP + a.P * a.Mp) * 8), &bar);

  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = n < a.npts;
  double part[PART_DW + DIMP];
#pragma unroll
  for (int k = 0; k < PART_DW + DIMP; ++k) part[k] = 0.0;

  if (!valid && n < ((a.npts + 1) & ~(int64_t)1)) {
    for (int p = 0; p < a.P; ++p) { a.wq[p * a.ldz + n] = 0.0; a.wp[p * a.ldz + n] = 0.0; }
  }
  if (valid) {
    LikParams lp;
    lp.kind = a.lik_kind;
    lp.sigma2 = a.lik_kind == GPFY_LIK_GAUSSIAN ? a.s2ptr[0] : 1.0;
    const double v = a.vptr[0];
    const double r = a.R[n];
    double xh[DIMP];
#pragma unroll
    for (int c = 0; c < DIMP; c += 2) {
      double2 t2 = *reinterpret_cast<const double2*>(a.XH + n * DIMP + c);
      xh[c] = t2.x; xh[c + 1] = t2.y;
    }
    double psi[PMAX], mz[PMAX];
#pragma unroll
    for (int p = 0; p < PMAX; ++p) psi[p] = mz[p] = 0.0;
    const double* zc = a.Z + n;
    const double* cc = a.C + n;
#pragma unroll 4
    for (int i = 0; i < a.M; ++i) {
      const double z = zc[(int64_t)i * a.ldz];
#pragma unroll
      for (int p = 0; p < PMAX; ++p)
        if (p < a.P) {
          psi[p] = fma(z, cc[p * a.c_stride + (int64_t)i * a.ldz], psi[p]);
          mz[p] = fma(mu2s[p * a.Mp + i], z, mz[p]);
        }
    }
    // K_diag = v r^(2 order)  (spherical.py:400); its r-derivative
    double kd = v, dkd = 0.0;
    {
      double rp = 1.0;  // r^(2 order - 1)
      for (int k = 0; k < 2 * a.order - 1; ++k) rp *= r;
      if (a.order > 0) { kd = v * rp * r; dkd = 2.0 * a.order * v * rp; }
    }
    const double r2 = r * r;
    double cp_[PMAX], cq_[PMAX];  // r p  and  2 q r^2
    double dr = 0.0;
#pragma unroll
    for (int p = 0; p < PMAX; ++p) {
      cp_[p] = cq_[p] = 0.0;
      if (p < a.P) {
        const double fmu = r * mz[p];
        const double fvar = kd + r2 * psi[p];
        double e, pp, qq, ds2;
        lik_terms(lp, fmu, fvar, a.Y[n * a.P + p], e, pp, qq, ds2);
        part[PART_E] += e;
        part[PART_DS2] += ds2;
        part[PART_DV] += qq * (kd / v);
        a.wq[p * a.ldz + n] = qq * r2;
        a.wp[p * a.ldz + n] = pp * r;
        cp_[p] = r * pp;
        cq_[p] = 2.0 * qq * r2;
        dr += pp * mz[p] + qq * (dkd + 2.0 * r * psi[p]);
      }
    }
    // pass 2
    double dxh[DIMP];
#pragma unroll
    for (int c = 0; c < DIMP; ++c) dxh[c] = 0.0;
    int i = 0;
    for (int l = 0; l < a.deg.F; ++l) {
      const int ml = a.deg.m[l];
      for (int j = 0; j < ml; ++j, ++i) {
        const double* vv = vs + i * DIMP;
        double vr[DIMP];
        double t = 0.0;
#pragma unroll
        for (int c = 0; c < DIMP; c += 2) {
          double2 v2 = *reinterpret_cast<const double2*>(vv + c);
          vr[c] = v2.x; vr[c + 1] = v2.y;
          t = fma(v2.x, xh[c], t);
          t = fma(v2.y, xh[c + 1], t);
        }
        const double zp = geg_deriv(a.rc, l, t);
        double dz = 0.0;
#pragma unroll
        for (int p = 0; p < PMAX; ++p)
          if (p < a.P) dz += cp_[p] * mu2s[p * a.Mp + i] + cq_[p] * a.C[p * a.c_stride + (int64_t)i * a.ldz + n];
        const double dt = dz * zp;
        a.C[(int64_t)i * a.ldz + n] = dt;
#pragma unroll
        for (int c = 0; c < DIMP; ++c) dxh[c] = fma(vr[c], dt, dxh[c]);
      }
    }
    // du = (dxh - xh (xh . dxh)) / r + dr xh      (VJP of spherical.py:248-249)
    double dot = 0.0;
#pragma unroll
    for (int c = 0; c < DIMP; ++c) dot = fma(xh[c], dxh[c], dot);
#pragma unroll
    for (int c = 0; c < DIMP; ++c) {
      const double du = (dxh[c] - xh[c] * dot) / r + dr * xh[c];
      const double u = xh[c] * r;
      if (c < a.d) part[PART_DW + c] = du * u;   // d w_j = sum du_j x_j / (2 sqrt w_j) = sum du_j u_j / (2 w_j)
      else if (c == a.d) part[PART_DB] = du;     // d b   = sum du_dim / (2 sqrt b)
    }
  }
  // block reduction (fixed order => deterministic)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < PART_DW + DIMP; ++k) {
    double s = warp_sum(part[k]);
    if (lane == 0) red[warp][k] = s;
  }
  __syncthreads();
  if (threadIdx.x < a.npart) {
    const int k = threadIdx.x;
    double s = ((red[0][k] + red[1][k]) + (red[2][k] + red[3][k]));
    double* dst = a.partials + (int64_t)blockIdx.x * a.npart + k;
    *dst = a.accumulate ? *dst + s : s;
  }
}

// predict_diag: pass 1 only, stores fmu / fvar [n, P]  (model.py:220-225)
struct PredictArgs {
  const double* Z;
  const double* C;
  int64_t ldz, c_stride;
  const double* R;
  const double* mu2;   // [P][Mp]
  const double* vptr;
  int Mp, M, P, order;
  int64_t npts;
  double* fmu;         // [npts][P]
  double* fvar;
};

template <int PMAX>
__global__ void __launch_bounds__(128) predict_kernel(PredictArgs a) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.npts) return;
  double psi[PMAX], mz[PMAX];
#pragma unroll
  for (int p = 0; p < PMAX; ++p) psi[p] = mz[p] = 0.0;
  const double* zc = a.Z + n;
  const double* cc = a.C + n;
#pragma unroll 4
  for (int i = 0; i < a.M; ++i) {
    const double z = zc[(int64_t)i * a.ldz];
#pragma unroll
    for (int p = 0; p < PMAX; ++p)
      if (p < a.P) {
        psi[p] = fma(z, cc[p * a.c_stride + (int64_t)i * a.ldz], psi[p]);
        mz[p] = fma(__ldg(a.mu2 + p * a.Mp + i), z, mz[p]);
      }
  }
  const double r = a.R[n];
  double kd = a.vptr[0];
  for (int k = 0; k < 2 * a.order; ++k) kd *= r;
#pragma unroll
  for (int p = 0; p < PMAX; ++p)
    if (p < a.P) {
      a.fmu[n * a.P + p] = r * mz[p];
      a.fvar[n * a.P + p] = kd + r * r * psi[p];
    }
}

// ------------------------------------------------------------------------------------------------
// K_D  weighted SYRK on the FP64 tensor cores:  G (lower 96x96 blocks) += Z diag(wq) Z^T  and, on the
// diagonal blocks, bz += Z wp through one extra MMA column.  Split over points: CTA (blk, ks) owns a
// contiguous range of the chunk and accumulates into its own slot of Gpart (no atomics, deterministic).
// ------------------------------------------------------------------------------------------------
constexpr int SYRK_B = 96, SYRK_BK = 16, SYRK_STAGES = 3, SYRK_THREADS = 192;
constexpr int SYRK_ZS = SYRK_BK + 4;
constexpr int SYRK_STAGE_DBL = 2 * SYRK_B * SYRK_ZS + 2 * SYRK_BK;
constexpr int SYRK_SMEM = SYRK_STAGES * SYRK_STAGE_DBL * 8;

struct SyrkArgs {
  const double* Z;
  int64_t ldz;
  int Mp;
  int64_t npts;
  const double* wq;   // [npts]
  const double* wp;   // [npts]
  double* Gpart;      // [KS][nblk][96*96]
  double* bzpart;     // [KS][nb*96]
  int nb, nblk, KS;
  int64_t range;      // points per split (multiple of 16)
  int accumulate;
};

__global__ void __launch_bounds__(SYRK_THREADS, 2) syrk_dmma_kernel(SyrkArgs a) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp % 3, wn = warp / 3;
  const int g = lane >> 2, t4 = lane & 3;
  const int blk = blockIdx.x % a.nblk, ks = blockIdx.x / a.nblk;
  // blk -> (bi >= bj)
  int bi = 0, rem = blk;
  while (rem > bi) { rem -= bi + 1; ++bi; }
  const int bj = rem;
  const bool diag = bi == bj;
  const int64_t p0 = (int64_t)ks * a.range;
  int64_t p1 = p0 + a.range;
  if (p1 > a.npts) p1 = a.npts;
  const int kt = p1 > p0 ? (int)((p1 - p0 + SYRK_BK - 1) / SYRK_BK) : 0;

  auto load_stage = [&](int st, int kb) {
    double* zi = smem + st * SYRK_STAGE_DBL;
    double* zj = zi + SYRK_B * SYRK_ZS;
    double* wqs = zj + SYRK_B * SYRK_ZS;
    const int64_t pt0 = p0 + (int64_t)kb * SYRK_BK;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int c = tid + i * SYRK_THREADS;  // 2 * 768 chunks
      int which = c >= 768;
      int cc = which ? c - 768 : c;
      int r = cc >> 3, kc = (cc & 7) * 2;
      int row = (which ? bj : bi) * SYRK_B + r;
      bool p = row < a.Mp && (pt0 + kc) < p1;
      const double* src = p ? a.Z + (int64_t)row * a.ldz + pt0 + kc : a.Z;
      cp_async16((which ? zj : zi) + r * SYRK_ZS + kc, src, p);
    }
    if (tid < 16) {
      int which = tid >> 3, kc = (tid & 7) * 2;
      bool p = (pt0 + kc) < p1;
      const double* base = which ? a.wp : a.wq;
      cp_async16(wqs + which * SYRK_BK + kc, p ? base + pt0 + kc : base, p);
    }
  };

  double acc[4][6][2];
  double accx[4][2];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    accx[r][0] = accx[r][1] = 0.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;
  }
#pragma unroll
  for (int s = 0; s < SYRK_STAGES - 1; ++s) {
    if (s < kt) load_stage(s, s);
    cp_async_commit();
  }
  for (int kb = 0; kb < kt; ++kb) {
    cp_async_wait<SYRK_STAGES - 2>();
    __syncthreads();
    if (kb + SYRK_STAGES - 1 < kt) load_stage((kb + SYRK_STAGES - 1) % SYRK_STAGES, kb + SYRK_STAGES - 1);
    cp_async_commit();
    const double* base = smem + (kb % SYRK_STAGES) * SYRK_STAGE_DBL;
    const double* zi = base + (wm * 32 + g) * SYRK_ZS + t4;
    const double* zj = base + SYRK_B * SYRK_ZS + (wn * 48 + g) * SYRK_ZS + t4;
    const double* wqs = base + 2 * SYRK_B * SYRK_ZS;
#pragma unroll
    for (int k4 = 0; k4 < SYRK_BK / 4; ++k4) {
      double af[4], bf[6];
      const double w = wqs[k4 * 4 + t4];
#pragma unroll
      for (int r = 0; r < 4; ++r) af[r] = zi[r * 8 * SYRK_ZS + k4 * 4];
#pragma unroll
      for (int c = 0; c < 6; ++c) bf[c] = zj[c * 8 * SYRK_ZS + k4 * 4] * w;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) dmma884(acc[r][c][0], acc[r][c][1], af[r], bf[c]);
      if (diag && wn == 0) {
        const double bx = (g == 0) ? wqs[SYRK_BK + k4 * 4 + t4] : 0.0;
#pragma unroll
        for (int r = 0; r < 4; ++r) dmma884(accx[r][0], accx[r][1], af[r], bx);
      }
    }
  }
  cp_async_wait<0>();

  double* gp = a.Gpart + ((int64_t)ks * a.nblk + blk) * (SYRK_B * SYRK_B);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int row = wm * 32 + r * 8 + g;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const int col = wn * 48 + c * 8 + 2 * t4;
      double2* dst = reinterpret_cast<double2*>(gp + row * SYRK_B + col);
      double2 v = make_double2(acc[r][c][0], acc[r][c][1]);
      if (a.accumulate) { double2 o = *dst; v.x += o.x; v.y += o.y; }
      *dst = v;
    }
    if (diag && wn == 0 && t4 == 0) {
      double* dst = a.bzpart + (int64_t)ks * (a.nb * SYRK_B) + bi * SYRK_B + row;
      *dst = a.accumulate ? *dst + accx[r][0] : accx[r][0];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K_E  dVhat = dT Xhat^T  (VJP of spherical_harmonics.py:308).  A warp owns 4 feature rows and a
// range of points; lanes stride over points, 4 x DIMP accumulators per lane, one warp reduction at
// the end.  dVpart[ks][Mp][DIMP].
// ------------------------------------------------------------------------------------------------
struct DvhatArgs {
  const double* dT;
  int64_t ldz;
  const double* XH;
  int Mp;
  int64_t npts;
  double* dVpart;
  int KS;
  int64_t range;
  int accumulate;
};

template <int DIMP>
__global__ void __launch_bounds__(128) dvhat_kernel(DvhatArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rg = blockIdx.x % (a.Mp / 16), ks = blockIdx.x / (a.Mp / 16);
  const int row0 = rg * 16 + warp * 4;
  const int64_t p0 = (int64_t)ks * a.range;
  int64_t p1 = p0 + a.range;
  if (p1 > a.npts) p1 = a.npts;
  double acc[4][DIMP];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < DIMP; ++c) acc[r][c] = 0.0;
  for (int64_t n = p0 + lane; n < p1; n += 32) {
    double xh[DIMP];
#pragma unroll
    for (int c = 0; c < DIMP; c += 2) {
      double2 t2 = *reinterpret_cast<const double2*>(a.XH + n * DIMP + c);
      xh[c] = t2.x; xh[c + 1] = t2.y;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const double dt = a.dT[(int64_t)(row0 + r) * a.ldz + n];
#pragma unroll
      for (int c = 0; c < DIMP; ++c) acc[r][c] = fma(dt, xh[c], acc[r][c]);
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < DIMP; ++c) {
      double s = warp_sum(acc[r][c]);
      if (lane == 0) {
        double* dst = a.dVpart + ((int64_t)ks * a.Mp + row0 + r) * DIMP + c;
        *dst = a.accumulate ? *dst + s : s;
      }
    }
}

}  // namespace gpfy
