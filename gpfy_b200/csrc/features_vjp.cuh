// Reverse pass of the feature matrix (gpfy_sh_features_vjp): the VJP `jax.grad` builds for
// SphericalHarmonics.polynomial_expansion / Kuf / project_fun (spherical_harmonics.py:290-320, 344-362, 385-387).
//
//   out = diag(scale_l r_n) L^-1 z_n = r_n A z_n,   A = diag(scale) blockdiag(L_l^-1)
// Given the cotangent dPhi [M, N]:
//   Dp      = dPhi * r                     (panel [Mp][ld], padded with zeros)
//   C       = A^T Dp                       (dz = cotangent of the zonal values; FP64 tensor-core GEMM, blocked layout)
//   dr_n    = z_n . C_n / r_n ... folded:  the GEMM's psi epilogue dots C with the Z panel (GemmArgs::psi_B)
//   dT      = C * z'(t)  ->  dVhat, dxhat -> dw, db            (zonal_bwd4 with cq = 1, cp = 0)
//   dA      = Dp Z^T restricted to the degree blocks           (cross_gram_kernel below, FP64 tensor cores, split over N)
//   dL_l    = -tril(L_l^-T (diag(scale) dA)_l^T ... )          (the dl_step kernels of the ELBO path, with dE = dA)
//   dscale_l= sum_{i in l} sum_j dA_ij Linv_ij                  (dd_kernel)
#pragma once
#include "svgp_kernels.cuh"

namespace gpfy {

// Dp[i][n] = dPhi[i][n] * (r ? r[n] : 1) for i < M, n < npts; zero in the even-padding column
__global__ void scale_panel_kernel(const double* dPhi, int64_t ldA, int M, int64_t npts, const double* r, double* Dp, int64_t ldz) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (n < ((npts + 1) & ~(int64_t)1)) Dp[(int64_t)i * ldz + n] = (n < npts) ? dPhi[(int64_t)i * ldA + n] * (r ? r[n] : 1.0) : 0.0;
}

__global__ void fill_kernel(double* x, int64_t n, double v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = v;
}

// dr[n] = mul_r ? psi[n] / r[n] : 0   (psi = z . C with C = A^T (dPhi r), so psi / r = (A z) . dPhi = d out / d r)
// per-row scale vector of A = diag(scale) Linv
__global__ void row_scale_kernel(const double* scale, DegreeTable deg, int M, int Mp, double* dvec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Mp) return;
  double v = 0.0;
  if (i < M) {
    int l = 0, o = 0;
    while (i >= o + deg.m[l]) { o += deg.m[l]; ++l; }
    v = scale ? scale[l] : 1.0;
  }
  dvec[i] = v;
}

__global__ void dr_from_psi_kernel(const double* psi_part, int npsi, int64_t ldz, const double* r, int mul_r, int64_t npts, double* dr) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= npts) return;
  double s = 0.0;
  if (mul_r) {
    for (int k = 0; k < npsi; ++k) s += psi_part[(int64_t)k * ldz + n];
    s /= r[n];
  }
  dr[n] = s;
}

// Block-diagonal cross Gram matrix on the FP64 tensor cores:  part[ks][i][c] = sum_{n in split ks} P[i][n] Q[col0(i) + c][n]
// for the 16 rows of row group rg = i / 16 and a window of XG_W columns starting at col0 (the first column of the degree
// block of the group's first row).  A warp owns (row group, split); fragments are read straight from the L2-resident panels.
constexpr int XG_W = 128;
struct CrossGramArgs {
  const double* P;     // [Mp][ld]
  const double* Q;     // [Mp][ld]
  int64_t ld;
  int Mp;
  int64_t npts;
  int col0[72];        // [Mp / 16] first window column of each row group
  double* part;        // [KS][Mp][XG_W]
  int KS;
  int64_t range;       // points per split (multiple of 4)
  int accumulate;
};

__global__ void __launch_bounds__(128) cross_gram_kernel(CrossGramArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int nrg = a.Mp / 16;
  const int wid = blockIdx.x * 4 + warp;
  const int rg = wid % nrg, ks = wid / nrg;
  if (ks >= a.KS) return;
  const int row0 = rg * 16, c0 = a.col0[rg];
  const int64_t p0 = (int64_t)ks * a.range;
  int64_t p1 = p0 + a.range;
  if (p1 > a.npts) p1 = a.npts;
  double acc[2][XG_W / 8][2];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < XG_W / 8; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;
  for (int64_t n0 = p0; n0 < p1; n0 += 4) {
    const int64_t n = n0 + t4;
    const bool ok = n < p1;
    double af[2], bf[XG_W / 8];
#pragma unroll
    for (int r = 0; r < 2; ++r) af[r] = ok ? a.P[(int64_t)(row0 + r * 8 + g) * a.ld + n] : 0.0;
#pragma unroll
    for (int c = 0; c < XG_W / 8; ++c) {
      const int col = c0 + c * 8 + g;
      bf[c] = (ok && col < a.Mp) ? a.Q[(int64_t)col * a.ld + n] : 0.0;
    }
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < XG_W / 8; ++c) dmma884(acc[r][c][0], acc[r][c][1], af[r], bf[c]);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < XG_W / 8; ++c) {
      double2* dst = reinterpret_cast<double2*>(a.part + ((int64_t)ks * a.Mp + row0 + r * 8 + g) * XG_W + c * 8 + 2 * t4);
      double2 v = make_double2(acc[r][c][0], acc[r][c][1]);
      if (a.accumulate) { const double2 o = *dst; v.x += o.x; v.y += o.y; }
      *dst = v;
    }
}

// dA [Mp][Mp] (dense, zero outside the degree blocks) = sum over splits of the windows
struct XgCols { int col0[72]; };
__global__ void cross_gram_reduce_kernel(const double* part, XgCols xc, int KS, int Mp, int M, DegreeTable deg, double* dA) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)Mp * Mp) return;
  const int i = (int)(e / Mp), j = (int)(e % Mp);
  double s = 0.0;
  if (i < M && j < M) {
    int li = 0, o = 0;
    while (i >= o + deg.m[li]) { o += deg.m[li]; ++li; }
    if (j >= o && j < o + deg.m[li]) {            // same degree block
      const int c = j - xc.col0[i / 16];
      if (c >= 0 && c < XG_W)
        for (int ks = 0; ks < KS; ++ks) s += part[((int64_t)ks * Mp + i) * XG_W + c];
    }
  }
  dA[e] = s;
}

struct FeatVjpFinalArgs {
  const double* partials;   // [npart] summed partial sums (PART_DB, PART_DW..)
  const double* dd;         // [M] sum_j dA_ij Linv_ij
  const double* scale;      // [F] or null
  const double *w, *b;
  int d, scalar_w, proj, nw, apply_to_sphere;
  DegreeTable deg;
  double *dw, *db, *dscale;
};
__global__ void features_vjp_finalize_kernel(FeatVjpFinalArgs a) {
  const int lane = threadIdx.x;
  if (a.apply_to_sphere) {
    if (lane == 0 && a.db) a.db[0] = a.partials[PART_DB] / (2.0 * sqrt(a.b[0]));
    if (a.dw) {
      if (a.scalar_w) {
        double s = 0.0;
        for (int j = lane; j < a.d; j += 32) s += a.partials[PART_DW + j];
        s = warp_sum(s);
        if (lane == 0) a.dw[0] = s / (2.0 * a.w[0]);
      } else {
        for (int j = lane; j < a.d; j += 32) a.dw[j] = a.partials[PART_DW + j] / (2.0 * a.w[j]);
      }
    }
  }
  if (a.dscale) {
    int base = 0;
    for (int l = 0; l < a.deg.F; ++l) {
      double s = 0.0;
      for (int j = lane; j < a.deg.m[l]; j += 32) s += a.dd[base + j];
      s = warp_sum(s);
      if (lane == 0) a.dscale[l] = s;
      base += a.deg.m[l];
    }
  }
}

}  // namespace gpfy
