// N-independent inducing-basis algebra on the device: under phase truncation the reference
// re-normalises the fundamental points, rebuilds the per-degree Gram matrices and their Cholesky factors every step
// (spherical_harmonics.py:194-239, 265-288) and differentiates through all of it.  One CTA per degree; everything
// for a degree lives in shared memory (m_l <= 64); for 64 < m_l <= 256 the four m x m matrices live in caller-provided global scratch
// (L2-resident; O(m^3) once per step either way).
//   fwd : Vhat_l = V_l / |V_l| (rows; only in learned mode),  B_l = c_l C_l^alpha(Vhat_l Vhat_l^T) + 1e-16 I,
//         L_l = chol(B_l)                                      c_l = alpha / (alpha + l)
//   vjp : (dVhat_l, dL_l) -> dV_l :  dB = sym(L^-T Phi(L^T dL) L^-1)   (Phi: lower triangle, halved diagonal; the
//         symmetrisation is jnp.linalg.cholesky's),  dG = dB * c_l * 2 alpha C_{l-1}^{alpha+1}(G)  (gegenbauer.py:73),
//         dVhat += (dG + dG^T) Vhat,  dV = (dVhat - Vhat (Vhat . dVhat)) / |V|
#pragma once
#include "svgp_kernels.cuh"

namespace gpfy {

constexpr int BASIS_MAX_M = 64, BASIS_MAX_M_WS = 256;

struct BasisArgs {
  const double* V;      // [M][dim] raw parameters
  double* vhat;         // [M][dim]
  double* lcat;         // [sum m_l^2]
  const double* dvhat;  // [M][dim]  (vjp)
  const double* dlcat;  // [sum m_l^2] (vjp)
  double* dV;           // [M][dim]  (vjp)
  int dim, normalise, vjp;
  double alpha;
  DegreeTable deg;
  RecCoef rc;
  double* scratch = nullptr;   // [4 * sum m_l^2] or null (all matrices in shared memory)
};

static inline int basis_smem_bytes(int m, int dim, bool scratch = false) { return ((scratch ? 0 : 4 * m * m) + 2 * m * dim + 2 * m) * 8; }

__global__ void __launch_bounds__(256) inducing_basis_kernel(BasisArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int l = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  int off = 0;
  int64_t loff = 0;
  for (int k = 0; k < l; ++k) { off += a.deg.m[k]; loff += (int64_t)a.deg.m[k] * a.deg.m[k]; }
  const int m = a.deg.m[l], dim = a.dim;
  double* G = a.scratch ? a.scratch + 4 * loff : sm;   // [m][m] Gram Vhat Vhat^T
  double* L = G + m * m;          // [m][m]
  double* T1 = L + m * m;         // [m][m] scratch
  double* T2 = T1 + m * m;        // [m][m] scratch
  double* vh = a.scratch ? sm : T2 + m * m;        // [m][dim]
  double* dvh = vh + m * dim;     // [m][dim]
  double* nrm = dvh + m * dim;    // [m]
  double* dots = nrm + m;         // [m]
  const double c = a.alpha / (a.alpha + (double)l);

  for (int i = tid; i < m; i += nt) {
    double s = 0.0;
    for (int k = 0; k < dim; ++k) { const double v = a.V[(off + i) * dim + k]; s += v * v; }
    nrm[i] = a.normalise ? sqrt(s) : 1.0;
  }
  __syncthreads();
  for (int e = tid; e < m * dim; e += nt) {
    const double v = a.V[off * dim + e] / nrm[e / dim];
    vh[e] = v;
    if (!a.vjp) a.vhat[off * dim + e] = v;
  }
  __syncthreads();
  for (int e = tid; e < m * m; e += nt) {
    const int i = e / m, j = e % m;
    double s = 0.0;
    for (int k = 0; k < dim; ++k) s = fma(vh[i * dim + k], vh[j * dim + k], s);
    G[e] = s;
  }
  __syncthreads();
  if (!a.vjp) {
    // B = c C_l(G) + 1e-16 I, then the Cholesky factor column by column (spherical_harmonics.py:281-285)
    for (int e = tid; e < m * m; e += nt) L[e] = c * geg_eval(a.rc, l, G[e]) + ((e / m == e % m) ? 1e-16 : 0.0);
    __syncthreads();
    for (int j = 0; j < m; ++j) {
      if (tid == 0) L[j * m + j] = sqrt(L[j * m + j]);
      __syncthreads();
      const double d = L[j * m + j];
      for (int i = j + 1 + tid; i < m; i += nt) L[i * m + j] /= d;
      __syncthreads();
      for (int e = tid; e < (m - j - 1) * (m - j - 1); e += nt) {
        const int i = j + 1 + e / (m - j - 1), k = j + 1 + e % (m - j - 1);
        if (k <= i) L[i * m + k] -= L[i * m + j] * L[k * m + j];
      }
      __syncthreads();
    }
    for (int e = tid; e < m * m; e += nt) a.lcat[loff + e] = (e % m <= e / m) ? L[e] : 0.0;
    return;
  }

  // ---- VJP ----
  for (int e = tid; e < m * m; e += nt) L[e] = a.lcat[loff + e];
  for (int e = tid; e < m * dim; e += nt) dvh[e] = a.dvhat[off * dim + e];
  __syncthreads();
  // T1 = Phi(L^T tril(dL))
  for (int e = tid; e < m * m; e += nt) {
    const int i = e / m, j = e % m;
    double s = 0.0;
    if (j <= i) {
      for (int k = i; k < m; ++k) s = fma(L[k * m + i], a.dlcat[loff + k * m + j], s);  // (L^T dL)[i][j], dL lower: k >= j; L^T[i][k] = L[k][i], k >= i
      if (i == j) s *= 0.5;
    }
    T1[e] = s;
  }
  __syncthreads();
  // T2 = Linv: column j by forward substitution (one thread per column)
  for (int j = tid; j < m; j += nt) {
    for (int i = 0; i < m; ++i) {
      double s = (i == j) ? 1.0 : 0.0;
      for (int k = j; k < i; ++k) s -= L[i * m + k] * T2[k * m + j];
      T2[i * m + j] = (i >= j) ? s / L[i * m + i] : 0.0;
    }
  }
  __syncthreads();
  // G2 := T1 Linv  (stored over L, which is no longer needed)
  for (int e = tid; e < m * m; e += nt) {
    const int i = e / m, j = e % m;
    double s = 0.0;
    for (int k = j; k <= i; ++k) s = fma(T1[i * m + k], T2[k * m + j], s);  // T1 lower (k <= i), Linv lower (k >= j)
    L[e] = s;
  }
  __syncthreads();
  // T1 := S = Linv^T (T1 Linv)
  for (int e = tid; e < m * m; e += nt) {
    const int i = e / m, j = e % m;
    double s = 0.0;
    for (int k = i; k < m; ++k) s = fma(T2[k * m + i], L[k * m + j], s);
    T1[e] = s;
  }
  __syncthreads();
  // dG = sym(S) * c * C'_l(G);  T2 := dG + dG^T
  for (int e = tid; e < m * m; e += nt) {
    const int i = e / m, j = e % m;
    const double dB = 0.5 * (T1[i * m + j] + T1[j * m + i]);
    L[e] = dB * c * geg_deriv(a.rc, l, G[e]);
  }
  __syncthreads();
  for (int e = tid; e < m * dim; e += nt) {
    const int i = e / dim, k = e % dim;
    double s = dvh[e];
    for (int j = 0; j < m; ++j) s = fma(L[i * m + j] + L[j * m + i], vh[j * dim + k], s);
    dvh[e] = s;  // total cotangent of Vhat, in place (each thread touches only its own element)
  }
  __syncthreads();
  for (int i = tid; i < m; i += nt) {
    double s = 0.0;
    for (int k = 0; k < dim; ++k) s = fma(vh[i * dim + k], dvh[i * dim + k], s);
    dots[i] = s;
  }
  __syncthreads();
  for (int e = tid; e < m * dim; e += nt) {
    const int i = e / dim;
    a.dV[off * dim + e] = a.normalise ? (dvh[e] - vh[e] * dots[i]) / nrm[i] : dvh[e];
  }
}

}  // namespace gpfy
