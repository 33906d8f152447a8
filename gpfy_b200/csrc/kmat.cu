// Kernel-matrix side of the model: the dense N x N2 prior covariance of the spherical kernels,
// their shape functions, and the N x N projected covariance of q(u) that `GP.predict / GP.cov` add to it.
//   spherical.py:318-382 (K, K2), :438-455 / :479-500 / :589-611 (shape functions), variational.py:330-351 / :443-464
//   (project_variance), model.py:129-131, 231-259 (conditional covariance, predict).
// Output is O(N N2) doubles, so N is small here (thousands); the kernels are tiled for coalesced 128-byte row segments
// and FP64 tensor-core contractions, but this is not the 10 M-row path and carries no pipeline machinery.
#include "common.cuh"

namespace gpfy {

#define KM_LAUNCH(kernel, grid, block, smem, st, ...) \
  do {                                                \
    kernel<<<grid, block, smem, st>>>(__VA_ARGS__);   \
    GPFY_COUNT_LAUNCH();                              \
    GPFY_LAUNCH_OK();                                 \
  } while (0)

// ------------------------------------------------------------------------------------------------
// shape functions
// ------------------------------------------------------------------------------------------------
struct ShapeParams {
  int kind, depth, order, num_levels, grid_resolution;
  double alpha;
  double coef[GPFY_MAX_SHAPE_LEVELS];   // PolynomialDecay: lambda_n (n + alpha) / alpha (spherical.py:607-608)
  double c1[GPFY_MAX_SHAPE_LEVELS];     // C_n = c1[n] x C_{n-1} - c2[n] C_{n-2}: 2 (n + alpha - 1) / n and
  double c2[GPFY_MAX_SHAPE_LEVELS];     // (n + 2 alpha - 2) / n of gegenbauer.py:49, divided on the host
};

constexpr double kPi = 3.14159265358979323846;
constexpr double kInvPi = 0.31830988618379067154;   // x * (1 / pi) instead of x / pi: one rounding apart, no FP64 divide chain

// kappa_0 and kappa_1 share acos and sqrt (spherical.py:299-303)
__device__ __forceinline__ void kappa01(double u, double& k0, double& k1) {
  const double ac = kPi - acos(u);
  k0 = ac * kInvPi;
  k1 = (u * ac + sqrt(1.0 - u * u)) * kInvPi;
}

__device__ __forceinline__ double kappa(double u, int order) {
  const double ac = kPi - acos(u);
  if (order == 0) return ac * kInvPi;
  const double sq = sqrt(1.0 - u * u);
  if (order == 1) return (u * ac + sq) * kInvPi;
  return ((1.0 + 2.0 * (u * u)) / 3 * ac + sq) * kInvPi;   // spherical.py:305-309 as written
}

// sum_n coef_n C_n^alpha on the lookup grid: the reference tabulates C_n on linspace(-1, 1, G) and interpolates linearly
// (gegenbauer.py:111-119, 152-167).  The two bracketing nodes are re-evaluated here with the same recurrence instead of
// being fetched from an 8.8 MB table: 2 x 3 FP64 operations per level against two dependent uncoalesced loads.
__device__ __forceinline__ double polydecay_shape(const ShapeParams& sp, double x) {
  const int G = sp.grid_resolution;
  double xi = (x + 1.0) / 2.0 * (double)(G - 1);
  xi = fmin(fmax(xi, 0.0), (double)(G - 1));
  const double lo = floor(xi);
  const double hi = fmin(lo + 1.0, (double)(G - 1));
  const double t = xi - lo;
  const double step = 2.0 / (double)(G - 1);
  const double xa = lo == (double)(G - 1) ? 1.0 : lo * step + -1.0;   // numpy.linspace: arange * step + start, last = stop
  const double xb = hi == (double)(G - 1) ? 1.0 : hi * step + -1.0;
  const double al = sp.alpha;
  double a0 = 1.0, b0 = 1.0;                 // C_0
  double acc = sp.coef[0] * ((1.0 - t) * a0 + t * b0);
  if (sp.num_levels == 1) return acc;
  double a1 = 2.0 * al * xa, b1 = 2.0 * al * xb;   // C_1
  acc += sp.coef[1] * ((1.0 - t) * a1 + t * b1);
  for (int n = 2; n < sp.num_levels; ++n) {
    const double a2 = sp.c1[n] * xa * a1 - sp.c2[n] * a0;   // gegenbauer.py:49 with the division folded into the table
    const double b2 = sp.c1[n] * xb * b1 - sp.c2[n] * b0;
    acc += sp.coef[n] * ((1.0 - t) * a2 + t * b2);
    a0 = a1; a1 = a2;
    b0 = b1; b1 = b2;
  }
  return acc;
}

template <int KIND>
__device__ __forceinline__ double shape_eval(const ShapeParams& sp, double x) {
  x = fmin(fmax(x, -1.0), 1.0);   // _squash (spherical.py:265-277)
  if (KIND == GPFY_SHAPE_NTK) {   // spherical.py:493-500
    double a = x, y = x;
    for (int i = 0; i < sp.depth; ++i) {
      double k0, k1;
      kappa01(a, k0, k1);
      y = y * k0 + k1;
      a = k1;
    }
    return y / (double)(sp.depth + 1);
  }
  if (KIND == GPFY_SHAPE_ARCCOSINE) {   // spherical.py:452-455
    double y = x;
    for (int i = 0; i < sp.depth; ++i) y = kappa(y, sp.order);
    return y;
  }
  return polydecay_shape(sp, x);
}

#define KM_KIND_SWITCH(kind, ...)                                                        \
  switch (kind) {                                                                        \
    case GPFY_SHAPE_ARCCOSINE: { constexpr int KIND = GPFY_SHAPE_ARCCOSINE; __VA_ARGS__; } break; \
    case GPFY_SHAPE_NTK: { constexpr int KIND = GPFY_SHAPE_NTK; __VA_ARGS__; } break;    \
    default: { constexpr int KIND = GPFY_SHAPE_POLYDECAY; __VA_ARGS__; } break;          \
  }

template <int KIND>
__global__ void shape_function_kernel(const __grid_constant__ ShapeParams sp, const double* __restrict__ x, int64_t n,
                                      double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = shape_eval<KIND>(sp, x[i]);
}

// ------------------------------------------------------------------------------------------------
// K[i, j] = variance * shape(<xhat_i, xhat2_j>) * (r_i r2_j)^order       (spherical.py:336-352, 373-382)
// One CTA per 64 x 64 tile, 256 threads.  Phase 1: each thread forms 4 x 4 cosines (plain FMA from shared memory, the
// contraction is only dim long) and parks them in a shared 64 x 64 tile.  Phase 2: ONE rolled instance of the shape
// function walks the tile row-major, so a warp evaluates 32 neighbouring columns and stores a full 256-byte row segment;
// keeping a single copy of the acos / sqrt chains keeps the kernel a few hundred instructions instead of sixteen inlined
// copies (the first version: 6 k instructions, instruction-fetch bound).
// ------------------------------------------------------------------------------------------------
constexpr int KT = 64;

template <int KIND>
__global__ void __launch_bounds__(256, 4) kernel_matrix_kernel(const __grid_constant__ ShapeParams sp, const double* __restrict__ xs1,
                                                            const double* __restrict__ r1, int n1, const double* __restrict__ xs2,
                                                            const double* __restrict__ r2, int n2, int dim,
                                                            const double* __restrict__ variance, int kernel_order, int unit_diagonal,
                                                            double* __restrict__ K) {
  extern __shared__ double sm[];
  const int ld = dim | 1;   // odd row pitch: the 16 rows a half-warp reads fall on 16 distinct bank pairs
  double* a = sm;                  // [64][ld]
  double* b = a + KT * ld;         // [64][ld]
  double* ct = b + KT * ld;        // [64][64] cosines
  double* ra = ct + KT * KT;       // [64] radii of the tile rows
  double* rb = ra + KT;            // [64] radii of the tile columns
  const int i0 = blockIdx.y * KT, j0 = blockIdx.x * KT;
  for (int e = threadIdx.x; e < KT * dim; e += 256) {
    const int row = e / dim, c = e - row * dim;
    a[row * ld + c] = i0 + row < n1 ? xs1[(int64_t)(i0 + row) * dim + c] : 0.0;
    b[row * ld + c] = j0 + row < n2 ? xs2[(int64_t)(j0 + row) * dim + c] : 0.0;
  }
  if (threadIdx.x < KT) ra[threadIdx.x] = i0 + threadIdx.x < n1 ? r1[i0 + threadIdx.x] : 0.0;
  else if (threadIdx.x < 2 * KT) rb[threadIdx.x - KT] = j0 + threadIdx.x - KT < n2 ? r2[j0 + threadIdx.x - KT] : 0.0;
  __syncthreads();
  {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
    for (int c = 0; c < dim; ++c) {
      double av[4], bv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) av[u] = a[(ty + 16 * u) * ld + c];
#pragma unroll
      for (int v = 0; v < 4; ++v) bv[v] = b[(tx + 16 * v) * ld + c];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fma(av[u], bv[v], acc[u][v]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) ct[(ty + 16 * u) * KT + tx + 16 * v] = acc[u][v];
  }
  __syncthreads();
  const double var = variance[0];
#pragma unroll 1
  for (int e = threadIdx.x; e < KT * KT; e += 256) {
    const int row = e >> 6, col = e & 63;
    const int i = i0 + row, j = j0 + col;
    if (i >= n1 || j >= n2) continue;
    double c = ct[e];
    if (unit_diagonal && i == j) c = 1.0;   // spherical.py:340-341
    const double k = shape_eval<KIND>(sp, c);
    const double rr = ra[row] * rb[col];
    double rp = 1.0;
    for (int o = 0; o < kernel_order; ++o) rp *= rr;
    K[(int64_t)i * n2 + j] = var * k * rp;
  }
}

// ------------------------------------------------------------------------------------------------
// out[i, j] = (base ? base[i, j] : 0) + scale * sum_k A(k, i) B(k, j),   A(k, i) = A[k * sak + i * sai]
// FP64 tensor cores (DMMA m8n8k4): CTA tile 64 x 64, 8 warps as 4 x 2, warp tile 16 x 32, k tile 16.
// Shared rows are padded to 68 doubles: a half-warp fragment load touches 4 rows x 4 doubles and the 8-bank row skew
// makes that conflict free.  `base` may alias `out` (each element is read and written by the same lane).
// ------------------------------------------------------------------------------------------------
constexpr int GT = 64, GK = 16, GLD = 68;

__global__ void __launch_bounds__(256) gemm_tn_kernel(int Kd, int I, int J, const double* __restrict__ A, int64_t sak, int64_t sai,
                                                      const double* __restrict__ B, int64_t sbk, int64_t sbj, double scale,
                                                      const double* base, double* out, int64_t ldo) {
  __shared__ double As[GK * GLD], Bs[GK * GLD];
  const int i0 = blockIdx.y * GT, j0 = blockIdx.x * GT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const int wi = (warp >> 1) * 16, wj = (warp & 1) * 32;
  double acc[2][4][2];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v][0] = acc[u][v][1] = 0.0;
  for (int k0 = 0; k0 < Kd; k0 += GK) {
    for (int e = threadIdx.x; e < GK * GT; e += 256) {
      // walk the contiguous direction of each operand with consecutive threads
      int ka, ia, kb, jb;
      if (sai == 1) { ka = e >> 6; ia = e & 63; } else { ka = e & 15; ia = e >> 4; }
      if (sbj == 1) { kb = e >> 6; jb = e & 63; } else { kb = e & 15; jb = e >> 4; }
      As[ka * GLD + ia] = (k0 + ka < Kd && i0 + ia < I) ? A[(int64_t)(k0 + ka) * sak + (int64_t)(i0 + ia) * sai] : 0.0;
      Bs[kb * GLD + jb] = (k0 + kb < Kd && j0 + jb < J) ? B[(int64_t)(k0 + kb) * sbk + (int64_t)(j0 + jb) * sbj] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < GK; ks += 4) {
      double af[2], bf[4];
#pragma unroll
      for (int u = 0; u < 2; ++u) af[u] = As[(ks + t4) * GLD + wi + 8 * u + g];
#pragma unroll
      for (int v = 0; v < 4; ++v) bf[v] = Bs[(ks + t4) * GLD + wj + 8 * v + g];
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) dmma884(acc[u][v][0], acc[u][v][1], af[u], bf[v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int i = i0 + wi + 8 * u + g;
    if (i >= I) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int j = j0 + wj + 8 * v + 2 * t4 + c;
        if (j >= J) continue;
        const int64_t o = (int64_t)i * ldo + j;
        out[o] = (base ? base[o] : 0.0) + scale * acc[u][v][c];
      }
  }
}

static int gemm_tn(cudaStream_t st, int Kd, int I, int J, const double* A, int64_t sak, int64_t sai, const double* B, int64_t sbk,
                   int64_t sbj, double scale, const double* base, double* out, int64_t ldo) {
  dim3 grid((J + GT - 1) / GT, (I + GT - 1) / GT);
  KM_LAUNCH(gemm_tn_kernel, grid, 256, 0, st, Kd, I, J, A, sak, sai, B, sbk, sbj, scale, base, out, ldo);
  return GPFY_OK;
}

static int make_shape_params(const GpfyShapeFn* fn, double alpha, ShapeParams* sp) {
  if (!fn || fn->struct_size != sizeof(GpfyShapeFn)) { set_error("GpfyShapeFn: null or struct_size mismatch"); return GPFY_EINVAL; }
  if (fn->kind < GPFY_SHAPE_ARCCOSINE || fn->kind > GPFY_SHAPE_POLYDECAY) { set_error("unknown shape function kind %d", fn->kind); return GPFY_EINVAL; }
  sp->kind = fn->kind;
  sp->depth = fn->depth;
  sp->order = fn->order;
  sp->num_levels = fn->num_levels;
  sp->grid_resolution = fn->grid_resolution;
  sp->alpha = alpha;
  if (fn->kind == GPFY_SHAPE_POLYDECAY) {
    if (fn->num_levels < 1 || fn->num_levels > GPFY_MAX_SHAPE_LEVELS) { set_error("truncation level %d outside [1, %d]", fn->num_levels, GPFY_MAX_SHAPE_LEVELS); return GPFY_EUNSUPPORTED; }
    if (fn->grid_resolution < 2) { set_error("grid_resolution must be >= 2"); return GPFY_EINVAL; }
    for (int i = 0; i < fn->num_levels; ++i) {
      sp->coef[i] = fn->coef[i];
      sp->c1[i] = i ? 2.0 * (i + alpha - 1.0) / i : 0.0;
      sp->c2[i] = i ? (i + 2.0 * alpha - 2.0) / i : 0.0;
    }
  } else {
    if (fn->depth < 0) { set_error("negative depth"); return GPFY_EINVAL; }
    if (fn->kind == GPFY_SHAPE_ARCCOSINE && (fn->order < 0 || fn->order > 2)) { set_error("ArcCosine order %d not in {0, 1, 2}", fn->order); return GPFY_EINVAL; }
  }
  return GPFY_OK;
}

}  // namespace gpfy

using namespace gpfy;

extern "C" {

int gpfy_kmat_workspace_bytes(const GpfyDesc* desc, int64_t n1, int64_t n2, size_t* bytes) {
  Shape s;
  GPFY_TRY(make_shape(desc, &s));
  if (!bytes || n1 < 0 || n2 < 0) { set_error("bad argument"); return GPFY_EINVAL; }
  const int64_t sphere = (n1 + n2) * (int64_t)(s.dim + 1);
  const int64_t tmp = (int64_t)s.M * (n1 > n2 ? n1 : n2);
  *bytes = (size_t)round_up64((sphere > tmp ? sphere : tmp) * 8, 256);
  return GPFY_OK;
}

int gpfy_shape_function(void* stream, const GpfyShapeFn* fn, double alpha, const double* x, int64_t n, double* out) {
  ShapeParams sp;
  GPFY_TRY(make_shape_params(fn, alpha, &sp));
  if (n == 0) return GPFY_OK;
  KM_KIND_SWITCH(sp.kind, KM_LAUNCH(shape_function_kernel<KIND>, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, sp, x, n, out));
  return GPFY_OK;
}

int gpfy_kernel_matrix(void* stream, const GpfyDesc* desc, const GpfyShapeFn* fn, int64_t n1, const double* X1, int64_t n2,
                       const double* X2, const double* w, const double* b, const double* variance, int unit_diagonal, double* K,
                       void* workspace, size_t workspace_bytes) {
  cudaStream_t st = (cudaStream_t)stream;
  Shape s;
  GPFY_TRY(make_shape(desc, &s));
  ShapeParams sp;
  GPFY_TRY(make_shape_params(fn, desc->alpha, &sp));
  if (n1 < 0 || n2 < 0 || n1 > INT32_MAX || n2 > INT32_MAX) { set_error("row counts out of range"); return GPFY_EINVAL; }
  const bool self = X2 == nullptr;
  if (self) n2 = n1;
  if (n1 == 0 || n2 == 0) return GPFY_OK;
  size_t need;
  GPFY_TRY(gpfy_kmat_workspace_bytes(desc, n1, self ? 0 : n2, &need));
  if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 255)) { set_error("workspace too small or misaligned (%zu needed)", need); return GPFY_EWORKSPACE; }
  double* ws = (double*)workspace;
  double* xs1 = ws;
  double* r1 = xs1 + n1 * s.dim;
  double* xs2 = self ? xs1 : r1 + n1;
  double* r2 = self ? r1 : xs2 + n2 * s.dim;
  GPFY_TRY(gpfy_to_sphere(stream, desc, n1, X1, w, b, xs1, r1));
  if (!self) GPFY_TRY(gpfy_to_sphere(stream, desc, n2, X2, w, b, xs2, r2));
  dim3 grid((unsigned)((n2 + KT - 1) / KT), (unsigned)((n1 + KT - 1) / KT));
  const size_t smem = ((size_t)2 * KT * (s.dim | 1) + KT * KT + 2 * KT) * 8;
  KM_KIND_SWITCH(sp.kind, {
    if (smem > 48 * 1024) GPFY_CUDA_OK(cudaFuncSetAttribute(kernel_matrix_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KM_LAUNCH(kernel_matrix_kernel<KIND>, grid, 256, smem, st, sp, xs1, r1, (int)n1, xs2, r2, (int)n2, s.dim, variance,
              desc->kernel_order, self && unit_diagonal, K);
  });
  return GPFY_OK;
}

int gpfy_project_variance(void* stream, const GpfyDesc* desc, int64_t n, const double* A, const double* sigma, int subtract_identity,
                          const double* base, double* out, void* workspace, size_t workspace_bytes) {
  cudaStream_t st = (cudaStream_t)stream;
  Shape s;
  GPFY_TRY(make_shape(desc, &s));
  if (n < 0 || n > INT32_MAX) { set_error("row count out of range"); return GPFY_EINVAL; }
  if (n == 0) return GPFY_OK;
  size_t need;
  GPFY_TRY(gpfy_kmat_workspace_bytes(desc, n, 0, &need));
  if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 255)) { set_error("workspace too small or misaligned (%zu needed)", need); return GPFY_EWORKSPACE; }
  double* tmp = (double*)workspace;   // [M, N]
  const int M = s.M, N = (int)n;
  for (int q = 0; q < s.P; ++q) {
    const double* Sq = sigma + (int64_t)q * M * M;
    double* Oq = out + (int64_t)q * N * N;
    if (desc->q_kind == GPFY_Q_TRIL) {
      // tmp = L^T A, out = tmp^T tmp (variational.py:346-351)
      GPFY_TRY(gemm_tn(st, M, M, N, Sq, M, 1, A, N, 1, 1.0, nullptr, tmp, N));
      GPFY_TRY(gemm_tn(st, M, N, N, tmp, N, 1, tmp, N, 1, 1.0, base, Oq, N));
    } else {
      // tmp = S A, out = A^T tmp (variational.py:459-464); S is read as stored, symmetry is not assumed
      GPFY_TRY(gemm_tn(st, M, M, N, Sq, 1, M, A, N, 1, 1.0, nullptr, tmp, N));
      GPFY_TRY(gemm_tn(st, M, N, N, A, N, 1, tmp, N, 1, 1.0, base, Oq, N));
    }
    // conditional covariance of the prior part: - A^T A (spherical_harmonics.py:391-393)
    if (subtract_identity) GPFY_TRY(gemm_tn(st, M, N, N, A, N, 1, A, N, 1, -1.0, Oq, Oq, N));
  }
  return GPFY_OK;
}

}  // extern "C"
