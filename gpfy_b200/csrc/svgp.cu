// Host orchestration + C-ABI of the gpfy_b200 hot path.  Stream-ordered, no allocation, no sync.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>

#include "gemm.cuh"
#include "svgp_kernels.cuh"
#include "zonal_bwd3.cuh"
#include "zonal_bwd4.cuh"
#include "zonal_fwd2.cuh"
#include "features3.cuh"
#include "zonal_fwd3.cuh"
#include "basis.cuh"
#include "features_vjp.cuh"
#include "syrk_tf32.cuh"
#include "gemm_i8.cuh"

namespace gpfy {

// ------------------------------------------------------------------------------------------------
// errors / bookkeeping
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launch_count{0};
void count_launch() { g_launch_count.fetch_add(1, std::memory_order_relaxed); }

static std::atomic<int> g_debug_opt[DBG_COUNT];
int debug_option(int which) { return (which >= 0 && which < DBG_COUNT) ? g_debug_opt[which].load(std::memory_order_relaxed) : 0; }

constexpr int MAX_DEVICES = 64;
struct DeviceState { int sms = 0; unsigned configured = 0; };
static std::mutex g_dev_mu;
static DeviceState g_dev[MAX_DEVICES];

int device_sms() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return 148;
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (!g_dev[dev].sms) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    g_dev[dev].sms = sms;
  }
  return g_dev[dev].sms;
}

int device_configure_once(unsigned bit, const void* kernel, int smem_bytes) {
  int dev = 0;
  GPFY_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= MAX_DEVICES) { set_error("device ordinal %d out of range", dev); return GPFY_EUNSUPPORTED; }
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (!(g_dev[dev].configured & bit)) {
    GPFY_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    g_dev[dev].configured |= bit;
  }
  return GPFY_OK;
}

// Optional per-kernel device timing (bench.py's roofline line): when enabled, the per-chunk kernels
// of the ELBO step are bracketed by cudaEvents on the launching stream; gpfy_debug_kernel_time
// synchronises those events and returns the summed durations.  Off by default: zero cost.
enum { TK_ZONAL_FWD = 0, TK_GEMM = 1, TK_LIK = 2, TK_ZONAL_BWD = 3, TK_SYRK = 4, TK_DVHAT = 5, TK_SLICE = 6, TK_COUNT = 7 };
struct KernelTimer {
  bool on = false;
  std::vector<cudaEvent_t> ev[TK_COUNT];   // pairs (start, stop)
  size_t used[TK_COUNT] = {0, 0, 0, 0, 0, 0, 0};
  cudaEvent_t get(int k) {
    if (used[k] == ev[k].size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      ev[k].push_back(e);
    }
    return ev[k][used[k]++];
  }
};
static KernelTimer g_timer;
struct ScopedKernelTime {
  int k;
  cudaStream_t st;
  ScopedKernelTime(int k_, cudaStream_t st_) : k(k_), st(st_) {
    if (g_timer.on) cudaEventRecord(g_timer.get(k), st);
  }
  ~ScopedKernelTime() {
    if (g_timer.on) cudaEventRecord(g_timer.get(k), st);
  }
};
#define GPFY_TIMED(k, st) ScopedKernelTime timed__(k, st)

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int make_shape(const GpfyDesc* desc, Shape* s) {
  if (!desc) { set_error("descriptor is NULL"); return GPFY_EINVAL; }
  if (desc->struct_size != sizeof(GpfyDesc) || desc->abi_version != GPFY_ABI_VERSION) {
    set_error("descriptor size/version mismatch (got %u/%u, want %zu/%d)", desc->struct_size, desc->abi_version,
              sizeof(GpfyDesc), GPFY_ABI_VERSION);
    return GPFY_EINVAL;
  }
  if (desc->input_dim < 1 || desc->num_degrees < 1 || desc->num_degrees > GPFY_MAX_DEGREES || desc->num_outputs < 1) {
    set_error("bad input_dim/num_degrees/num_outputs (%d/%d/%d)", desc->input_dim, desc->num_degrees, desc->num_outputs);
    return GPFY_EINVAL;
  }
  s->d = desc->input_dim;
  s->proj = desc->weight_mode == GPFY_W_PROJECTION ? desc->projection_dim : 0;
  if (desc->weight_mode == GPFY_W_PROJECTION && s->proj < 1) { set_error("projection_dim must be >= 1 for GPFY_W_PROJECTION"); return GPFY_EINVAL; }
  s->sd = s->proj > 0 ? s->proj : s->d;
  s->nw = s->proj > 0 ? s->d * s->proj : (desc->weight_mode == GPFY_W_SCALAR ? 1 : s->d);
  s->dim = s->sd + 1;
  s->dimp = round_up(s->dim, 4);
  s->F = desc->num_degrees;
  s->P = desc->num_outputs;
  s->M = 0;
  s->lsize = 0;
  s->mmax = 0;
  for (int l = 0; l < s->F; ++l) {
    if (desc->m[l] < 1) { set_error("m[%d] = %d must be >= 1", l, desc->m[l]); return GPFY_EINVAL; }
    s->off[l] = s->M;
    s->loff[l] = s->lsize;
    s->M += desc->m[l];
    s->lsize += (int64_t)desc->m[l] * desc->m[l];
    s->mmax = std::max(s->mmax, desc->m[l]);
  }
  s->off[s->F] = s->M;
  s->loff[s->F] = s->lsize;
  s->Mp = round_up(s->M, 32);
  if (s->dimp > 36) { set_error("input_dim %d > 35 is not built (ambient dim padded to %d > 36)", s->d, s->dimp); return GPFY_EUNSUPPORTED; }
  if (s->P > 8) { set_error("num_outputs %d > 8 is not built", s->P); return GPFY_EUNSUPPORTED; }
  if (desc->kernel_order < 0 || desc->kernel_order > 8) { set_error("kernel_order %d out of range", desc->kernel_order); return GPFY_EINVAL; }
  if (desc->likelihood != GPFY_LIK_GAUSSIAN && desc->likelihood != GPFY_LIK_BERNOULLI) { set_error("unknown likelihood %d", desc->likelihood); return GPFY_EINVAL; }
  if (desc->q_kind != GPFY_Q_TRIL && desc->q_kind != GPFY_Q_FULLCOV) { set_error("unknown q_kind %d", desc->q_kind); return GPFY_EINVAL; }
  if (desc->weight_mode < GPFY_W_ARD || desc->weight_mode > GPFY_W_PROJECTION) { set_error("unknown weight_mode %d", desc->weight_mode); return GPFY_EINVAL; }
  if (desc->precision != GPFY_PREC_F64 && desc->precision != GPFY_PREC_TF32X3 && desc->precision != GPFY_PREC_F64_I8) { set_error("unknown precision %d", desc->precision); return GPFY_EINVAL; }
  if (s->nw > 1024) { set_error("projection with %d weights > 1024 is not built", s->nw); return GPFY_EUNSUPPORTED; }
  if (s->mmax > 1024) { set_error("m_l = %d > 1024 is not built", s->mmax); return GPFY_EUNSUPPORTED; }
  return GPFY_OK;
}

// ------------------------------------------------------------------------------------------------
// small parameter-algebra kernels (O(M^2) / O(M^3) per step, N-independent)
// ------------------------------------------------------------------------------------------------
struct PrepArgs {
  const double *w, *b, *v, *eig, *vhat;
  double *sw, *sb, *dvec, *vhat_p;
  int d, dim, dimp, M, Mp, scalar_w, proj;
  DegreeTable deg;
};

__global__ void prep_kernel(PrepArgs a) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  if (a.proj > 0) {
    for (int j = tid; j < a.d * a.proj; j += nth) a.sw[j] = a.w[j];  // [d, p] projection: identity bijector, no sqrt
  } else {
    for (int j = tid; j < a.d; j += nth) a.sw[j] = sqrt(a.w[a.scalar_w ? 0 : j]);
  }
  if (tid == 0) a.sb[0] = sqrt(a.b[0]);
  if (a.dvec) {
    for (int i = tid; i < a.Mp; i += nth) {
      double val = 0.0;
      if (i < a.M) {
        int l = 0, o = 0;
        while (i >= o + a.deg.m[l]) { o += a.deg.m[l]; ++l; }
        val = sqrt(a.eig[l] * a.v[0]);  // sqrt(1 / Kuu) * sqrt(variance): spherical_harmonics.py:362,385
      }
      a.dvec[i] = val;
    }
  }
  for (int e = tid; e < a.Mp * a.dimp; e += nth) {
    int i = e / a.dimp, c = e % a.dimp;
    a.vhat_p[e] = (i < a.M && c < a.dim) ? a.vhat[i * a.dim + c] : 0.0;
  }
}

// Linv (dense [Mp, Mp], block diagonal): column j of block l by forward substitution with L_l.
// One block per degree, one thread per column (strided).  Optionally scaled: out = rs[i] * Linv.
__global__ void linv_kernel(const double* lcat, double* Linv, int Mp, DegreeTable deg) {
  const int l = blockIdx.x;
  int off = 0;
  int64_t loff = 0;
  for (int k = 0; k < l; ++k) { off += deg.m[k]; loff += (int64_t)deg.m[k] * deg.m[k]; }
  const int m = deg.m[l];
  const double* L = lcat + loff;
  for (int j = threadIdx.x; j < m; j += blockDim.x) {
    for (int i = j; i < m; ++i) {
      double s = (i == j) ? 1.0 : 0.0;
      for (int k = j; k < i; ++k) s -= L[i * m + k] * Linv[(int64_t)(off + k) * Mp + off + j];
      Linv[(int64_t)(off + i) * Mp + off + j] = s / L[i * m + i];
    }
  }
}

// out[i][j] = rs ? rs[i] * in[i][j] : in[i][j];  outT[j][i] = same      (square Mp x Mp)
__global__ void scale_rows_transpose_kernel(const double* in, const double* rs, double* out, double* outT, int Mp) {
  __shared__ double tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = by + r, j = bx + threadIdx.x;
    double val = in[(int64_t)i * Mp + j];
    if (rs) val *= rs[i];
    if (out) out[(int64_t)i * Mp + j] = val;
    tile[r][threadIdx.x] = val;
  }
  __syncthreads();
  if (outT)
    for (int r = threadIdx.y; r < 32; r += blockDim.y) outT[(int64_t)(bx + r) * Mp + by + threadIdx.x] = tile[threadIdx.x][r];
}

// pad [M, M] -> [Mp, Mp] (zeros outside), optionally subtracting the identity on i < M
__global__ void pad_square_kernel(const double* in, double* out, int M, int Mp, int sub_identity) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)Mp * Mp) return;
  const int i = (int)(e / Mp), j = (int)(e % Mp);
  double val = (i < M && j < M) ? in[(int64_t)i * M + j] : 0.0;
  if (sub_identity && i == j && i < M) val -= 1.0;
  out[e] = val;
}

__global__ void sub_identity_kernel(double* A, int M, int Mp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M) A[(int64_t)i * Mp + i] -= 1.0;
}

// unpad [Mp, Mp] -> [M, M] with a scale
__global__ void unpad_square_kernel(const double* in, double* out, int M, int Mp, double scale) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)M * M) return;
  const int i = (int)(e / M), j = (int)(e % M);
  out[e] = scale * in[(int64_t)i * Mp + j];
}

// y[i] = sum_j op(A)[i][j] x[j * xs]   (A [Mp, Mp]; trans: use A^T), one warp per row
__global__ void gemv_kernel(const double* A, int Mp, int trans, const double* x, int xs, int xn, double* y, int ys, int yn) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= yn) return;
  double s = 0.0;
  for (int j = lane; j < xn; j += 32) s += (trans ? A[(int64_t)j * Mp + row] : A[(int64_t)row * Mp + j]) * x[(int64_t)j * xs];
  s = warp_sum(s);
  if (lane == 0) y[(int64_t)row * ys] = s;
}

// Gz [Mp, Mp] (full symmetric) = sum_ks Gpart blocks;  bz [Mp] = sum_ks bzpart
__global__ void reduce_g_kernel(const double* Gpart, const double* bzpart, double* Gz, double* bz, int Mp, SyrkSplit sp) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < (int64_t)Mp * Mp) {
    const int i = (int)(e / Mp), j = (int)(e % Mp);
    int bi = i / SYRK_B, bj = j / SYRK_B, ri = i % SYRK_B, rj = j % SYRK_B;
    // lower blocks only; inside a diagonal block only the lower half is guaranteed (upper warp tiles are skipped)
    if (bi < bj || (bi == bj && ri < rj)) { int t = bi; bi = bj; bj = t; t = ri; ri = rj; rj = t; }
    const int KS = sp.ks[sp.type(bi, bj)];
    const double* src = Gpart + (int64_t)sp.gslot(bi, bj) * (SYRK_B * SYRK_B) + ri * SYRK_B + rj;
    double s = 0.0;
    for (int ks = 0; ks < KS; ++ks) s += src[(int64_t)ks * (SYRK_B * SYRK_B)];
    Gz[e] = s;
  }
  if (e < Mp) {
    const int bi = (int)(e / SYRK_B);
    const int KS = sp.ks[sp.type(bi, bi)];
    const double* src = bzpart + (int64_t)sp.bslot(bi) * SYRK_B + (e % SYRK_B);
    double s = 0.0;
    for (int ks = 0; ks < KS; ++ks) s += src[(int64_t)ks * SYRK_B];
    bz[e] = s;
  }
}

// dVhat[i][c] = sum_ks dVpart[ks][i][c]  -> packed [M, dim].  Eight lanes share one element (strided over ks,
// then a fixed xor tree), so the KS-long dependent chain of loads is 8 times shorter.
__global__ void reduce_dv_kernel(const double* dVpart, double* out, int M, int Mp, int dim, int dimp, int KS) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int e = t >> 3, sub = t & 7;
  const bool ok = e < M * dim;
  const int i = ok ? e / dim : 0, c = ok ? e % dim : 0;
  const int ldv = (dimp + 7) / 8 * 8;
  double s = 0.0;
  if (ok)
    for (int ks = sub; ks < KS; ks += 8) s += dVpart[((int64_t)ks * Mp + i) * ldv + c];
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (ok && sub == 0) out[e] = s;
}

// dd[i] = sum_j (dE[i][j] + sum_p mu[i][p] bz[p][j]) * Linv[i][j]      (one warp per row, i < M)
__global__ void dd_kernel(const double* dE, const double* mu, const double* bz, const double* Linv, double* dd, int M, int Mp,
                          int P) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  double s = 0.0;
  for (int j = lane; j < M; j += 32) {
    double val = dE[(int64_t)row * Mp + j];
    for (int p = 0; p < P; ++p) val += mu[row * P + p] * bz[p * Mp + j];
    s += val * Linv[(int64_t)row * Mp + j];
  }
  s = warp_sum(s);
  if (lane == 0) dd[row] = s;
}

// per-degree:  B = tril(dvec_i (dE + mu bz^T))_ll ;  X = Linv_l^T B  (stored in tmp)
__global__ void dl_step1_kernel(const double* dE, const double* mu, const double* bz, const double* dvec, const double* Linv,
                                double* tmp, int Mp, int P, DegreeTable deg) {
  const int l = blockIdx.x;
  int off = 0;
  for (int k = 0; k < l; ++k) off += deg.m[k];
  const int m = deg.m[l];
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
    const int i = e / m, j = e % m;
    // X[i][j] = sum_{k >= max(i, j)} Linv[k][i] * B[k][j],  B[k][j] = dvec_k * dEfull[k][j] for j <= k
    double s = 0.0;
    for (int k = (i > j ? i : j); k < m; ++k) {
      const int gk = off + k, gj = off + j;
      double b = dE[(int64_t)gk * Mp + gj];
      for (int p = 0; p < P; ++p) b += mu[gk * P + p] * bz[p * Mp + gj];
      s += Linv[(int64_t)gk * Mp + off + i] * dvec[gk] * b;
    }
    tmp[(int64_t)(off + i) * Mp + off + j] = s;
  }
}
// dL_l = -(X Linv_l^T): dL[i][j] = -sum_{k <= j} X[i][k] Linv[j][k]
__global__ void dl_step2_kernel(const double* tmp, const double* Linv, double* dlcat, int Mp, DegreeTable deg) {
  const int l = blockIdx.x;
  int off = 0;
  int64_t loff = 0;
  for (int k = 0; k < l; ++k) { off += deg.m[k]; loff += (int64_t)deg.m[k] * deg.m[k]; }
  const int m = deg.m[l];
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
    const int i = e / m, j = e % m;
    double s = 0.0;
    for (int k = 0; k <= j; ++k) s += tmp[(int64_t)(off + i) * Mp + off + k] * Linv[(int64_t)(off + j) * Mp + off + k];
    dlcat[loff + e] = (j <= i) ? -s : 0.0;
  }
}

struct FinalArgs {
  const double* partials;
  int nblocks, npart;
  const double *dd, *dvec, *eig, *w, *b, *v;
  int M, d, scalar_w, proj, nw;
  DegreeTable deg;
  double* packed;
  GpfyElboLayout lay;
};
// column k of the per-block partials, one block per k: deterministic tree
__global__ void __launch_bounds__(256) partial_sums_kernel(const double* partials, int nblocks, int npart, double* out) {
  __shared__ double red[256];
  const int k = blockIdx.x;
  double s0 = 0.0, s1 = 0.0;
  int bI = threadIdx.x;
  for (; bI + 256 < nblocks; bI += 512) { s0 += partials[(int64_t)bI * npart + k]; s1 += partials[(int64_t)(bI + 256) * npart + k]; }
  if (bI < nblocks) s0 += partials[(int64_t)bI * npart + k];
  red[threadIdx.x] = s0 + s1;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[k] = red[0];
}

// the scalar chain rules on the summed partials (a.partials = [npart] sums); one warp, lanes stride the rows of a
// degree and combine with the fixed shuffle tree (deterministic)
__global__ void finalize_kernel(FinalArgs a) {
  const double* sums = a.partials;
  const int lane = threadIdx.x;
  if (lane == 0) {
    a.packed[a.lay.data_term] = sums[PART_E];
    a.packed[a.lay.d_sigma2] = sums[PART_DS2];
    a.packed[a.lay.d_b] = sums[PART_DB] / (2.0 * sqrt(a.b[0]));
  }
  if (a.proj > 0) {
    for (int j = lane; j < a.nw; j += 32) a.packed[a.lay.d_w + j] = sums[PART_DW + j];  // W has the identity bijector, no sqrt
  } else if (a.scalar_w) {
    double s = 0.0;
    for (int j = lane; j < a.d; j += 32) s += sums[PART_DW + j];
    s = warp_sum(s);
    if (lane == 0) a.packed[a.lay.d_w] = s / (2.0 * a.w[0]);
  } else {
    for (int j = lane; j < a.d; j += 32) a.packed[a.lay.d_w + j] = sums[PART_DW + j] / (2.0 * a.w[j]);
  }
  // eigenvalue / variance chain:  d_i = sqrt(eig_l v)
  double dv = sums[PART_DV];
  int base = 0;
  for (int l = 0; l < a.deg.F; ++l) {
    double s = 0.0;
    for (int j = lane; j < a.deg.m[l]; j += 32) s = fma(a.dd[base + j], a.dvec[base + j], s);
    s = warp_sum(s);
    if (lane == 0) a.packed[a.lay.d_eig + l] = s / (2.0 * a.eig[l]);
    dv += s / (2.0 * a.v[0]);
    base += a.deg.m[l];
  }
  if (lane == 0) a.packed[a.lay.d_v] = dv;
}

// KL(q || N(0, I)) for q = N(mu, L L^T) (variational.py:225-240, 262-286):
//   KL = -MP/2 - 1/2 sum log diag(L)^2 + 1/2 |mu|^2 + 1/2 |L|_F^2
// The gradients are elementwise (many blocks); the scalar is one block's fixed-order tree (deterministic).
__global__ void kl_tril_grad_kernel(const double* mu, const double* L, double* out, int M, int P) {
  const int64_t nmu = (int64_t)M * P, nL = (int64_t)P * M * M;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nmu) out[1 + e] = mu[e];
  if (e < nL) {
    const int i = (int)((e / M) % M), j = (int)(e % M);
    const double v = L[e];
    out[1 + nmu + e] = (i == j) ? v - 1.0 / v : v;
  }
}
__global__ void __launch_bounds__(1024) kl_tril_value_kernel(const double* mu, const double* L, double* out, int M, int P) {
  __shared__ double red[1024];
  const int64_t nmu = (int64_t)M * P, nL = (int64_t)P * M * M;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  for (int64_t e = threadIdx.x; e < nmu; e += blockDim.x) s0 = fma(0.5 * mu[e], mu[e], s0);
  int64_t e = threadIdx.x;
  for (; e + 3 * (int64_t)blockDim.x < nL; e += 4 * (int64_t)blockDim.x) {
    const double a0 = L[e], a1 = L[e + blockDim.x], a2 = L[e + 2 * (int64_t)blockDim.x], a3 = L[e + 3 * (int64_t)blockDim.x];
    s0 = fma(0.5 * a0, a0, s0); s1 = fma(0.5 * a1, a1, s1); s2 = fma(0.5 * a2, a2, s2); s3 = fma(0.5 * a3, a3, s3);
  }
  for (; e < nL; e += blockDim.x) s0 = fma(0.5 * L[e], L[e], s0);
  for (int64_t d = threadIdx.x; d < (int64_t)P * M; d += blockDim.x) {
    const double v = L[(d / M) * M * M + (d % M) * (int64_t)(M + 1)];
    s1 -= 0.5 * log(v * v);
  }
  red[threadIdx.x] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0] - 0.5 * (double)nmu;
}

// KL(q || N(0, I)) for q = N(mu, Sigma) with a dense covariance (variational.py:375-399).  One CTA per output p:
//   A <- Sigma_p, in-place right-looking Cholesky (lower), logdet = 2 sum log diag;  X = L^-1 by forward substitution, one
//   thread per column;  dKL/dSigma = 1/2 I - 1/2 X^T X (the symmetrised cotangent of jnp.linalg.cholesky's logdet).
__global__ void __launch_bounds__(1024) kl_full_chol_kernel(const double* Sigma, double* A, double* X, double* logdet, int M) {
  const int p = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int64_t MM = (int64_t)M * M;
  double* a = A + p * MM;
  double* x = X + p * MM;
  const double* s = Sigma + p * MM;
  for (int64_t e = tid; e < MM; e += nt) a[e] = s[e];
  __syncthreads();
  for (int j = 0; j < M; ++j) {
    const double d = sqrt(a[(int64_t)j * M + j]);       // every thread reads the same (updated) pivot
    __syncthreads();
    for (int i = j + tid; i < M; i += nt) a[(int64_t)i * M + j] = (i == j) ? d : a[(int64_t)i * M + j] / d;
    __syncthreads();
    const int rem = M - j - 1;                            // trailing update of the lower triangle: (i, k), j < k <= i
    for (int64_t e = tid; e < (int64_t)rem * rem; e += nt) {
      const int i = j + 1 + (int)(e / rem), k = j + 1 + (int)(e % rem);
      if (k <= i) a[(int64_t)i * M + k] -= a[(int64_t)i * M + j] * a[(int64_t)k * M + j];
    }
    __syncthreads();
  }
  // X = L^-1: column c by forward substitution (entries above the diagonal are zero)
  for (int c = tid; c < M; c += nt) {
    for (int i = 0; i < c; ++i) x[(int64_t)i * M + c] = 0.0;
    for (int i = c; i < M; ++i) {
      double acc = (i == c) ? 1.0 : 0.0;
      for (int k = c; k < i; ++k) acc -= a[(int64_t)i * M + k] * x[(int64_t)k * M + c];
      x[(int64_t)i * M + c] = acc / a[(int64_t)i * M + i];
    }
  }
  __shared__ double red[1024];
  double ld = 0.0;
  for (int i = tid; i < M; i += nt) { const double v = a[(int64_t)i * M + i]; ld += log(v * v); }
  red[tid] = ld;
  __syncthreads();
  for (int o = nt / 2; o > 0; o >>= 1) {
    if (tid < o) red[tid] += red[tid + o];
    __syncthreads();
  }
  if (tid == 0) logdet[p] = red[0];
}
// out = [KL, dKL/dmu, dKL/dSigma]: gradients elementwise; the scalar from one block's fixed-order tree
__global__ void kl_full_grad_kernel(const double* mu, const double* X, double* out, int M, int P) {
  const int64_t nmu = (int64_t)M * P, nS = (int64_t)P * M * M;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nmu) out[1 + e] = mu[e];
  if (e < nS) {
    const int p = (int)(e / ((int64_t)M * M)), i = (int)((e / M) % M), j = (int)(e % M);
    const double* x = X + (int64_t)p * M * M;
    double s = 0.0;
    for (int k = (i > j ? i : j); k < M; ++k) s = fma(x[(int64_t)k * M + i], x[(int64_t)k * M + j], s);
    out[1 + nmu + e] = (i == j ? 0.5 : 0.0) - 0.5 * s;
  }
}
__global__ void __launch_bounds__(1024) kl_full_value_kernel(const double* mu, const double* Sigma, const double* logdet, double* out, int M, int P) {
  __shared__ double red[1024];
  const int64_t nmu = (int64_t)M * P;
  double s = 0.0;
  for (int64_t e = threadIdx.x; e < nmu; e += blockDim.x) s = fma(0.5 * mu[e], mu[e], s);
  for (int64_t d = threadIdx.x; d < (int64_t)P * M; d += blockDim.x) s += 0.5 * Sigma[(d / M) * M * M + (d % M) * (int64_t)(M + 1)];
  for (int p = threadIdx.x; p < P; p += blockDim.x) s -= 0.5 * logdet[p];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0] - 0.5 * (double)nmu;
}

// ------------------------------------------------------------------------------------------------
// workspace plan
// ------------------------------------------------------------------------------------------------
struct Plan {
  Shape s;
  int sms;
  int64_t nc;       // points per chunk (multiple of 128)
  int nb;           // SYRK row blocks
  SyrkSplit split;  // per block type: how many CTAs share a chunk's points
  int KSV;          // dVhat splits
  int64_t nblocks_c; // lik_bwd blocks per full chunk
  int npart;
  // offsets (doubles)
  int64_t o_sw, o_sb, o_dvec, o_vhat_p, o_mu2, o_linv, o_E, o_ET, o_tmp1, o_tmp2, o_tmp3, o_LqP, o_W0, o_W2, o_dE, o_Gz, o_bz,
      o_dd, o_partials, o_Gpart, o_bzpart, o_dVpart, o_Z, o_C, o_XH, o_R, o_wq, o_wp, o_mz, o_psi, o_cq, o_dr, o_dT, o_W2f, o_Zf, o_Dp, o_xg, o_G32, o_bz32, total;
  int KSX;          // cross-Gram splits (features VJP)
  bool f32s;        // the weighted Gram matrix also runs on tcgen05 (syrk_tf32x3_kernel): no fp64 Z panel is written at all
  int H32, NC32, KS32;  // its column parts, columns per part, point splits
  bool f32;         // GPFY_PREC_TF32X3: float hi / lo planes of W2 (o_W2f) and Z (o_Zf) for the tcgen05 contraction
  bool i8;          // GPFY_PREC_F64_I8: C = W2 Z as exact int8 digit products on tcgen05 (gemm_i8.cuh); float64 results
  GemmI8Blocks i8blk;
  int i8nsa;
  int64_t o_Zi8, o_Wi8, o_rs8, o_ks8, o_bd8, o_ctr;
  int npsi;
};

static bool fused_bwd_ok(const Shape& s);
static bool bwd4_ok(const Shape& s);

static int64_t choose_chunk(int sms, int64_t n, const Shape& s) {
  // GEMM tiles are 96 x 192 with one persistent CTA per SM: chunks of 192 * sms * mult points give every CTA the same
  // number of tiles (no tail wave).  Larger chunks mean fewer launches and pipeline fills per step (measured at the
  // headline shape: mult 8 / 16 / 32 / 44 -> 85.1 / 86.0 / 86.4 / 86.5 M points/s), so the fused path defaults to
  // mult = 32, lowered (not below 8) where the per-chunk panels Z, C_p, dT would exceed 8 GB.  The general path keeps
  // mult = 8: its split-K dVhat kernel has a fixed number of warps per launch and loses more on long chunks than the
  // GEMM gains (13 degrees, N = 2 M: 46.9 ms at mult 8, 48.5 ms at 32).  Small inputs use one chunk.
  int mult = std::max(0, debug_option(DBG_CHUNK_MULT));  // 0 = the defaults below
  if (!mult) {
    const double bytes_per_point = (double)(s.P + 2) * s.Mp * 8.0;
    const int64_t cap = (int64_t)(8.0e9 / bytes_per_point) / ((int64_t)192 * sms);
    mult = (fused_bwd_ok(s) || bwd4_ok(s)) ? (int)std::min<int64_t>(32, std::max<int64_t>(8, cap)) : 8;
  }
  const int64_t unit = (int64_t)192 * sms, nc_max = unit * mult;
  const int64_t need = round_up64(n, 384);   // whole 128-point tcgen05 tiles and whole 192-column GEMM tiles
  if (need <= nc_max) return need;
  // equal chunks: k = ceil(n / nc_max) chunks of ceil(n / k) rows rounded up to whole tile rounds, so the last chunk is
  // not a sliver (n = 2 M: 3 x 0.68 M instead of 0.91 + 0.91 + 0.18 M)
  const int64_t k = (n + nc_max - 1) / nc_max;
  const int64_t per = (n + k - 1) / k;
  const int64_t nc = round_up64(per, unit);
  return nc < nc_max ? nc : nc_max;
}

static void make_plan(const Shape& s, int64_t n, int op_in, Plan* p, int prec = GPFY_PREC_F64) {
  const bool f32 = prec == GPFY_PREC_TF32X3;
  const bool fvjp = op_in == GPFY_OP_FEATURES_VJP;   // the features VJP shares the ELBO step's buffers plus two of its own
  const int op = fvjp ? GPFY_OP_ELBO : op_in;
  p->s = s;
  p->f32 = f32;
  p->sms = device_sms();
  p->nc = choose_chunk(p->sms, n, s);
  p->nb = (s.Mp + SYRK_B - 1) / SYRK_B;
  {
    SyrkSplit& sp = p->split;
    sp.nb = p->nb;
    sp.last_rows = s.Mp - (p->nb - 1) * SYRK_B;
    const int noff = p->nb * (p->nb - 1) / 2, slots = 2 * p->sms;
    // One uniform split count.  Measured (r02c): splits proportional to the tensor-core tile count of each block type
    // (12 diagonal / 18 off-diagonal / 12 partial) made the SYRK 5 - 10 % SLOWER at Mp = 352 and 256 - the kernel is not
    // bound by its DMMA count alone - so the per-type table stays at one value.
    // Round 2: the CTAs are persistent and pull (block, point range) units from a counter, heaviest block type first, so the table
    // only sets the granularity: about four units per resident CTA for a full chunk (a launch with fewer rows uses fewer, see
    // syrk_ks_eff).
    (void)noff;
    // (three block rows keep one unit per resident CTA, in the fixed order: no need for the finer slots)
    for (int t = 0; t < 4; ++t) sp.ks[t] = std::max(1, (p->nb == 3 ? 1 : 4) * slots / (p->nb * (p->nb + 1) / 2));
  }
  p->KSV = std::max(p->sms, (8 * p->sms) / (s.Mp / 32));  // dVhat partial slots: >= one per persistent CTA of the fused backward kernel
  p->nblocks_c = p->nc / 32;
  p->npsi = gemm_psi_slabs(s.Mp);
  p->npart = PART_DW + (s.proj > 0 ? s.nw : s.d);
  const int64_t Mp = s.Mp, Mp2 = Mp * Mp;
  int64_t o = 0;
  auto take = [&](int64_t cnt) { int64_t r = o; o += round_up64(cnt, 32); return r; };  // 256-byte aligned
  p->o_sw = take(std::max(s.d, s.nw));
  p->o_sb = take(1);
  p->o_dvec = take(Mp);
  p->o_vhat_p = take(Mp * s.dimp + (int64_t)s.P * Mp);  // mu2 sits right behind vhat_p (one TMA transaction)
  p->o_mu2 = p->o_vhat_p + Mp * s.dimp;
  p->o_linv = take(Mp2);
  p->o_E = take(Mp2);
  p->o_ET = take(Mp2);
  p->o_tmp1 = take(Mp2);
  p->o_tmp2 = take(Mp2);
  p->o_tmp3 = take(Mp2);
  p->o_LqP = take(s.P * Mp2);
  p->o_W0 = take(s.P * Mp2);
  p->o_W2 = take(s.P * Mp2);
  p->o_dE = take(Mp2);
  p->o_Gz = take(Mp2);
  p->o_bz = take(s.P * Mp);
  p->o_dd = take(Mp);
  p->o_partials = take(p->nblocks_c * p->npart);
  if (op == GPFY_OP_ELBO) {
    p->o_Gpart = take((int64_t)s.P * p->split.total() * SYRK_B * SYRK_B);
    p->o_bzpart = take((int64_t)s.P * p->split.bslot(p->nb) * SYRK_B);
    p->o_dVpart = take((int64_t)p->KSV * Mp * ((s.dimp + 7) / 8) * 8);
  } else {
    p->o_Gpart = p->o_bzpart = p->o_dVpart = o;
  }
  p->o_Z = take((Mp + 32) * p->nc);
  p->o_C = take((op == GPFY_OP_FEATURES ? 0 : (int64_t)s.P * Mp * p->nc));
  p->o_XH = take(p->nc * s.dimp);
  p->o_R = take(p->nc);
  p->o_wq = take((int64_t)s.P * p->nc);
  p->o_wp = take((int64_t)s.P * p->nc);
  p->o_mz = take((int64_t)s.P * p->nc);
  p->o_psi = take((int64_t)s.P * p->npsi * p->nc);
  p->o_cq = take((int64_t)s.P * p->nc);
  p->o_dr = take(p->nc);
  p->o_dT = take(op == GPFY_OP_ELBO ? Mp * p->nc : 0);
  // two float planes = one double per element
  p->o_W2f = take(f32 ? (int64_t)s.P * Mp2 : 0);
  p->o_Zf = take(f32 && op != GPFY_OP_FEATURES ? Mp * p->nc : 0);
  p->f32s = f32 && op == GPFY_OP_ELBO && s.P == 1 && s32_plan(s.Mp, &p->H32, &p->NC32) && !debug_option(DBG_NO_SYRK32);
  p->KS32 = p->f32s ? std::max(1, p->sms / p->H32) : 0;
  p->o_G32 = take(p->f32s ? (int64_t)p->KS32 * ((Mp + 127) / 128) * Mp * 128 : 0);
  p->o_bz32 = take(p->f32s ? (int64_t)p->KS32 * Mp : 0);
  p->KSX = std::max(1, (8 * p->sms) / std::max(1, s.Mp / 16));
  p->o_Dp = take(fvjp ? Mp * p->nc : 0);
  p->o_xg = take(fvjp ? (int64_t)p->KSX * Mp * XG_W : 0);
  // int8 digit planes (GPFY_PREC_F64_I8): built for one output and the blocked C layout of the fused backward kernels (or psi only);
  // every other shape keeps the DMMA contraction - the results are float64-accurate either way
  p->i8 = prec == GPFY_PREC_F64_I8 && !fvjp && op != GPFY_OP_FEATURES && s.P == 1 && (op == GPFY_OP_PREDICT || fused_bwd_ok(s) || bwd4_ok(s)) &&
          gi8_plan(s.Mp, p->sms, &p->i8blk, &p->i8nsa) && !debug_option(DBG_NO_I8);
  p->o_Zi8 = take(p->i8 ? (gi8_z_image_bytes(s.Mp, p->nc) + 7) / 8 : 0);
  p->o_Wi8 = take(p->i8 ? (gi8_w_image_bytes(s.Mp) + 7) / 8 : 0);
  p->o_rs8 = take(p->i8 ? Mp : 0);
  p->o_ks8 = take(p->i8 ? Mp : 0);
  p->o_bd8 = take(p->i8 ? Mp : 0);
  p->o_ctr = take(32);   // work-unit counter of the weighted-Gram kernel
  if (p->i8) p->npsi = std::max(p->npsi, p->i8blk.nblk);
  p->total = o;
}

static DegreeTable degree_table(const GpfyDesc* desc) {
  DegreeTable t;
  t.F = desc->num_degrees;
  for (int l = 0; l < GPFY_MAX_DEGREES; ++l) t.m[l] = l < t.F ? desc->m[l] : 0;
  return t;
}

template <typename K>
static int set_smem(K kernel, int bytes) {
  GPFY_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return GPFY_OK;
}

// dispatch on the padded ambient dimension
#define GPFY_DISPATCH_DIMP(dimp, CALL)          \
  switch (dimp) {                               \
    case 4: { constexpr int DIMP = 4; CALL; } break;   \
    case 8: { constexpr int DIMP = 8; CALL; } break;   \
    case 12: { constexpr int DIMP = 12; CALL; } break; \
    case 16: { constexpr int DIMP = 16; CALL; } break; \
    case 20: { constexpr int DIMP = 20; CALL; } break; \
    case 24: { constexpr int DIMP = 24; CALL; } break; \
    case 28: { constexpr int DIMP = 28; CALL; } break; \
    case 32: { constexpr int DIMP = 32; CALL; } break; \
    case 36: { constexpr int DIMP = 36; CALL; } break; \
    default: set_error("unsupported padded dim %d", dimp); return GPFY_EUNSUPPORTED; \
  }

template <int DIMP>
static int launch_zonal(cudaStream_t st, const ZonalArgs& a) {
  const unsigned grid = (unsigned)((((a.npts + 1) & ~(int64_t)1) + 31) / 32);
  if (a.P <= 4 && !debug_option(DBG_OLD_FWD)) {  // tensor-core forward kernel
    ZonalFwd2Args fa;
    fa.z = a;
    fa.mt = make_monic(0.5 * a.rc.two_alpha);
#ifdef GPFY_ZB_CLOCKS   // knock-out bits for time attribution: instrumented builds only (tools/zb_clocks.py)
    { const char* e = getenv("GPFY_ZF_DBG"); fa.dbg = e ? atoi(e) : 0; }
#else
    fa.dbg = 0;
#endif
    const int pmax = a.P == 1 ? 1 : (a.P == 2 ? 2 : 4);
    // barrier-free variant (one warp = one 32-point tile through all row tiles) when its per-warp tiles fit twice per SM
    const int smem3 = zf3_smem_bytes(a.Mp, DIMP, a.P, a.mz != nullptr, pmax);
    // (and when there are enough tiles to give every resident warp one: a small batch spreads better over zonal_fwd2_kernel's
    // seven warps per tile)
    if (smem3 <= 113 * 1024 && grid >= (unsigned)(2 * ZF3_WARPS * device_sms()) && !debug_option(DBG_NO_FWD3)) {
      const unsigned grid3 = (unsigned)std::min<int64_t>((grid + ZF3_WARPS - 1) / ZF3_WARPS, 2 * device_sms());
#define GPFY_ZF3(PM, ZP)                                                            \
  {                                                                                 \
    GPFY_TRY(set_smem((zonal_fwd3_kernel<DIMP, PM, ZP>), smem3));                   \
    zonal_fwd3_kernel<DIMP, PM, ZP><<<grid3, 32 * ZF3_WARPS, smem3, st>>>(fa);      \
  }
      if (a.ZPb) { if (pmax != 1) { set_error("the derivative panel is built for P = 1"); return GPFY_EINVAL; } GPFY_ZF3(1, true) }
      else if (pmax == 1) GPFY_ZF3(1, false) else if (pmax == 2) GPFY_ZF3(2, false) else GPFY_ZF3(4, false)
#undef GPFY_ZF3
      GPFY_COUNT_LAUNCH();
      GPFY_LAUNCH_OK();
      return GPFY_OK;
    }
    const int smem = zf2_smem_bytes(a.Mp, DIMP, a.P, a.mz != nullptr, pmax);
    const unsigned pgrid = std::min(grid, (unsigned)(2 * device_sms()));  // persistent: two CTAs per SM
#define GPFY_ZF(PM, ZP)                                                \
  {                                                                    \
    GPFY_TRY(set_smem(zonal_fwd2_kernel<DIMP, PM, ZP>, smem));         \
    zonal_fwd2_kernel<DIMP, PM, ZP><<<pgrid, 256, smem, st>>>(fa);     \
  }
    if (a.ZPb) { if (pmax != 1) { set_error("the derivative panel is built for P = 1"); return GPFY_EINVAL; } GPFY_ZF(1, true) }
    else if (pmax == 1) GPFY_ZF(1, false) else if (pmax == 2) GPFY_ZF(2, false) else GPFY_ZF(4, false)
#undef GPFY_ZF
    GPFY_COUNT_LAUNCH();
    GPFY_LAUNCH_OK();
    return GPFY_OK;
  }
  if (a.ZPb) { set_error("the legacy forward kernel has no derivative panel"); return GPFY_EINVAL; }
  const int smem = (a.Mp * DIMP + (a.mz ? a.P * a.Mp : 0)) * 8;
#define GPFY_ZN(PM)                                               \
  {                                                               \
    GPFY_TRY(set_smem(zonal_fwd_kernel<DIMP, PM>, smem));         \
    zonal_fwd_kernel<DIMP, PM><<<grid, 256, smem, st>>>(a);       \
  }
  if (a.P == 1) GPFY_ZN(1) else if (a.P == 2) GPFY_ZN(2) else if (a.P <= 4) GPFY_ZN(4) else GPFY_ZN(8)
#undef GPFY_ZN
  GPFY_COUNT_LAUNCH();
  GPFY_LAUNCH_OK();
  return GPFY_OK;
}

template <int DIMP>
static int launch_features2(cudaStream_t st, const FeaturesArgs& fa) {
  const int smem = ft_smem_bytes(fa.z.Mp, DIMP);
  const int tiles = (int)((fa.z.npts + 31) / 32);
  const unsigned grid = (unsigned)std::min(tiles, 2 * device_sms());
  GPFY_TRY(set_smem(features2_kernel<DIMP>, smem));
  features2_kernel<DIMP><<<grid, 256, smem, st>>>(fa);
  GPFY_COUNT_LAUNCH();
  GPFY_LAUNCH_OK();
  return GPFY_OK;
}

template <int DIMP>
static int launch_features3(cudaStream_t st, const Features3Args& fa) {
  if constexpr (DIMP <= 32) {
    const int64_t units = ((fa.f.z.npts + 31) / 32) * fa.ngrp;
    const unsigned grid = (unsigned)std::min<int64_t>((units + FT3_WARPS - 1) / FT3_WARPS, device_sms());
    const int smem_l = ft3_smem_bytes(fa.f.z.Mp, DIMP, fa.linv_doubles);
    if (smem_l <= 227 * 1024 - 64) {   // the triangles of Linv fit next to the private tiles
      GPFY_TRY(set_smem((features3_kernel<DIMP, true>), smem_l));
      features3_kernel<DIMP, true><<<grid, 32 * FT3_WARPS, smem_l, st>>>(fa);
    } else {
      const int smem = ft3_smem_bytes(fa.f.z.Mp, DIMP, 0);
      GPFY_TRY(set_smem((features3_kernel<DIMP, false>), smem));
      features3_kernel<DIMP, false><<<grid, 32 * FT3_WARPS, smem, st>>>(fa);
    }
    GPFY_COUNT_LAUNCH();
    GPFY_LAUNCH_OK();
    return GPFY_OK;
  } else {
    set_error("features3: padded dim %d > 32", DIMP);
    return GPFY_EUNSUPPORTED;
  }
}

template <int DIMP>
static int launch_zonal_bwd(cudaStream_t st, const ZonalBwdArgs& a) {
  const int smem = (a.Mp * DIMP + a.P * a.Mp + 8 * DIMP * 32) * 8;
  const unsigned grid = (unsigned)((a.npts + 31) / 32);
#define GPFY_ZB(PM)                                               \
  {                                                               \
    GPFY_TRY(set_smem(zonal_bwd_kernel<DIMP, PM>, smem));         \
    zonal_bwd_kernel<DIMP, PM><<<grid, 256, smem, st>>>(a);       \
  }
  if (a.P == 1) GPFY_ZB(1) else if (a.P == 2) GPFY_ZB(2) else if (a.P <= 4) GPFY_ZB(4) else GPFY_ZB(8)
#undef GPFY_ZB
  GPFY_COUNT_LAUNCH();
  GPFY_LAUNCH_OK();
  return GPFY_OK;
}

static int launch_dvhat(cudaStream_t st, const DvhatArgs& a) {
  const int warps = (a.Mp / 32) * a.KS;
  const unsigned grid = (warps + 3) / 4;
  switch ((a.dimp + 7) / 8) {
    case 1: dvhat_dmma_kernel<1><<<grid, 128, 0, st>>>(a); break;
    case 2: dvhat_dmma_kernel<2><<<grid, 128, 0, st>>>(a); break;
    case 3: dvhat_dmma_kernel<3><<<grid, 128, 0, st>>>(a); break;
    case 4: dvhat_dmma_kernel<4><<<grid, 128, 0, st>>>(a); break;
    default: dvhat_dmma_kernel<5><<<grid, 128, 0, st>>>(a); break;
  }
  GPFY_COUNT_LAUNCH();
  GPFY_LAUNCH_OK();
  return GPFY_OK;
}

static int launch_predict(cudaStream_t st, const PredictArgs& a) {
  predict_kernel<<<(unsigned)((a.npts + 255) / 256), 256, 0, st>>>(a);
  GPFY_COUNT_LAUNCH();
  GPFY_LAUNCH_OK();
  return GPFY_OK;
}

// The fused backward kernel covers P = 1 and shapes whose [Mp x 32] tile pair fits in shared memory.
static int fused_red_sep(const Shape& s) { return (s.Mp / 8) < 2 * ZB_WARPS ? 1 : 0; }
static bool fused_bwd_ok(const Shape& s) {
  return s.P == 1 && s.proj == 0 && s.dimp <= 16 && (s.Mp / 8) <= ZB_WARPS * ZB_SLOTS &&
         zb3_smem_bytes(s.Mp, s.dimp, fused_red_sep(s)) <= 227 * 1024 && !debug_option(DBG_NO_FUSED_BWD);
}

// The general fused backward kernel (zonal_bwd4): P = 1, no projection weights, up to 8 row tiles per compute warp.
static bool bwd4_ok(const Shape& s) {
  int RG, NST;
  return s.P == 1 && s.proj == 0 && (s.Mp / 8 + ZB4_CW - 1) / ZB4_CW <= (s.dimp > 16 ? 4 : 8) && zb4_plan(s.Mp, s.dimp, &RG, &NST) &&
         !debug_option(DBG_NO_BWD4) && !debug_option(DBG_OLD_FWD);
}

template <int DIMP>
static int launch_zonal_bwd4(cudaStream_t st, ZonalBwd4Args& a, int sms) {
  int RG = 0, NST = 0;
  if (!zb4_plan(a.Mp, DIMP, &RG, &NST)) { set_error("zonal_bwd4: no row split of Mp = %d fits in shared memory", a.Mp); return GPFY_EUNSUPPORTED; }
  a.RG = RG; a.G = a.Mp / RG; a.NST = NST; a.HB = zb4_header_ring(a.G, NST);
  const int smem = zb4_smem_bytes(a.Mp, DIMP, RG, NST);
  const int tiles = (int)((a.npts + 31) / 32);
  const unsigned grid = (unsigned)std::min(tiles, sms);
  const int slots = (a.Mp / 8 + ZB4_CW - 1) / ZB4_CW;
  if (slots <= 4) {
    GPFY_TRY(set_smem(zonal_bwd4_kernel<DIMP, 4>, smem));
    zonal_bwd4_kernel<DIMP, 4><<<grid, ZB4_THREADS, smem, st>>>(a);
  } else {
    if constexpr (DIMP <= 16) {
      GPFY_TRY(set_smem(zonal_bwd4_kernel<DIMP, 8>, smem));
      zonal_bwd4_kernel<DIMP, 8><<<grid, ZB4_THREADS, smem, st>>>(a);
    } else {
      set_error("zonal_bwd4: Mp = %d with ambient dim > 16 is not built", a.Mp);
      return GPFY_EUNSUPPORTED;
    }
  }
  GPFY_COUNT_LAUNCH();
  GPFY_LAUNCH_OK();
  return GPFY_OK;
}

template <int DIMP>
static int launch_zonal_bwd3(cudaStream_t st, const ZonalBwd3Args& a, bool fuse_lik, int sms) {
  if constexpr (DIMP <= 16) {
    const int smem = zb3_smem_bytes(a.Mp, DIMP, a.red_sep);
    const int tiles = (int)((a.npts + 31) / 32);
    const unsigned grid = (unsigned)std::min(tiles, sms);
    if (fuse_lik) {
      GPFY_TRY(set_smem(zonal_bwd3_kernel<DIMP, true>, smem));
      zonal_bwd3_kernel<DIMP, true><<<grid, ZB_THREADS, smem, st>>>(a);
    } else {
      GPFY_TRY(set_smem(zonal_bwd3_kernel<DIMP, false>, smem));
      zonal_bwd3_kernel<DIMP, false><<<grid, ZB_THREADS, smem, st>>>(a);
    }
    GPFY_COUNT_LAUNCH();
    GPFY_LAUNCH_OK();
    return GPFY_OK;
  } else {
    set_error("fused backward kernel is not built for padded dim %d", DIMP);
    return GPFY_EUNSUPPORTED;
  }
}

// C [Mp, Mp] = alpha A B + beta C for the O(M^3) parameter algebra of a step (7 products).  The big-tile GEMM
// would put these on 6 CTAs (54 us each); here every CTA owns a 32 x 32 tile (4 warps x 16 x 16, DMMA fragments
// straight from L2-resident operands), so an Mp = 288 product spreads over 81 CTAs.
__global__ void __launch_bounds__(256) gemm_small_kernel(const double* __restrict__ A, const double* __restrict__ B, double* C, int Mp,
                                                         double alpha, double beta) {
  __shared__ double part[4][32][8];  // second K half of each 16 x 16 warp tile, added in a fixed order
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const int wt = warp & 3, kh = warp >> 2;
  const int row0 = blockIdx.y * 32 + (wt >> 1) * 16, col0 = blockIdx.x * 32 + (wt & 1) * 16;
  const int kb = kh * (Mp / 2), ke = kb + Mp / 2;  // Mp is a multiple of 32
  double acc[2][2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const double* a0 = A + (int64_t)(row0 + g) * Mp + t4;
  const double* a1 = a0 + 8 * (int64_t)Mp;
  const double* b0 = B + (int64_t)t4 * Mp + col0 + g;
#pragma unroll 8
  for (int k = kb; k < ke; k += 4) {
    const double af0 = __ldg(a0 + k), af1 = __ldg(a1 + k);
    const double bf0 = __ldg(b0 + (int64_t)k * Mp), bf1 = __ldg(b0 + (int64_t)k * Mp + 8);
    dmma884(acc[0][0][0], acc[0][0][1], af0, bf0);
    dmma884(acc[0][1][0], acc[0][1][1], af0, bf1);
    dmma884(acc[1][0][0], acc[1][0][1], af1, bf0);
    dmma884(acc[1][1][0], acc[1][1][1], af1, bf1);
  }
  if (kh == 1) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) { part[wt][lane][4 * i + 2 * j] = acc[i][j][0]; part[wt][lane][4 * i + 2 * j + 1] = acc[i][j][1]; }
  }
  __syncthreads();
  if (kh == 0) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        double2* dst = reinterpret_cast<double2*>(C + (int64_t)(row0 + 8 * i + g) * Mp + col0 + 8 * j + 2 * t4);
        double2 v = make_double2(alpha * (acc[i][j][0] + part[wt][lane][4 * i + 2 * j]), alpha * (acc[i][j][1] + part[wt][lane][4 * i + 2 * j + 1]));
        if (beta != 0.0) { const double2 o = *dst; v.x += beta * o.x; v.y += beta * o.y; }
        *dst = v;
      }
  }
}

static int gemm_sq(cudaStream_t st, const double* A, const double* B, double* C, int Mp, double alpha, double beta, int sms) {
  (void)sms;
  dim3 grid(Mp / 32, Mp / 32);
  gemm_small_kernel<<<grid, 256, 0, st>>>(A, B, C, Mp, alpha, beta);
  GPFY_COUNT_LAUNCH();
  GPFY_LAUNCH_OK();
  return GPFY_OK;
}

#define GPFY_K(kernel, grid, block, smem, st, ...)    \
  do {                                                \
    kernel<<<grid, block, smem, st>>>(__VA_ARGS__);   \
    GPFY_COUNT_LAUNCH();                              \
    GPFY_LAUNCH_OK();                                 \
  } while (0)

// ---- the f32 instantiation of the contraction (GPFY_PREC_TF32X3): tcgen05 kernel of gemm_tf32.cuh ----
static bool want_f32(const GpfyDesc* desc) { return desc->precision == GPFY_PREC_TF32X3; }

typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {   // the driver API entry point through the runtime: no link-time libcuda dependency
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (TensorMapEncodeFn)f;
  });
  return fn;
}
// 2-D float tensor [outer][inner] (row stride = inner_ld floats), box = box_inner x box_outer elements
static int make_map_f32(CUtensorMap* m, const float* base, uint64_t inner, uint64_t outer, uint64_t inner_ld, uint32_t box_inner,
                        uint32_t box_outer, CUtensorMapSwizzle swz) {
  TensorMapEncodeFn enc = tensor_map_encoder();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return GPFY_EUNSUPPORTED; }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner_ld * sizeof(float)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return GPFY_ECUDA; }
  return GPFY_OK;
}

// The shapes the tcgen05 contraction is built for.
static int check_f32_shape(const Shape& s) {
  if (s.P != 1 || s.proj != 0 || s.Mp > 512 || g32_nsb(s.Mp) < 2) {
    set_error("GPFY_PREC_TF32X3 is built for num_outputs = 1, no projection weights and M <= 512 (got P = %d, M = %d)", s.P, s.M);
    return GPFY_EUNSUPPORTED;
  }
  return GPFY_OK;
}

// C (blocked, may be null) and / or psi = z . C for npts points of the chunk panels, output q
static int launch_gemm32(cudaStream_t st, const Plan& p, double* ws, int q, int64_t npts, double* C, double* psi) {
  const Shape& s = p.s;
  const int Mp = s.Mp;
  const int64_t Mp2 = (int64_t)Mp * Mp;
  float* zh = reinterpret_cast<float*>(ws + p.o_Zf);
  float* zl = zh + (int64_t)Mp * p.nc;
  float* wh = reinterpret_cast<float*>(ws + p.o_W2f) + 2 * q * Mp2;
  float* wl = wh + Mp2;
  Gemm32Maps maps;
  const uint64_t prow = (uint64_t)(p.nc / 32) * Mp;   // the tile-contiguous planes as [tiles * Mp rows][32 columns]
  GPFY_TRY(make_map_f32(&maps.zh, zh, 32, prow, 32, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  GPFY_TRY(make_map_f32(&maps.zl, zl, 32, prow, 32, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  GPFY_TRY(make_map_f32(&maps.wh, wh, (uint64_t)Mp, (uint64_t)Mp, (uint64_t)Mp, 32, (uint32_t)(Mp / 2), CU_TENSOR_MAP_SWIZZLE_128B));
  GPFY_TRY(make_map_f32(&maps.wl, wl, (uint64_t)Mp, (uint64_t)Mp, (uint64_t)Mp, 32, (uint32_t)(Mp / 2), CU_TENSOR_MAP_SWIZZLE_128B));
  Gemm32Args ga;
  ga.Zh = zh; ga.Zl = zl; ga.ldz = p.nc; ga.C = C; ga.psi = psi; ga.Mp = Mp; ga.npts = npts; ga.nsb = g32_nsb(Mp);
  const int smem = g32_smem_bytes(Mp);
  GPFY_TRY(set_smem(gemm_tf32x3_kernel, smem));
  const int tiles = (int)((npts + G32_BM - 1) / G32_BM);
  gemm_tf32x3_kernel<<<std::min(tiles, p.sms), G32_THREADS, smem, st>>>(maps, ga);
  GPFY_COUNT_LAUNCH();
  GPFY_LAUNCH_OK();
  return GPFY_OK;
}

// ---- float64-accurate contraction on the int8 tensor cores (GPFY_PREC_F64_I8): gemm_i8.cuh ----
// max over the sphere of |C_l^alpha(t)| = C_l^alpha(1) = prod_{i = 1..l} (2 alpha + i - 1) / i   (gegenbauer.py:27-60 at t = 1)
struct I8Bounds { double deg[GPFY_MAX_DEGREES]; };
__global__ void i8_bound_rows_kernel(I8Bounds bd, DegreeTable deg, int M, int Mp, double* bound) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= Mp) return;
  double v = 1.0;   // padding rows hold zeros: any scale
  if (k < M) {
    int o = 0, l = 0;
    while (k >= o + deg.m[l]) { o += deg.m[l]; ++l; }
    v = bd.deg[l];
  }
  bound[k] = v;
}
// digit planes + row scales of W2 and the per-feature scales of Z (once per step)
static int slice_w2_i8(cudaStream_t st, const GpfyDesc* desc, const Plan& p, double* ws) {
  const Shape& s = p.s;
  I8Bounds bd;
  for (int l = 0; l < GPFY_MAX_DEGREES; ++l) {
    double v = 1.0;
    for (int i = 1; i <= l; ++i) v *= (2.0 * desc->alpha + i - 1.0) / i;
    bd.deg[l] = v > 1.0 ? v : 1.0;
  }
  DegreeTable deg = degree_table(desc);
  GPFY_K(i8_bound_rows_kernel, (s.Mp + 127) / 128, 128, 0, st, bd, deg, s.M, s.Mp, ws + p.o_bd8);
  GPFY_K(w2_slice_i8_kernel, (s.Mp + 3) / 4, 128, 0, st, ws + p.o_W2, s.Mp, ws + p.o_bd8, p.i8blk, reinterpret_cast<int8_t*>(ws + p.o_Wi8),
         ws + p.o_rs8, ws + p.o_ks8);
  return GPFY_OK;
}
// Z panel -> digit planes, then C (blocked, may be null) and / or the psi slabs for npts points of the chunk
static int launch_gemm_i8(cudaStream_t st, const Plan& p, double* ws, int64_t npts, double* C, double* psi, bool sliced = false) {
  const Shape& s = p.s;
  const int Mp = s.Mp;
  int8_t* zs = reinterpret_cast<int8_t*>(ws + p.o_Zi8);
  const int tiles = (int)((npts + GI8_BM - 1) / GI8_BM);
  if (!sliced) {   // (on the fused Gaussian path the weighted-Gram kernel, launched first, has already written the digit planes)
    GPFY_TIMED(TK_SLICE, st);
    GPFY_K(z_slice_i8_kernel, dim3((unsigned)tiles, (unsigned)(Mp / 16)), GI8_BM, 0, st, ws + p.o_Z, p.nc, Mp, npts, ws + p.o_ks8, zs);
  }
  GemmI8Args ga;
  ga.Zs = zs; ga.Ws = reinterpret_cast<const int8_t*>(ws + p.o_Wi8); ga.rs = ws + p.o_rs8;
  ga.Z = psi ? ws + p.o_Z : nullptr; ga.ldz = p.nc; ga.C = C; ga.psi = psi; ga.ld_psi = p.nc;
  ga.Mp = Mp; ga.npts = npts; ga.nsa = p.i8nsa; ga.blk = p.i8blk;
  const int smem = gi8_smem_bytes(Mp, p.i8blk, p.i8nsa);
  GPFY_TRY(set_smem(gemm_i8_kernel, smem));
  GPFY_TIMED(TK_GEMM, st);
  gemm_i8_kernel<<<p.i8blk.cta0[p.i8blk.nblk], GI8_THREADS, smem, st>>>(ga);
  GPFY_COUNT_LAUNCH();
  GPFY_LAUNCH_OK();
  return GPFY_OK;
}

// G += Z diag(wq) Z^T, bz += Z wp over npts points of the chunk planes, on tcgen05 (P = 1)
static int launch_syrk32(cudaStream_t st, const Plan& p, double* ws, int64_t npts) {
  const Shape& s = p.s;
  const int Mp = s.Mp;
  float* zh = reinterpret_cast<float*>(ws + p.o_Zf);
  float* zl = zh + (int64_t)Mp * p.nc;
  Syrk32Args sa;
  sa.nbox = (Mp + 255) / 256;
  while (Mp % sa.nbox || (Mp / sa.nbox) % 8) ++sa.nbox;
  Syrk32Maps maps;
  const uint64_t prow = (uint64_t)(p.nc / 32) * Mp;
  const CUtensorMapSwizzle swz = S32_KP == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  GPFY_TRY(make_map_f32(&maps.zh, zh, 32, prow, 32, S32_KP, (uint32_t)(Mp / sa.nbox), swz));
  GPFY_TRY(make_map_f32(&maps.zl, zl, 32, prow, 32, S32_KP, (uint32_t)(Mp / sa.nbox), swz));
  sa.wq = ws + p.o_wq; sa.wp = ws + p.o_wp; sa.Mp = Mp; sa.NC = p.NC32; sa.H = p.H32; sa.ntile = (Mp + 127) / 128;
  sa.npts = npts; sa.range = round_up64((npts + p.KS32 - 1) / p.KS32, 32);
  sa.Gpart = ws + p.o_G32; sa.bzpart = ws + p.o_bz32;
  sa.kf = debug_option(DBG_SYRK32_KF) > 0 ? debug_option(DBG_SYRK32_KF) : S32_KF;
  const int smem = s32_smem_bytes(Mp, p.NC32);
  GPFY_TRY(set_smem(syrk_tf32x3_kernel, smem));
  syrk_tf32x3_kernel<<<p.H32 * p.KS32, S32_THREADS, smem, st>>>(maps, sa);
  GPFY_COUNT_LAUNCH();
  GPFY_LAUNCH_OK();
  return GPFY_OK;
}

// Shared prologue: sqrt weights, padded Vhat, Linv, E = diag(sqrt(eig v)) Linv, E^T, and per output
// W0 = L L^T - I (or Sigma - I), W2 = E^T W0 E, mu2 = E^T mu.
static int run_prologue(cudaStream_t st, const GpfyDesc* desc, const Plan& p, double* ws, const double* w, const double* b,
                        const double* v, const double* eig, const double* vhat, const double* lcat, const double* mu,
                        const double* sigma, bool need_model) {
  const Shape& s = p.s;
  const int Mp = s.Mp;
  const int64_t Mp2 = (int64_t)Mp * Mp;
  DegreeTable deg = degree_table(desc);
  PrepArgs pa;
  pa.w = w; pa.b = b; pa.v = v; pa.eig = eig; pa.vhat = vhat;
  pa.sw = ws + p.o_sw; pa.sb = ws + p.o_sb; pa.dvec = need_model ? ws + p.o_dvec : nullptr; pa.vhat_p = ws + p.o_vhat_p;
  pa.d = s.d; pa.dim = s.dim; pa.dimp = s.dimp; pa.M = s.M; pa.Mp = Mp; pa.scalar_w = desc->weight_mode == GPFY_W_SCALAR; pa.proj = s.proj;
  pa.deg = deg;
  GPFY_K(prep_kernel, 32, 256, 0, st, pa);
  GPFY_CUDA_OK(cudaMemsetAsync(ws + p.o_linv, 0, Mp2 * 8, st));
  GPFY_K(linv_kernel, s.F, 128, 0, st, lcat, ws + p.o_linv, Mp, deg);
  if (!need_model) return GPFY_OK;
  dim3 tg(Mp / 32, Mp / 32), tb(32, 8);
  GPFY_K(scale_rows_transpose_kernel, tg, tb, 0, st, ws + p.o_linv, ws + p.o_dvec, ws + p.o_E, ws + p.o_ET, Mp);
  const unsigned sqg = (unsigned)((Mp2 + 255) / 256);
  for (int q = 0; q < s.P; ++q) {
    double* W0 = ws + p.o_W0 + q * Mp2;
    if (desc->q_kind == GPFY_Q_TRIL) {
      double* LqP = ws + p.o_LqP + q * Mp2;
      GPFY_K(pad_square_kernel, sqg, 256, 0, st, sigma + (int64_t)q * s.M * s.M, LqP, s.M, Mp, 0);
      GPFY_K(scale_rows_transpose_kernel, tg, tb, 0, st, LqP, (const double*)nullptr, (double*)nullptr, ws + p.o_tmp1, Mp);
      GPFY_TRY(gemm_sq(st, LqP, ws + p.o_tmp1, W0, Mp, 1.0, 0.0, p.sms));  // L L^T   (variational.py:326-327)
      GPFY_K(sub_identity_kernel, (s.M + 127) / 128, 128, 0, st, W0, s.M, Mp);
    } else {
      GPFY_K(pad_square_kernel, sqg, 256, 0, st, sigma + (int64_t)q * s.M * s.M, W0, s.M, Mp, 1);  // Sigma - I (variational.py:438-440)
    }
    GPFY_TRY(gemm_sq(st, W0, ws + p.o_E, ws + p.o_tmp1, Mp, 1.0, 0.0, p.sms));              // W0 E
    GPFY_TRY(gemm_sq(st, ws + p.o_ET, ws + p.o_tmp1, ws + p.o_W2 + q * Mp2, Mp, 1.0, 0.0, p.sms));  // E^T W0 E
    if (p.f32) {   // hi / lo TF32 planes of W2 for the tcgen05 contraction
      float* wh = reinterpret_cast<float*>(ws + p.o_W2f) + 2 * q * Mp2;
      GPFY_K(tf32_split_kernel, sqg, 256, 0, st, ws + p.o_W2 + q * Mp2, wh, wh + Mp2, Mp2);
    }
    if (p.i8) GPFY_TRY(slice_w2_i8(st, desc, p, ws));   // digit planes of W2 for the int8 contraction (P = 1)
    // mu2_q = E^T mu[:, q]
    GPFY_K(gemv_kernel, (Mp + 3) / 4, 128, 0, st, ws + p.o_E, Mp, 1, mu + q, s.P, s.M, ws + p.o_mu2 + (int64_t)q * Mp, 1, Mp);
  }
  return GPFY_OK;
}

static int check_ws(const Plan& p, void* workspace, size_t bytes) {
  if (!workspace || ((uintptr_t)workspace & 255) != 0) { set_error("workspace must be 256-byte aligned and non-NULL"); return GPFY_EWORKSPACE; }
  if (bytes < (size_t)p.total * 8) { set_error("workspace too small: %zu < %zu bytes", bytes, (size_t)p.total * 8); return GPFY_EWORKSPACE; }
  return GPFY_OK;
}

}  // namespace gpfy

using namespace gpfy;

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int gpfy_abi_version(void) { return GPFY_ABI_VERSION; }
const char* gpfy_last_error_string(void) { return g_err; }
GPFY_API unsigned long long gpfy_debug_launch_count(void) { return g_launch_count.load(); }

// Development switches (tests / tools only; process-wide, default 0).  which: 0 = rows per chunk as a multiple of
// 192 * SMs (0 = default), 1 = legacy forward kernel, 2 = general (unfused) backward, 3 = two-kernel feature path.
GPFY_API int gpfy_debug_set_option(int which, int value) {
  if (which < 0 || which >= DBG_COUNT) { set_error("unknown debug option %d", which); return GPFY_EINVAL; }
  g_debug_opt[which].store(value, std::memory_order_relaxed);
  return GPFY_OK;
}

#ifdef GPFY_ZB_CLOCKS
GPFY_API int gpfy_debug_zb_clocks(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, g_zb_clk, sizeof(g_zb_clk)) == cudaSuccess ? 0 : -4;
}
#endif

// The int8 contraction's work split for Mp rows on `sms` SMs (host arithmetic only; tests/test_i8_plan_cpu.py):
// out[0] = built (0 / 1), out[1] = out-blocks, out[2] = ring stages, out[3] = shared-memory bytes, then per block (row0, rows, first CTA), then the CTA total.
GPFY_API int gpfy_debug_i8_plan(int Mp, int sms, int* out, int out_len) {
  GemmI8Blocks b;
  int nsa = 0;
  if (!out || out_len < 5 + 3 * GI8_MAXBLK) return GPFY_EINVAL;
  const bool ok = gi8_plan(Mp, sms, &b, &nsa);
  out[0] = ok ? 1 : 0;
  if (!ok) return GPFY_OK;
  out[1] = b.nblk; out[2] = nsa; out[3] = gi8_smem_bytes(Mp, b, nsa) + GI8_STATIC_SMEM;
  for (int i = 0; i < b.nblk; ++i) { out[4 + 3 * i] = b.row0[i]; out[5 + 3 * i] = b.rows[i]; out[6 + 3 * i] = b.cta0[i]; }
  out[4 + 3 * b.nblk] = b.cta0[b.nblk];
  return GPFY_OK;
}

// The barrier-free feature kernel's degree groups for degrees of m[0..F) rows (host arithmetic only; tests/test_features3_plan_cpu.py):
// out[0] = taken (0 / 1: every m_l <= 32), out[1] = groups, out[2] = doubles of the staged Linv triangles, out[3] / out[4] = shared-memory
// bytes with / without the staged triangles, then the first degree of every group and F.
GPFY_API int gpfy_debug_features3_plan(const int* m, int F, int Mp, int dimp, int* out, int out_len) {
  if (!m || !out || F < 1 || F > GPFY_MAX_DEGREES || out_len < 6 + GPFY_MAX_DEGREES + 1) return GPFY_EINVAL;
  DegreeTable deg;
  deg.F = F;
  for (int l = 0; l < F; ++l) deg.m[l] = m[l];
  Features3Args f3;
  const bool ok = ft3_groups(&f3, deg);
  out[0] = ok ? 1 : 0;
  if (!ok) return GPFY_OK;
  out[1] = f3.ngrp; out[2] = ft3_linv_doubles(deg);
  out[3] = ft3_smem_bytes(Mp, dimp, out[2]); out[4] = ft3_smem_bytes(Mp, dimp, 0);
  for (int g = 0; g <= f3.ngrp; ++g) out[5 + g] = f3.grp_lo[g];
  return GPFY_OK;
}

GPFY_API int gpfy_debug_kernel_timing(int enable) {
  g_timer.on = enable != 0;
  for (int k = 0; k < TK_COUNT; ++k) g_timer.used[k] = 0;
  return GPFY_OK;
}

// Sum of the recorded durations of kernel class `which` (TK_*), in milliseconds, and the number of
// launches.  Synchronises on the recorded events (a debug call, not part of the stream-ordered ABI).
GPFY_API int gpfy_debug_kernel_time(int which, double* ms, int64_t* launches) {
  if (which < 0 || which >= TK_COUNT || !ms || !launches) { set_error("bad argument"); return GPFY_EINVAL; }
  double total = 0.0;
  const size_t n = g_timer.used[which] / 2;
  for (size_t i = 0; i < n; ++i) {
    float t = 0.f;
    GPFY_CUDA_OK(cudaEventSynchronize(g_timer.ev[which][2 * i + 1]));
    GPFY_CUDA_OK(cudaEventElapsedTime(&t, g_timer.ev[which][2 * i], g_timer.ev[which][2 * i + 1]));
    total += t;
  }
  *ms = total;
  *launches = (int64_t)n;
  return GPFY_OK;
}

int64_t gpfy_num_inducing(const GpfyDesc* desc) {
  Shape s;
  int rc = make_shape(desc, &s);
  return rc == GPFY_OK ? s.M : rc;
}

int gpfy_elbo_packed_layout(const GpfyDesc* desc, GpfyElboLayout* out) {
  Shape s;
  GPFY_TRY(make_shape(desc, &s));
  if (!out) { set_error("out is NULL"); return GPFY_EINVAL; }
  int64_t o = 0;
  out->data_term = o; o += 1;
  out->d_mu = o; o += (int64_t)s.M * s.P;
  out->d_sigma = o; o += (int64_t)s.P * s.M * s.M;
  out->d_eig = o; o += s.F;
  out->d_vhat = o; o += (int64_t)s.M * s.dim;
  out->d_lcat = o; o += s.lsize;
  out->d_w = o; o += s.nw;
  out->d_b = o; o += 1;
  out->d_v = o; o += 1;
  out->d_sigma2 = o; o += 1;
  out->total = o;
  return GPFY_OK;
}

int gpfy_chunk_rows(const GpfyDesc* desc, int64_t n_points, int64_t* rows) {
  Shape s;
  GPFY_TRY(make_shape(desc, &s));
  if (!rows || n_points < 0) { set_error("bad argument"); return GPFY_EINVAL; }
  *rows = choose_chunk(device_sms(), n_points, s);
  return GPFY_OK;
}

int gpfy_contraction_path(const GpfyDesc* desc, int64_t n_points, int op, int32_t* path) {
  Shape s;
  GPFY_TRY(make_shape(desc, &s));
  if (!path || n_points < 0 || (op != GPFY_OP_ELBO && op != GPFY_OP_PREDICT)) { set_error("bad argument"); return GPFY_EINVAL; }
  Plan p;
  make_plan(s, n_points, op, &p, desc->precision);
  *path = p.f32 ? GPFY_PREC_TF32X3 : (p.i8 ? GPFY_PREC_F64_I8 : GPFY_PREC_F64);
  return GPFY_OK;
}

int gpfy_workspace_bytes(const GpfyDesc* desc, int64_t n_points, int op, size_t* bytes) {
  Shape s;
  GPFY_TRY(make_shape(desc, &s));
  if (!bytes || n_points < 0 || op < GPFY_OP_FEATURES || op > GPFY_OP_FEATURES_VJP) { set_error("bad argument"); return GPFY_EINVAL; }
  Plan p;
  if (op == GPFY_OP_FEATURES_VJP) s.P = 1;
  make_plan(s, n_points, op, &p, (op != GPFY_OP_FEATURES && op != GPFY_OP_FEATURES_VJP) ? desc->precision : GPFY_PREC_F64);
  *bytes = (size_t)p.total * 8;
  return GPFY_OK;
}

int gpfy_gegenbauer(void* stream, int level, double alpha, const double* x, int64_t n, double* out, double* dout) {
  if (level < 0 || level > GPFY_MAX_DEGREES) { set_error("level %d out of range", level); return GPFY_EINVAL; }
  if (n == 0) return GPFY_OK;
  RecCoef rc = make_rec_coef(alpha);
  GPFY_K(gegenbauer_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, rc, level, x, n, out, dout);
  return GPFY_OK;
}

int gpfy_sh_features_fwd(void* stream, const GpfyDesc* desc, int64_t n, const double* X, int apply_to_sphere, const double* w,
                         const double* b, const double* vhat, const double* lcat, const double* degree_scale, int mul_r,
                         double* out, double* r_out, void* workspace, size_t workspace_bytes) {
  cudaStream_t st = (cudaStream_t)stream;
  Shape s;
  GPFY_TRY(make_shape(desc, &s));
  if (n == 0) return GPFY_OK;
  Plan p;
  make_plan(s, n, GPFY_OP_FEATURES, &p);
  GPFY_TRY(check_ws(p, workspace, workspace_bytes));
  double* ws = (double*)workspace;
  const int Mp = s.Mp;
  DegreeTable deg = degree_table(desc);
  // parameter prologue; when X is already on the sphere w/b are unused: reuse vhat as a dummy pointer
  GPFY_TRY(run_prologue(st, desc, p, ws, apply_to_sphere ? w : vhat, apply_to_sphere ? b : vhat, nullptr, nullptr, vhat, lcat,
                        nullptr, nullptr, false));
  RecCoef rc = make_rec_coef(desc->alpha);
  Features3Args f3;
  // barrier-free fused kernel (one warp = one (tile, degree group) unit) when every degree fits its 32-row private tile
  const bool use_f3 = !debug_option(DBG_OLD_FEATURES) && !debug_option(DBG_NO_FEATURES3) && s.dimp <= 32 &&
                      ft3_smem_bytes(Mp, s.dimp, 0) <= 227 * 1024 - 64 && ft3_groups(&f3, deg);
  if (use_f3 || (ft_smem_bytes(Mp, s.dimp) <= 113 * 1024 && !debug_option(DBG_OLD_FEATURES))) {
    // one fused kernel: the zonal values stay in shared memory
    FeaturesArgs fa;
    ZonalArgs& za = fa.z;
    za.Yg = nullptr; za.ZPb = nullptr; za.Zh = za.Zl = nullptr; za.s2g = nullptr; za.wq = za.wp = za.cq = nullptr; za.proj = apply_to_sphere ? s.proj : 0;
    za.X = X; za.xcols = apply_to_sphere ? s.d : s.dim; za.apply_to_sphere = apply_to_sphere;
    za.sw = ws + p.o_sw; za.sb = ws + p.o_sb; za.vhat_p = ws + p.o_vhat_p;
    za.Mp = Mp; za.M = s.M; za.dim = s.dim; za.npts = n;
    za.Z = nullptr; za.ldz = 0; za.XH = nullptr; za.R = nullptr; za.mz = nullptr; za.P = 1;
    za.deg = deg; za.rc = rc;
    fa.mt = make_monic(desc->alpha);
    fa.Linv = ws + p.o_linv; fa.scale = degree_scale; fa.mul_r = mul_r; fa.out = out; fa.ldo = n; fa.r_out = r_out;
    if (use_f3) {
      f3.f = fa;
      f3.linv_doubles = ft3_linv_doubles(deg);
      GPFY_DISPATCH_DIMP(s.dimp, GPFY_TRY(launch_features3<DIMP>(st, f3)));
      return GPFY_OK;
    }
    ft_deal_items(&fa, deg);
    GPFY_DISPATCH_DIMP(s.dimp, GPFY_TRY(launch_features2<DIMP>(st, fa)));
    return GPFY_OK;
  }
  for (int64_t c0 = 0; c0 < n; c0 += p.nc) {
    const int64_t npts = std::min(p.nc, n - c0);
    ZonalArgs za;
    za.Yg = nullptr; za.ZPb = nullptr; za.Zh = za.Zl = nullptr; za.s2g = nullptr; za.wq = za.wp = za.cq = nullptr; za.proj = apply_to_sphere ? s.proj : 0;
    za.X = X + c0 * (apply_to_sphere ? s.d : s.dim);
    za.xcols = apply_to_sphere ? s.d : s.dim;
    za.apply_to_sphere = apply_to_sphere;
    za.sw = ws + p.o_sw; za.sb = ws + p.o_sb; za.vhat_p = ws + p.o_vhat_p;
    za.Mp = Mp; za.M = s.M; za.dim = s.dim; za.npts = npts;
    za.Z = ws + p.o_Z; za.ldz = p.nc; za.XH = nullptr; za.R = ws + p.o_R; za.mz = nullptr; za.P = 1;
    za.deg = deg; za.rc = rc;
    GPFY_DISPATCH_DIMP(s.dimp, GPFY_TRY(launch_zonal<DIMP>(st, za)));
    WhitenArgs wa;
    wa.Z = ws + p.o_Z; wa.ldz = p.nc; wa.Linv = ws + p.o_linv; wa.Mp = Mp;
    wa.scale = degree_scale; wa.R = mul_r ? ws + p.o_R : nullptr; wa.npts = npts;
    wa.out = out + c0; wa.ldo = n; wa.deg = deg;
    GPFY_K(whiten_kernel, (unsigned)((npts + 127) / 128), 128, 0, st, wa);
    if (r_out) GPFY_CUDA_OK(cudaMemcpyAsync(r_out + c0, ws + p.o_R, npts * 8, cudaMemcpyDeviceToDevice, st));
  }
  return GPFY_OK;
}

int gpfy_sh_features_vjp(void* stream, const GpfyDesc* desc, int64_t n, const double* X, int apply_to_sphere, const double* w,
                         const double* b, const double* vhat, const double* lcat, const double* degree_scale, int mul_r,
                         const double* dPhi, double* dvhat, double* dlcat, double* dscale, double* dw, double* db, void* workspace,
                         size_t workspace_bytes) {
  cudaStream_t st = (cudaStream_t)stream;
  Shape s;
  GPFY_TRY(make_shape(desc, &s));
  s.P = 1;   // the feature matrix does not depend on the number of outputs
  if (s.proj > 0 && apply_to_sphere) { set_error("gpfy_sh_features_vjp is not built for projection weights"); return GPFY_EUNSUPPORTED; }
  if (!bwd4_ok(s)) { set_error("gpfy_sh_features_vjp: shape outside the fused backward kernel (M = %d, ambient dim %d)", s.M, s.dim); return GPFY_EUNSUPPORTED; }
  const int Mp = s.Mp;
  const int64_t Mp2 = (int64_t)Mp * Mp;
  DegreeTable deg = degree_table(desc);
  // column window of every 16-row group of the block-diagonal cross Gram matrix
  XgCols xc;
  for (int rg = 0; rg < 72; ++rg) xc.col0[rg] = 0;
  for (int rg = 0; rg * 16 < Mp; ++rg) {
    const int first = rg * 16, last = std::min(rg * 16 + 15, s.M - 1);
    if (first >= s.M) continue;
    int lf = 0, ll = 0;
    while (first >= s.off[lf + 1]) ++lf;
    while (last >= s.off[ll + 1]) ++ll;
    xc.col0[rg] = s.off[lf];
    if (s.off[ll + 1] - s.off[lf] > XG_W) { set_error("gpfy_sh_features_vjp: degree blocks wider than the %d-column window", XG_W); return GPFY_EUNSUPPORTED; }
  }
  Plan p;
  make_plan(s, n, GPFY_OP_FEATURES_VJP, &p);
  GPFY_TRY(check_ws(p, workspace, workspace_bytes));
  double* ws = (double*)workspace;
  const int d_nw = desc->weight_mode == GPFY_W_SCALAR ? 1 : s.d;
  if (n == 0) {
    if (dvhat) GPFY_CUDA_OK(cudaMemsetAsync(dvhat, 0, (size_t)s.M * s.dim * 8, st));
    if (dlcat) GPFY_CUDA_OK(cudaMemsetAsync(dlcat, 0, (size_t)s.lsize * 8, st));
    if (dscale) GPFY_CUDA_OK(cudaMemsetAsync(dscale, 0, (size_t)s.F * 8, st));
    if (dw && apply_to_sphere) GPFY_CUDA_OK(cudaMemsetAsync(dw, 0, (size_t)d_nw * 8, st));
    if (db && apply_to_sphere) GPFY_CUDA_OK(cudaMemsetAsync(db, 0, 8, st));
    return GPFY_OK;
  }
  GPFY_TRY(run_prologue(st, desc, p, ws, apply_to_sphere ? w : vhat, apply_to_sphere ? b : vhat, nullptr, nullptr, vhat, lcat,
                        nullptr, nullptr, false));
  // A = diag(scale) Linv (E) and its transpose; mu2 = 0 (it sits behind vhat_p and is staged by the backward kernel)
  GPFY_K(row_scale_kernel, (Mp + 127) / 128, 128, 0, st, degree_scale, deg, s.M, Mp, ws + p.o_dvec);
  dim3 tg(Mp / 32, Mp / 32), tb(32, 8);
  GPFY_K(scale_rows_transpose_kernel, tg, tb, 0, st, ws + p.o_linv, ws + p.o_dvec, ws + p.o_E, ws + p.o_ET, Mp);
  GPFY_CUDA_OK(cudaMemsetAsync(ws + p.o_mu2, 0, (size_t)Mp * 8, st));
  GPFY_CUDA_OK(cudaMemsetAsync(ws + p.o_Z + (int64_t)s.M * p.nc, 0, (size_t)(Mp + 32 - s.M) * p.nc * 8, st));
  GPFY_CUDA_OK(cudaMemsetAsync(ws + p.o_Dp + (int64_t)s.M * p.nc, 0, (size_t)(Mp - s.M) * p.nc * 8, st));
  GPFY_CUDA_OK(cudaMemsetAsync(ws + p.o_partials, 0, (size_t)(p.o_Z - p.o_partials) * 8, st));
  GPFY_CUDA_OK(cudaMemsetAsync(ws + p.o_xg, 0, (size_t)p.KSX * Mp * XG_W * 8, st));
  RecCoef rc = make_rec_coef(desc->alpha);
  for (int64_t c0 = 0; c0 < n; c0 += p.nc) {
    const int64_t npts = std::min(p.nc, n - c0);
    const int64_t npts_e = (npts + 1) & ~(int64_t)1;
    ZonalArgs za;
    za.Yg = nullptr; za.ZPb = ws + p.o_dT; za.Zh = za.Zl = nullptr; za.s2g = nullptr; za.wq = za.wp = za.cq = nullptr;
    za.proj = apply_to_sphere ? s.proj : 0;
    za.X = X + c0 * (apply_to_sphere ? s.d : s.dim); za.xcols = apply_to_sphere ? s.d : s.dim; za.apply_to_sphere = apply_to_sphere;
    za.sw = ws + p.o_sw; za.sb = ws + p.o_sb; za.vhat_p = ws + p.o_vhat_p;
    za.Mp = Mp; za.M = s.M; za.dim = s.dim; za.npts = npts;
    za.Z = ws + p.o_Z; za.ldz = p.nc; za.XH = ws + p.o_XH; za.R = ws + p.o_R; za.mz = nullptr; za.P = 1;
    za.deg = deg; za.rc = rc;
    GPFY_DISPATCH_DIMP(s.dimp, GPFY_TRY(launch_zonal<DIMP>(st, za)));
    dim3 cg((unsigned)((npts_e + 255) / 256), s.M);
    GPFY_K(scale_panel_kernel, cg, 256, 0, st, dPhi + c0, n, s.M, npts, mul_r ? ws + p.o_R : (const double*)nullptr, ws + p.o_Dp, p.nc);
    {
      GemmArgs g;   // C = A^T Dp (cotangent of the zonal values), psi = z . C
      g.A = ws + p.o_ET; g.lda = Mp; g.B = ws + p.o_Dp; g.ldb = p.nc; g.C = ws + p.o_C; g.ldc = p.nc;
      g.m = Mp; g.k = Mp; g.n = (int)npts_e; g.alpha = 1.0; g.beta = 0.0;
      g.psi_out = ws + p.o_psi; g.ld_psi = p.nc; g.psi_B = ws + p.o_Z; g.c_blocked = 1;
      GPFY_TRY(launch_gemm(st, g, p.sms));
    }
    GPFY_K(dr_from_psi_kernel, (unsigned)((npts + 255) / 256), 256, 0, st, ws + p.o_psi, p.npsi, p.nc, ws + p.o_R, mul_r, npts, ws + p.o_dr);
    GPFY_K(fill_kernel, (unsigned)((npts_e + 255) / 256), 256, 0, st, ws + p.o_cq, npts_e, 1.0);
    GPFY_CUDA_OK(cudaMemsetAsync(ws + p.o_wp, 0, (size_t)npts_e * 8, st));
    {
      ZonalBwd4Args zb;
      zb.C = ws + p.o_C; zb.ZP = ws + p.o_dT; zb.wp = ws + p.o_wp; zb.cq = ws + p.o_cq; zb.dr = ws + p.o_dr;
      zb.XH = ws + p.o_XH; zb.R = ws + p.o_R; zb.vhat_p = ws + p.o_vhat_p; zb.Mp = Mp; zb.M = s.M; zb.d = s.d; zb.npts = npts;
      zb.partials = ws + p.o_partials; zb.npart = p.npart; zb.dVpart = ws + p.o_dVpart;
      GPFY_DISPATCH_DIMP(s.dimp, GPFY_TRY(launch_zonal_bwd4<DIMP>(st, zb, p.sms)));
    }
    {
      CrossGramArgs xa;
      xa.P = ws + p.o_Dp; xa.Q = ws + p.o_Z; xa.ld = p.nc; xa.Mp = Mp; xa.npts = npts;
      for (int i = 0; i < 72; ++i) xa.col0[i] = xc.col0[i];
      xa.part = ws + p.o_xg; xa.KS = p.KSX; xa.range = round_up64((npts + p.KSX - 1) / p.KSX, 4); xa.accumulate = 1;
      const int warps = (Mp / 16) * p.KSX;
      GPFY_K(cross_gram_kernel, (warps + 3) / 4, 128, 0, st, xa);
    }
  }
  const unsigned sqg = (unsigned)((Mp2 + 255) / 256);
  GPFY_K(cross_gram_reduce_kernel, sqg, 256, 0, st, ws + p.o_xg, xc, p.KSX, Mp, s.M, deg, ws + p.o_dE);
  GPFY_K(dd_kernel, (s.M + 3) / 4, 128, 0, st, ws + p.o_dE, (const double*)nullptr, (const double*)nullptr, ws + p.o_linv, ws + p.o_dd, s.M, Mp, 0);
  if (dlcat) {
    GPFY_K(dl_step1_kernel, s.F, 256, 0, st, ws + p.o_dE, (const double*)nullptr, (const double*)nullptr, ws + p.o_dvec, ws + p.o_linv, ws + p.o_tmp1, Mp, 0, deg);
    GPFY_K(dl_step2_kernel, s.F, 256, 0, st, ws + p.o_tmp1, ws + p.o_linv, dlcat, Mp, deg);
  }
  if (dvhat) GPFY_K(reduce_dv_kernel, (8 * s.M * s.dim + 255) / 256, 256, 0, st, ws + p.o_dVpart, dvhat, s.M, Mp, s.dim, s.dimp, p.KSV);
  GPFY_K(partial_sums_kernel, p.npart, 256, 0, st, ws + p.o_partials, (int)p.nblocks_c, p.npart, ws + p.o_tmp3);
  FeatVjpFinalArgs fa;
  fa.partials = ws + p.o_tmp3; fa.dd = ws + p.o_dd; fa.scale = degree_scale; fa.w = w; fa.b = b; fa.d = s.d;
  fa.scalar_w = desc->weight_mode == GPFY_W_SCALAR; fa.proj = s.proj; fa.nw = s.nw; fa.apply_to_sphere = apply_to_sphere; fa.deg = deg;
  fa.dw = dw; fa.db = db; fa.dscale = dscale;
  GPFY_K(features_vjp_finalize_kernel, 1, 32, 0, st, fa);
  return GPFY_OK;
}

// ---- the ELBO step in three stream-ordered stages sharing one workspace ----------------------------
// begin : parameter prologue (folded matrices) + zeroed accumulators
// rows  : any number of calls, each adds the terms of `n` rows of (X, Y) to the accumulators
// finish: reduces the accumulators and applies the chain rules back to the reference's parameters
static int elbo_plan(const GpfyDesc* desc, int64_t n_capacity, void* workspace, size_t workspace_bytes, Shape* s, Plan* p) {
  GPFY_TRY(make_shape(desc, s));
  if (n_capacity < 0) { set_error("n_capacity < 0"); return GPFY_EINVAL; }
  make_plan(*s, n_capacity, GPFY_OP_ELBO, p, desc->precision);
  if (p->f32) {
    GPFY_TRY(check_f32_shape(*s));
    if (!fused_bwd_ok(*s) && !bwd4_ok(*s)) { set_error("GPFY_PREC_TF32X3 needs one of the fused backward kernels for this shape"); return GPFY_EUNSUPPORTED; }
  }
  return check_ws(*p, workspace, workspace_bytes);
}

int gpfy_svgp_elbo_begin(void* stream, const GpfyDesc* desc, int64_t n_capacity, const double* w, const double* b,
                         const double* variance, const double* eig, const double* vhat, const double* lcat, const double* mu,
                         const double* sigma, void* workspace, size_t workspace_bytes) {
  cudaStream_t st = (cudaStream_t)stream;
  Shape s;
  Plan p;
  GPFY_TRY(elbo_plan(desc, n_capacity, workspace, workspace_bytes, &s, &p));
  double* ws = (double*)workspace;
  GPFY_TRY(run_prologue(st, desc, p, ws, w, b, variance, eig, vhat, lcat, mu, sigma, true));
  // pad rows of Z are never written by the zonal kernel; accumulators start at zero
  GPFY_CUDA_OK(cudaMemsetAsync(ws + p.o_Z + (int64_t)s.M * p.nc, 0, (size_t)(s.Mp + 32 - s.M) * p.nc * 8, st));
  GPFY_CUDA_OK(cudaMemsetAsync(ws + p.o_partials, 0, (size_t)(p.o_Z - p.o_partials) * 8, st));  // partials, Gpart, bzpart, dVpart
  if (p.f32s) GPFY_CUDA_OK(cudaMemsetAsync(ws + p.o_G32, 0, (size_t)(p.o_bz32 - p.o_G32 + (int64_t)p.KS32 * s.Mp) * 8, st));
  return GPFY_OK;
}

int gpfy_svgp_elbo_rows(void* stream, const GpfyDesc* desc, int64_t n_capacity, int64_t n, const double* X, const double* Y,
                        const double* variance, const double* sigma2, void* workspace, size_t workspace_bytes) {
  cudaStream_t st = (cudaStream_t)stream;
  Shape s;
  Plan p;
  GPFY_TRY(elbo_plan(desc, n_capacity, workspace, workspace_bytes, &s, &p));
  if (n < 0) { set_error("n < 0"); return GPFY_EINVAL; }
  double* ws = (double*)workspace;
  const int Mp = s.Mp;
  const int64_t Mp2 = (int64_t)Mp * Mp;
  DegreeTable deg = degree_table(desc);
  RecCoef rc = make_rec_coef(desc->alpha);

  const bool fused = fused_bwd_ok(s);                                  // zonal_bwd3: dT stays on chip, dVhat fused
  const bool fuse_lik = fused && desc->likelihood == GPFY_LIK_GAUSSIAN;  // + psi / likelihood fused (no GEMM psi epilogue)
  const bool bwd4 = !fused && bwd4_ok(s);                              // zonal_bwd4: C + derivative panel, any Mp / dim
  for (int64_t c0 = 0; c0 < n; c0 += p.nc) {
    const int64_t npts = std::min(p.nc, n - c0);
    const int64_t npts_e = (npts + 1) & ~(int64_t)1;
    ZonalArgs za;
    za.Yg = nullptr; za.ZPb = nullptr; za.Zh = za.Zl = nullptr; za.s2g = nullptr; za.wq = za.wp = za.cq = nullptr; za.proj = s.proj;
    za.X = X + c0 * s.d; za.xcols = s.d; za.apply_to_sphere = 1;
    za.sw = ws + p.o_sw; za.sb = ws + p.o_sb; za.vhat_p = ws + p.o_vhat_p;
    za.Mp = Mp; za.M = s.M; za.dim = s.dim; za.npts = npts;
    za.Z = ws + p.o_Z; za.ldz = p.nc; za.XH = ws + p.o_XH; za.R = ws + p.o_R; za.mz = ws + p.o_mz; za.P = s.P;
    za.deg = deg; za.rc = rc;
    // Gaussian weights need only fmu: the forward kernel writes them (likelihoods.py:210-218).  The fused path uses them throughout;
    // on the zonal_bwd4 path with the int8 contraction they let the weighted-Gram kernel (and its digit slicing) run before the
    // contraction - lik_point_kernel recomputes the same weights afterwards together with the value terms.
    const bool early_w = fuse_lik || (p.i8 && bwd4 && s.P == 1 && desc->likelihood == GPFY_LIK_GAUSSIAN);
    if (early_w) { za.Yg = Y + c0; za.s2g = sigma2; za.wq = ws + p.o_wq; za.wp = ws + p.o_wp; za.cq = ws + p.o_cq; }
    za.ZPb = bwd4 ? ws + p.o_dT : nullptr;   // the derivative panel takes the place of the legacy path's dT buffer
    if (p.f32) { za.Zh = reinterpret_cast<float*>(ws + p.o_Zf); za.Zl = za.Zh + (int64_t)Mp * p.nc; }
    if (p.f32s) za.Z = nullptr;   // both contractions read the float planes: the fp64 panel is not written
    {
      GPFY_TIMED(TK_ZONAL_FWD, st);
      GPFY_DISPATCH_DIMP(s.dimp, GPFY_TRY(launch_zonal<DIMP>(st, za)));
    }

    // G += Z diag(wq) Z^T, bz += Z wp (P outputs).  On the fused Gaussian path the weights are known after the forward kernel, so
    // the weighted Gram matrix can run BEFORE the contraction - and then its diagonal-block CTAs also cut Z into the int8 digit planes.
    auto launch_syrk = [&](bool slice) -> int {
      if (p.f32s) {
        GPFY_TIMED(TK_SYRK, st);
        return launch_syrk32(st, p, ws, npts);
      }
      for (int q = 0; q < s.P; ++q) {
        GPFY_TIMED(TK_SYRK, st);
        SyrkArgs sa;
        sa.Z = ws + p.o_Z; sa.ldz = p.nc; sa.Mp = Mp; sa.npts = npts;
        sa.wq = ws + p.o_wq + (int64_t)q * p.nc; sa.wp = ws + p.o_wp + (int64_t)q * p.nc;
        sa.Gpart = ws + p.o_Gpart + (int64_t)q * p.split.total() * SYRK_B * SYRK_B;
        sa.bzpart = ws + p.o_bzpart + (int64_t)q * p.split.bslot(p.nb) * SYRK_B;
        sa.split = p.split; sa.chunk = npts;   // the splits cover THIS launch's rows (a short slab still uses every CTA)
        sa.accumulate = 1;
        if (slice) { sa.zs_out = reinterpret_cast<int8_t*>(ws + p.o_Zi8); sa.kscale = ws + p.o_ks8; }
        // units of at least 1024 points, never fewer ranges than two resident CTAs per SM can take
        const int nblocks = p.nb * (p.nb + 1) / 2;
        const int ks_fix = std::max(1, 2 * p.sms / nblocks);   // one unit per resident CTA
        const bool fixed = p.nb == 3;   // three block rows (3 diagonal + 3 off-diagonal blocks): the fixed order pairs the two kinds on every SM; measured
                                         // against the dynamic order: Mp = 288 11.5 vs 12.6 ms, Mp = 256 11.6 vs 11.9 ms per 4 M points; Mp = 352 23.3 vs 21.9, Mp = 576 24.8 vs 22.3
        sa.ks_eff = fixed ? ks_fix : (int)std::min<int64_t>(p.split.ks[0], std::max<int64_t>(ks_fix, npts / 1024));
        sa.counter = fixed ? nullptr : reinterpret_cast<int*>(ws + p.o_ctr);
        if (sa.counter) GPFY_CUDA_OK(cudaMemsetAsync(sa.counter, 0, sizeof(int), st));
        GPFY_TRY(device_configure_once(DEV_CFG_SYRK, (const void*)syrk_dmma_kernel, SYRK_SMEM));
        GPFY_K(syrk_dmma_kernel, fixed ? nblocks * sa.ks_eff : std::min(nblocks * sa.ks_eff, 2 * p.sms), SYRK_THREADS, SYRK_SMEM, st, sa);
      }
      return GPFY_OK;
    };
    const bool syrk_first = p.i8 && early_w && !debug_option(DBG_NO_SYRK_SLICE);
    if (syrk_first) GPFY_TRY(launch_syrk(true));

    if (p.f32) {  // C = W2 Z on tcgen05 in 3 x TF32 (P = 1); psi goes out as one slab
      GPFY_TIMED(TK_GEMM, st);
      GPFY_TRY(launch_gemm32(st, p, ws, 0, npts, ws + p.o_C, fuse_lik ? nullptr : ws + p.o_psi));
    } else if (p.i8) {  // C = W2 Z as exact int8 digit products on tcgen05 (P = 1); psi in one slab per out-block
      GPFY_TRY(launch_gemm_i8(st, p, ws, npts, ws + p.o_C, nullptr, syrk_first));   // psi, where needed, is formed by lik_point_kernel
    } else
    for (int q = 0; q < s.P; ++q) {  // C_q = W2_q Z
      GPFY_TIMED(TK_GEMM, st);
      GemmArgs g;
      g.A = ws + p.o_W2 + q * Mp2; g.lda = Mp;
      g.B = ws + p.o_Z; g.ldb = p.nc;
      g.C = ws + p.o_C + (int64_t)q * Mp * p.nc; g.ldc = p.nc;
      g.m = Mp; g.k = Mp; g.n = (int)npts_e; g.alpha = 1.0; g.beta = 0.0;
      g.psi_out = fuse_lik ? nullptr : ws + p.o_psi + (int64_t)q * p.npsi * p.nc; g.ld_psi = p.nc;
      g.c_blocked = (fused || bwd4) ? 1 : 0;  // tile-contiguous, pre-swizzled C for the fused backward kernels
      g.m_eff = round_up(s.M, 8); g.k_eff = round_up(s.M, GEMM_BK);   // rows / columns M..Mp of W2 and rows M.. of Z are zero
      GPFY_TRY(launch_gemm(st, g, p.sms));
    }

    if (!fuse_lik) {
      LikPointArgs lp;
      lp.psi_part = ws + p.o_psi; lp.npsi = p.f32 ? 1 : (p.i8 ? p.i8blk.nblk : p.npsi); lp.mz = ws + p.o_mz; lp.ldz = p.nc;
      lp.R = ws + p.o_R; lp.Y = Y + c0 * s.P; lp.vptr = variance; lp.s2ptr = sigma2;
      lp.P = s.P; lp.order = desc->kernel_order; lp.lik_kind = desc->likelihood; lp.npts = npts;
      lp.wq = ws + p.o_wq; lp.wp = ws + p.o_wp; lp.cq = ws + p.o_cq; lp.dr = ws + p.o_dr;
      lp.partials = ws + p.o_partials; lp.npart = p.npart; lp.accumulate = 1;
      if (p.i8) { lp.Zp = ws + p.o_Z; lp.Cb = ws + p.o_C; lp.Mp = Mp; lp.M = s.M; }
      GPFY_TIMED(TK_LIK, st);
      GPFY_K(lik_point_kernel, (unsigned)((npts_e + 255) / 256), 256, 0, st, lp);
    }

    if (fused) {
      ZonalBwd3Args zb;
      zb.C = ws + p.o_C; zb.ldz = p.nc; zb.wp = ws + p.o_wp; zb.cq = ws + p.o_cq; zb.dr = ws + p.o_dr;
      zb.XH = ws + p.o_XH; zb.R = ws + p.o_R; zb.mz = ws + p.o_mz; zb.Y = Y + c0; zb.vptr = variance; zb.s2ptr = sigma2;
      zb.order = desc->kernel_order; zb.vhat_p = ws + p.o_vhat_p; zb.Mp = Mp; zb.M = s.M; zb.d = s.d; zb.npts = npts;
      zb.partials = ws + p.o_partials; zb.npart = p.npart; zb.dVpart = ws + p.o_dVpart;
      zb.deg = deg; zb.red_sep = fused_red_sep(s);
#ifdef GPFY_ZB_CLOCKS
      { const char* e = getenv("GPFY_ZB_DBG"); zb.dbg = e ? atoi(e) : 0; }
#else
      zb.dbg = 0;
#endif
      zb3_setup(&zb, desc->alpha, deg, s.M, Mp);
      GPFY_TIMED(TK_ZONAL_BWD, st);
      GPFY_DISPATCH_DIMP(s.dimp, GPFY_TRY(launch_zonal_bwd3<DIMP>(st, zb, fuse_lik, p.sms)));
    } else if (bwd4) {
      ZonalBwd4Args zb;
      zb.C = ws + p.o_C; zb.ZP = ws + p.o_dT; zb.wp = ws + p.o_wp; zb.cq = ws + p.o_cq; zb.dr = ws + p.o_dr;
      zb.XH = ws + p.o_XH; zb.R = ws + p.o_R; zb.vhat_p = ws + p.o_vhat_p; zb.Mp = Mp; zb.M = s.M; zb.d = s.d; zb.npts = npts;
      zb.partials = ws + p.o_partials; zb.npart = p.npart; zb.dVpart = ws + p.o_dVpart;
      GPFY_TIMED(TK_ZONAL_BWD, st);
      GPFY_DISPATCH_DIMP(s.dimp, GPFY_TRY(launch_zonal_bwd4<DIMP>(st, zb, p.sms)));
    } else {
      ZonalBwdArgs zb;
      zb.C = ws + p.o_C; zb.dT = ws + p.o_dT; zb.ldz = p.nc; zb.c_stride = (int64_t)Mp * p.nc;
      zb.wp = ws + p.o_wp; zb.cq = ws + p.o_cq; zb.dr = ws + p.o_dr;
      zb.XH = ws + p.o_XH; zb.R = ws + p.o_R; zb.vhat_p = ws + p.o_vhat_p;
      zb.Mp = Mp; zb.M = s.M; zb.d = s.sd; zb.P = s.P; zb.npts = npts;
      zb.X = X + c0 * s.d; zb.W = ws + p.o_sw; zb.xcols = s.d; zb.proj = s.proj;
      zb.partials = ws + p.o_partials; zb.npart = p.npart; zb.accumulate = 1;
      zb.deg = deg; zb.rc = rc;
      GPFY_TIMED(TK_ZONAL_BWD, st);
      GPFY_DISPATCH_DIMP(s.dimp, GPFY_TRY(launch_zonal_bwd<DIMP>(st, zb)));
    }

    if (!syrk_first) GPFY_TRY(launch_syrk(false));

    if (!fused && !bwd4) {
      DvhatArgs da;
      da.dT = ws + p.o_dT; da.ldz = p.nc; da.XH = ws + p.o_XH; da.dimp = s.dimp; da.Mp = Mp; da.npts = npts;
      da.dVpart = ws + p.o_dVpart; da.KS = p.KSV; da.range = round_up64((npts + p.KSV - 1) / p.KSV, 16);
      da.accumulate = 1;
      GPFY_TIMED(TK_DVHAT, st);
      GPFY_TRY(launch_dvhat(st, da));
    }
  }
  return GPFY_OK;
}

int gpfy_svgp_elbo_finish(void* stream, const GpfyDesc* desc, int64_t n_capacity, const double* w, const double* b,
                          const double* variance, const double* eig, const double* mu, double* packed, void* workspace,
                          size_t workspace_bytes) {
  cudaStream_t st = (cudaStream_t)stream;
  Shape s;
  Plan p;
  GPFY_TRY(elbo_plan(desc, n_capacity, workspace, workspace_bytes, &s, &p));
  GpfyElboLayout lay;
  GPFY_TRY(gpfy_elbo_packed_layout(desc, &lay));
  double* ws = (double*)workspace;
  const int Mp = s.Mp;
  const int64_t Mp2 = (int64_t)Mp * Mp;
  DegreeTable deg = degree_table(desc);
  // ---- gradients of the folded matrices back to the reference's parameters ----
  const unsigned sqg = (unsigned)((Mp2 + 255) / 256);
  for (int q = 0; q < s.P; ++q) {
    double* Gz = ws + p.o_Gz;
    double* bz = ws + p.o_bz + (int64_t)q * Mp;
    if (p.f32s) {
      GPFY_K(reduce_g32_kernel, sqg, 256, 0, st, ws + p.o_G32, ws + p.o_bz32, Gz, bz, Mp, p.KS32, (Mp + 127) / 128);
    } else {
      GPFY_K(reduce_g_kernel, sqg, 256, 0, st, ws + p.o_Gpart + (int64_t)q * p.split.total() * SYRK_B * SYRK_B,
             ws + p.o_bzpart + (int64_t)q * p.split.bslot(p.nb) * SYRK_B, Gz, bz, Mp, p.split);
    }
    GPFY_TRY(gemm_sq(st, ws + p.o_E, Gz, ws + p.o_tmp1, Mp, 1.0, 0.0, p.sms));             // T1 = E Gz
    GPFY_TRY(gemm_sq(st, ws + p.o_tmp1, ws + p.o_ET, ws + p.o_tmp2, Mp, 1.0, 0.0, p.sms));  // dW0 = T1 E^T
    double* dsig = packed + lay.d_sigma + (int64_t)q * s.M * s.M;
    const unsigned ug = (unsigned)(((int64_t)s.M * s.M + 255) / 256);
    if (desc->q_kind == GPFY_Q_TRIL) {
      GPFY_TRY(gemm_sq(st, ws + p.o_tmp2, ws + p.o_LqP + q * Mp2, ws + p.o_tmp3, Mp, 2.0, 0.0, p.sms));  // 2 dW0 L
      GPFY_K(unpad_square_kernel, ug, 256, 0, st, ws + p.o_tmp3, dsig, s.M, Mp, 1.0);
    } else {
      GPFY_K(unpad_square_kernel, ug, 256, 0, st, ws + p.o_tmp2, dsig, s.M, Mp, 1.0);
    }
    // d mu[:, q] = E bz_q
    GPFY_K(gemv_kernel, (s.M + 3) / 4, 128, 0, st, ws + p.o_E, Mp, 0, bz, 1, Mp, packed + lay.d_mu + q, s.P, s.M);
    // dE += 2 W0_q T1
    GPFY_TRY(gemm_sq(st, ws + p.o_W0 + q * Mp2, ws + p.o_tmp1, ws + p.o_dE, Mp, 2.0, q > 0 ? 1.0 : 0.0, p.sms));
  }
  GPFY_K(dd_kernel, (s.M + 3) / 4, 128, 0, st, ws + p.o_dE, mu, ws + p.o_bz, ws + p.o_linv, ws + p.o_dd, s.M, Mp, s.P);
  GPFY_K(dl_step1_kernel, s.F, 256, 0, st, ws + p.o_dE, mu, ws + p.o_bz, ws + p.o_dvec, ws + p.o_linv, ws + p.o_tmp1, Mp, s.P, deg);
  GPFY_K(dl_step2_kernel, s.F, 256, 0, st, ws + p.o_tmp1, ws + p.o_linv, packed + lay.d_lcat, Mp, deg);
  GPFY_K(reduce_dv_kernel, (8 * s.M * s.dim + 255) / 256, 256, 0, st, ws + p.o_dVpart, packed + lay.d_vhat, s.M, Mp, s.dim, s.dimp, p.KSV);
  FinalArgs fa;
  GPFY_K(partial_sums_kernel, p.npart, 256, 0, st, ws + p.o_partials, (int)p.nblocks_c, p.npart, ws + p.o_tmp3);
  fa.partials = ws + p.o_tmp3; fa.nblocks = (int)p.nblocks_c; fa.npart = p.npart;
  fa.dd = ws + p.o_dd; fa.dvec = ws + p.o_dvec; fa.eig = eig; fa.w = w; fa.b = b; fa.v = variance;
  fa.M = s.M; fa.d = s.d; fa.scalar_w = desc->weight_mode == GPFY_W_SCALAR; fa.proj = s.proj; fa.nw = s.nw; fa.deg = deg; fa.packed = packed; fa.lay = lay;
  GPFY_K(finalize_kernel, 1, 32, 0, st, fa);
  return GPFY_OK;
}

int gpfy_svgp_elbo_valgrad(void* stream, const GpfyDesc* desc, int64_t n, const double* X, const double* Y, const double* w,
                           const double* b, const double* variance, const double* eig, const double* vhat, const double* lcat,
                           const double* mu, const double* sigma, const double* sigma2, double* packed, void* workspace,
                           size_t workspace_bytes) {
  if (n == 0) {  // empty shard: all sums are zero
    GpfyElboLayout lay;
    GPFY_TRY(gpfy_elbo_packed_layout(desc, &lay));
    GPFY_CUDA_OK(cudaMemsetAsync(packed, 0, lay.total * 8, (cudaStream_t)stream));
    return GPFY_OK;
  }
  GPFY_TRY(gpfy_svgp_elbo_begin(stream, desc, n, w, b, variance, eig, vhat, lcat, mu, sigma, workspace, workspace_bytes));
  GPFY_TRY(gpfy_svgp_elbo_rows(stream, desc, n, n, X, Y, variance, sigma2, workspace, workspace_bytes));
  return gpfy_svgp_elbo_finish(stream, desc, n, w, b, variance, eig, mu, packed, workspace, workspace_bytes);
}

static int basis_launch(void* stream, const GpfyDesc* desc, BasisArgs& ba, void* workspace = nullptr, size_t workspace_bytes = 0) {
  Shape s;
  GPFY_TRY(make_shape(desc, &s));
  ba.scratch = nullptr;
  if (s.mmax > BASIS_MAX_M) {
    if (!workspace) { set_error("inducing-basis kernels keep m_l <= %d in shared memory (got %d): use the _ws entry points", BASIS_MAX_M, s.mmax); return GPFY_EUNSUPPORTED; }
    if (s.mmax > BASIS_MAX_M_WS) { set_error("inducing-basis kernels are built for m_l <= %d (got %d)", BASIS_MAX_M_WS, s.mmax); return GPFY_EUNSUPPORTED; }
    const size_t need = (size_t)4 * s.lsize * 8;
    if (((uintptr_t)workspace & 15) != 0 || workspace_bytes < need) { set_error("inducing basis with m_l > %d needs %zu bytes of 16-byte aligned workspace", BASIS_MAX_M, need); return GPFY_EWORKSPACE; }
    ba.scratch = (double*)workspace;
  }
  ba.dim = s.dim; ba.alpha = desc->alpha; ba.deg = degree_table(desc); ba.rc = make_rec_coef(desc->alpha);
  const int smem = basis_smem_bytes(s.mmax, s.dim, ba.scratch != nullptr);
  GPFY_TRY(set_smem(inducing_basis_kernel, smem));
  GPFY_K(inducing_basis_kernel, s.F, 256, smem, (cudaStream_t)stream, ba);
  return GPFY_OK;
}

int gpfy_inducing_basis_fwd(void* stream, const GpfyDesc* desc, const double* V, int normalise, double* vhat, double* lcat) {
  BasisArgs ba;
  ba.V = V; ba.vhat = vhat; ba.lcat = lcat; ba.dvhat = nullptr; ba.dlcat = nullptr; ba.dV = nullptr;
  ba.normalise = normalise; ba.vjp = 0;
  return basis_launch(stream, desc, ba);
}

int gpfy_inducing_basis_vjp(void* stream, const GpfyDesc* desc, const double* V, const double* lcat, int normalise,
                            const double* dvhat, const double* dlcat, double* dV) {
  BasisArgs ba;
  ba.V = V; ba.vhat = nullptr; ba.lcat = const_cast<double*>(lcat); ba.dvhat = dvhat; ba.dlcat = dlcat; ba.dV = dV;
  ba.normalise = normalise; ba.vjp = 1;
  return basis_launch(stream, desc, ba);
}

int gpfy_inducing_basis_fwd_ws(void* stream, const GpfyDesc* desc, const double* V, int normalise, double* vhat, double* lcat,
                               void* workspace, size_t workspace_bytes) {
  BasisArgs ba;
  ba.V = V; ba.vhat = vhat; ba.lcat = lcat; ba.dvhat = nullptr; ba.dlcat = nullptr; ba.dV = nullptr;
  ba.normalise = normalise; ba.vjp = 0;
  return basis_launch(stream, desc, ba, workspace, workspace_bytes);
}

int gpfy_inducing_basis_vjp_ws(void* stream, const GpfyDesc* desc, const double* V, const double* lcat, int normalise,
                               const double* dvhat, const double* dlcat, double* dV, void* workspace, size_t workspace_bytes) {
  BasisArgs ba;
  ba.V = V; ba.vhat = nullptr; ba.lcat = const_cast<double*>(lcat); ba.dvhat = dvhat; ba.dlcat = dlcat; ba.dV = dV;
  ba.normalise = normalise; ba.vjp = 1;
  return basis_launch(stream, desc, ba, workspace, workspace_bytes);
}

int gpfy_prior_kl_valgrad(void* stream, const GpfyDesc* desc, const double* mu, const double* sigma, double* out) {
  Shape s;
  GPFY_TRY(make_shape(desc, &s));
  if (desc->q_kind != GPFY_Q_TRIL) {
    set_error("prior KL for a full-covariance q needs a Cholesky of Sigma: left to the host framework");
    return GPFY_EUNSUPPORTED;
  }
  const int64_t nL = (int64_t)s.P * s.M * s.M;
  GPFY_K(kl_tril_grad_kernel, (unsigned)((nL + 255) / 256), 256, 0, (cudaStream_t)stream, mu, sigma, out, s.M, s.P);
  GPFY_K(kl_tril_value_kernel, 1, 1024, 0, (cudaStream_t)stream, mu, sigma, out, s.M, s.P);
  return GPFY_OK;
}

int gpfy_prior_kl_valgrad_ws(void* stream, const GpfyDesc* desc, const double* mu, const double* sigma, double* out, void* workspace,
                             size_t workspace_bytes) {
  Shape s;
  GPFY_TRY(make_shape(desc, &s));
  if (desc->q_kind == GPFY_Q_TRIL) return gpfy_prior_kl_valgrad(stream, desc, mu, sigma, out);
  const int64_t nS = (int64_t)s.P * s.M * s.M;
  const size_t need = (size_t)(2 * nS + s.P) * 8;
  if (!workspace || ((uintptr_t)workspace & 15) != 0 || workspace_bytes < need) {
    set_error("full-covariance KL needs %zu bytes of 16-byte aligned workspace", need);
    return GPFY_EWORKSPACE;
  }
  double* A = (double*)workspace;
  double* X = A + nS;
  double* logdet = X + nS;
  cudaStream_t st = (cudaStream_t)stream;
  GPFY_K(kl_full_chol_kernel, s.P, 1024, 0, st, sigma, A, X, logdet, s.M);
  GPFY_K(kl_full_grad_kernel, (unsigned)((nS + 255) / 256), 256, 0, st, mu, X, out, s.M, s.P);
  GPFY_K(kl_full_value_kernel, 1, 1024, 0, st, mu, sigma, logdet, out, s.M, s.P);
  return GPFY_OK;
}

int gpfy_predict_diag(void* stream, const GpfyDesc* desc, int64_t n, const double* X, const double* w, const double* b,
                      const double* variance, const double* eig, const double* vhat, const double* lcat, const double* mu,
                      const double* sigma, double* fmu, double* fvar, void* workspace, size_t workspace_bytes) {
  cudaStream_t st = (cudaStream_t)stream;
  Shape s;
  GPFY_TRY(make_shape(desc, &s));
  if (n == 0) return GPFY_OK;
  Plan p;
  make_plan(s, n, GPFY_OP_PREDICT, &p, desc->precision);
  if (p.f32) GPFY_TRY(check_f32_shape(s));
  GPFY_TRY(check_ws(p, workspace, workspace_bytes));
  double* ws = (double*)workspace;
  const int Mp = s.Mp;
  const int64_t Mp2 = (int64_t)Mp * Mp;
  DegreeTable deg = degree_table(desc);
  RecCoef rc = make_rec_coef(desc->alpha);
  GPFY_TRY(run_prologue(st, desc, p, ws, w, b, variance, eig, vhat, lcat, mu, sigma, true));

  // only the quadratic form z^T W2 z is needed: fold W2 onto its lower triangle (W_ij + W_ji below the diagonal, zero
  // above) and let the GEMM skip the k range right of each row tile's diagonal block
  for (int q = 0; q < s.P && !p.i8; ++q)   // (the int8 contraction forms the full product: its digit planes were cut from the unfolded W2)
    GPFY_K(fold_lower_kernel, dim3((Mp + 31) / 32, (Mp + 31) / 32), dim3(32, 8), 0, st, ws + p.o_W2 + q * Mp2, Mp);
  GPFY_CUDA_OK(cudaMemsetAsync(ws + p.o_Z + (int64_t)s.M * p.nc, 0, (size_t)(Mp + 32 - s.M) * p.nc * 8, st));
  for (int64_t c0 = 0; c0 < n; c0 += p.nc) {
    const int64_t npts = std::min(p.nc, n - c0);
    const int64_t npts_e = (npts + 1) & ~(int64_t)1;
    ZonalArgs za;
    za.Yg = nullptr; za.ZPb = nullptr; za.Zh = za.Zl = nullptr; za.s2g = nullptr; za.wq = za.wp = za.cq = nullptr; za.proj = s.proj;
    za.X = X + c0 * s.d; za.xcols = s.d; za.apply_to_sphere = 1;
    za.sw = ws + p.o_sw; za.sb = ws + p.o_sb; za.vhat_p = ws + p.o_vhat_p;
    za.Mp = Mp; za.M = s.M; za.dim = s.dim; za.npts = npts;
    za.Z = ws + p.o_Z; za.ldz = p.nc; za.XH = nullptr; za.R = ws + p.o_R; za.mz = ws + p.o_mz; za.P = s.P;
    za.deg = deg; za.rc = rc;
    if (p.f32) { za.Zh = reinterpret_cast<float*>(ws + p.o_Zf); za.Zl = za.Zh + (int64_t)Mp * p.nc; }
    GPFY_DISPATCH_DIMP(s.dimp, GPFY_TRY(launch_zonal<DIMP>(st, za)));
    if (p.f32) {
      GPFY_TRY(launch_gemm32(st, p, ws, 0, npts, nullptr, ws + p.o_psi));   // psi only, 3 x TF32 on tcgen05
    } else if (p.i8) {
      GPFY_TRY(launch_gemm_i8(st, p, ws, npts, ws + p.o_C, nullptr));  // C by int8 digit products on tcgen05; psi = z . C in predict_kernel
    } else
    for (int q = 0; q < s.P; ++q) {
      GemmArgs g;
      g.A = ws + p.o_W2 + q * Mp2; g.lda = Mp;
      g.B = ws + p.o_Z; g.ldb = p.nc;
      g.C = nullptr; g.ldc = p.nc;  // only psi = z . (W2 z) is needed (model.py:136): the product is never stored
      g.lower_k = 1;
      g.m_eff = round_up(s.M, 8); g.k_eff = round_up(s.M, GEMM_BK);
      g.m = Mp; g.k = Mp; g.n = (int)npts_e; g.alpha = 1.0; g.beta = 0.0;
      g.psi_out = ws + p.o_psi + (int64_t)q * p.npsi * p.nc; g.ld_psi = p.nc;
      GPFY_TRY(launch_gemm(st, g, p.sms));
    }
    PredictArgs pa;
    pa.psi_part = ws + p.o_psi; pa.npsi = p.f32 ? 1 : (p.i8 ? p.i8blk.nblk : p.npsi); pa.mz = ws + p.o_mz; pa.ldz = p.nc;
    pa.R = ws + p.o_R; pa.vptr = variance; pa.P = s.P; pa.order = desc->kernel_order; pa.npts = npts;
    pa.fmu = fmu + c0 * s.P; pa.fvar = fvar + c0 * s.P;
    if (p.i8) { pa.Zp = ws + p.o_Z; pa.Cb = ws + p.o_C; pa.Mp = Mp; pa.M = s.M; }
    GPFY_TRY(launch_predict(st, pa));
  }
  return GPFY_OK;
}

int gpfy_to_sphere(void* stream, const GpfyDesc* desc, int64_t n, const double* X, const double* w, const double* b, double* xs,
                   double* r) {
  cudaStream_t st = (cudaStream_t)stream;
  Shape s;
  GPFY_TRY(make_shape(desc, &s));
  if (n == 0) return GPFY_OK;
  const unsigned grid = (unsigned)((n + 127) / 128);
  const int sc = desc->weight_mode == GPFY_W_SCALAR;
  GPFY_DISPATCH_DIMP(s.dimp, GPFY_K(to_sphere_kernel<DIMP>, grid, 128, 0, st, X, s.d, s.dim, w, sc, b, n, xs, r, s.proj));
  return GPFY_OK;
}

int gpfy_variational_expectations(void* stream, const GpfyDesc* desc, int64_t n, const double* fmu, const double* fvar,
                                  const double* Y, const double* sigma2, double* out) {
  Shape s;
  GPFY_TRY(make_shape(desc, &s));
  if (n == 0) return GPFY_OK;
  GPFY_K(var_exp_kernel, (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, desc->likelihood, sigma2, n, s.P, fmu, fvar, Y,
         out);
  return GPFY_OK;
}

int gpfy_project_q(void* stream, const GpfyDesc* desc, int64_t n, const double* A, const double* mu, const double* sigma,
                   double* mean, double* var, void* workspace, size_t workspace_bytes) {
  cudaStream_t st = (cudaStream_t)stream;
  Shape s;
  GPFY_TRY(make_shape(desc, &s));
  if (n == 0) return GPFY_OK;
  Plan p;
  make_plan(s, n, GPFY_OP_PREDICT, &p);
  GPFY_TRY(check_ws(p, workspace, workspace_bytes));
  double* ws = (double*)workspace;
  const int Mp = s.Mp;
  const int64_t Mp2 = (int64_t)Mp * Mp;
  // S_q = L L^T (variational.py:326-327) or Sigma (:438-440), padded, into W2
  dim3 tg(Mp / 32, Mp / 32), tb(32, 8);
  const unsigned sqg = (unsigned)((Mp2 + 255) / 256);
  for (int q = 0; q < s.P; ++q) {
    double* S = ws + p.o_W2 + q * Mp2;
    if (desc->q_kind == GPFY_Q_TRIL) {
      double* LqP = ws + p.o_LqP + q * Mp2;
      GPFY_K(pad_square_kernel, sqg, 256, 0, st, sigma + (int64_t)q * s.M * s.M, LqP, s.M, Mp, 0);
      GPFY_K(scale_rows_transpose_kernel, tg, tb, 0, st, LqP, (const double*)nullptr, (double*)nullptr, ws + p.o_tmp1, Mp);
      GPFY_TRY(gemm_sq(st, LqP, ws + p.o_tmp1, S, Mp, 1.0, 0.0, p.sms));
    } else {
      GPFY_K(pad_square_kernel, sqg, 256, 0, st, sigma + (int64_t)q * s.M * s.M, S, s.M, Mp, 0);
    }
    // diag(A^T S A) only: fold S onto its lower triangle, the GEMM then skips a third of its k-steps (see gpfy_predict_diag)
    GPFY_K(fold_lower_kernel, dim3((Mp + 31) / 32, (Mp + 31) / 32), dim3(32, 8), 0, st, S, Mp);
  }
  GPFY_CUDA_OK(cudaMemsetAsync(ws + p.o_Z + (int64_t)s.M * p.nc, 0, (size_t)(Mp + 32 - s.M) * p.nc * 8, st));
  for (int64_t c0 = 0; c0 < n; c0 += p.nc) {
    const int64_t npts = std::min(p.nc, n - c0);
    const int64_t npts_e = (npts + 1) & ~(int64_t)1;
    dim3 cg((unsigned)((npts_e + 255) / 256), s.M);
    GPFY_K(copy_panel_kernel, cg, 256, 0, st, A + c0, n, s.M, npts, ws + p.o_Z, p.nc);
    for (int q = 0; q < s.P; ++q) {
      GemmArgs g;
      g.A = ws + p.o_W2 + q * Mp2; g.lda = Mp;
      g.B = ws + p.o_Z; g.ldb = p.nc;
      g.C = nullptr; g.ldc = p.nc;  // only psi = z . (W2 z) is needed (model.py:136): the product is never stored
      g.lower_k = 1;
      g.m_eff = round_up(s.M, 8); g.k_eff = round_up(s.M, GEMM_BK);
      g.m = Mp; g.k = Mp; g.n = (int)npts_e; g.alpha = 1.0; g.beta = 0.0;
      g.psi_out = ws + p.o_psi + (int64_t)q * p.npsi * p.nc; g.ld_psi = p.nc;
      GPFY_TRY(launch_gemm(st, g, p.sms));
    }
    GPFY_K(project_finish_kernel, (unsigned)((npts + 127) / 128), 128, 0, st, ws + p.o_Z, p.nc, s.M, s.P, mu, ws + p.o_psi, p.npsi,
           npts, mean + c0 * s.P, var + c0 * s.P);
  }
  return GPFY_OK;
}

// NCCL is resolved from the process at run time (the host framework's copy)
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, void*);
}  // extern "C"

#include <dlfcn.h>
extern "C" int gpfy_allreduce_packed(void* nccl_comm, void* stream, double* packed, int64_t count) {
  static nccl_allreduce_fn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    fn = (nccl_allreduce_fn)dlsym(RTLD_DEFAULT, "ncclAllReduce");
    if (!fn) {
      void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (h) fn = (nccl_allreduce_fn)dlsym(h, "ncclAllReduce");
    }
  });
  if (!fn) { set_error("ncclAllReduce not found in the process or libnccl.so.2"); return GPFY_EUNSUPPORTED; }
  // ncclDouble = 8, ncclSum = 0
  int rc = fn(packed, packed, (size_t)count, 8, 0, nccl_comm, stream);
  if (rc != 0) { set_error("ncclAllReduce returned %d", rc); return GPFY_ECUDA; }
  return GPFY_OK;
}
