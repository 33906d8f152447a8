// Device-side step driver: everything around the hot path that the reference does in
// host JAX every step - bijector forward (param.py:255), PolynomialDecay eigenvalues (spherical.py:563-587), the
// bijector / eigenvalue chain rules, the scale / N and KL bookkeeping of optimization.py:151-156 and the Adam
// update of training.py:78-86 - as a few tiny kernels on flat device vectors, so that a training step needs no
// host round trip (and can be captured in a CUDA graph).  O(#parameters) work; not on the N-dependent path.
#include "common.cuh"

namespace gpfy {

__global__ void step_constrain_kernel(GpfyStepLayout L, const double* __restrict__ theta, const int* __restrict__ tri_map,
                                      const double* __restrict__ cf, double* __restrict__ cons, int M, int P, int F, int fullcov) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  // weights: Exp (ARD / scalar) or Identity (projection)
  for (int64_t i = tid; i < L.n_w; i += nth) cons[L.c_w + i] = L.w_identity ? theta[L.th_w + i] : exp(theta[L.th_w + i]);
  if (tid == 0) {
    cons[L.c_b] = exp(theta[L.th_b]);
    cons[L.c_v] = exp(theta[L.th_v]);
    if (L.th_s2 >= 0) cons[L.c_s2] = exp(theta[L.th_s2]);
    if (L.kernel_kind == 1) {
      // lambda_n = (1 + n)^-beta cf_n / sum_k (1 + k)^-beta, levels clipped to truncation_level - 1 (spherical.py:579-587)
      const double beta = exp(theta[L.th_beta]);
      double S = 0.0, dS = 0.0;
      for (int k = 0; k < L.truncation_level; ++k) {
        const double dk = pow(1.0 + k, -beta);
        S += dk;
        dS += -log1p((double)k) * dk;
      }
      for (int l = 0; l < F; ++l) {
        const int n = l < L.truncation_level ? l : L.truncation_level - 1;
        const double dn = pow(1.0 + n, -beta), ddn = -log1p((double)n) * dn;
        cons[L.c_eig + l] = dn * cf[n] / S;
        cons[L.c_deig + l] = ddn * cf[n] / S - dn * cf[n] * dS / (S * S);
      }
      cons[L.c_beta] = beta;
    } else {
      for (int l = 0; l < F; ++l) { cons[L.c_eig + l] = cf[l]; cons[L.c_deig + l] = 0.0; }
    }
  }
  for (int64_t i = tid; i < (int64_t)M * P; i += nth) cons[L.c_mu + i] = theta[L.th_mu + i];
  const int64_t nV = L.th_total - L.th_V;
  for (int64_t i = tid; i < nV; i += nth) cons[L.c_V + i] = theta[L.th_V + i];
  if (fullcov) {
    for (int64_t i = tid; i < (int64_t)P * M * M; i += nth) cons[L.c_sig + i] = theta[L.th_sig + i];
  } else {
    // FillTriangular (variational.py:260): zero the dense matrices here, step_fill_tril_kernel scatters the vector
    for (int64_t i = tid; i < (int64_t)P * M * M; i += nth) cons[L.c_sig + i] = 0.0;
  }
}

__global__ void step_fill_tril_kernel(GpfyStepLayout L, const double* __restrict__ theta, const int* __restrict__ tri_map,
                                      double* __restrict__ cons, int M, int P) {
  const int64_t ntri = (int64_t)M * (M + 1) / 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * ntri) return;
  const int64_t p = i / ntri, k = i % ntri;
  cons[L.c_sig + p * M * M + tri_map[k]] = theta[L.th_sig + i];
}

// gradient of the loss (-ELBO) w.r.t. the unconstrained vector, and the loss itself
__global__ void step_grad_kernel(GpfyStepLayout L, GpfyElboLayout E, const double* __restrict__ cons, const double* __restrict__ packed,
                                 const double* __restrict__ kl, const double* __restrict__ dV, const int* __restrict__ tri_map,
                                 double c, double* __restrict__ grad, double* __restrict__ loss_out, int M, int P, int F, int fullcov) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  const int64_t nmu = (int64_t)M * P, nsig = (int64_t)P * M * M;
  for (int64_t i = tid; i < L.n_w; i += nth) {
    const double g = c * packed[E.d_w + i];
    grad[L.th_w + i] = -(L.w_identity ? g : g * cons[L.c_w + i]);   // Exp: d exp(x) / dx = exp(x)
  }
  if (tid == 0) {
    grad[L.th_b] = -c * packed[E.d_b] * cons[L.c_b];
    grad[L.th_v] = -c * packed[E.d_v] * cons[L.c_v];
    if (L.th_s2 >= 0) grad[L.th_s2] = -c * packed[E.d_sigma2] * cons[L.c_s2];
    if (L.kernel_kind == 1) {
      double s = 0.0;
      for (int l = 0; l < F; ++l) s += c * packed[E.d_eig + l] * cons[L.c_deig + l];
      grad[L.th_beta] = -s * cons[L.c_beta];
    }
    loss_out[0] = -(c * packed[E.data_term] - kl[0]);   // optimization.py:154-156
  }
  for (int64_t i = tid; i < nmu; i += nth) grad[L.th_mu + i] = -(c * packed[E.d_mu + i] - kl[1 + i]);
  if (fullcov) {
    for (int64_t i = tid; i < nsig; i += nth) grad[L.th_sig + i] = -(c * packed[E.d_sigma + i] - kl[1 + nmu + i]);
  } else {
    const int64_t ntri = (int64_t)M * (M + 1) / 2;
    for (int64_t i = tid; i < P * ntri; i += nth) {
      const int64_t p = i / ntri, k = i % ntri;
      const int64_t e = p * M * M + tri_map[k];
      grad[L.th_sig + i] = -(c * packed[E.d_sigma + e] - kl[1 + nmu + e]);   // FillTriangular's VJP: gather
    }
  }
  const int64_t nV = L.th_total - L.th_V;
  for (int64_t i = tid; i < nV; i += nth) grad[L.th_V + i] = -c * dV[i];
}

// optax.adam + masked set_to_zero (model.py:288-289, training.py:78-86)
__global__ void step_adam_kernel(int64_t n, double* __restrict__ theta, const double* __restrict__ grad, double* __restrict__ m1,
                                 double* __restrict__ m2, const unsigned char* __restrict__ mask, double lr, double b1, double b2,
                                 double eps, double bc1, double bc2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double g = grad[i];
  const double a = b1 * m1[i] + (1.0 - b1) * g;
  const double v = b2 * m2[i] + (1.0 - b2) * g * g;
  m1[i] = a;
  m2[i] = v;
  if (mask[i]) theta[i] += -lr * (a / bc1) / (sqrt(v / bc2) + eps);
}

// ------------------------------------------------------------------------------------------------
// Shuffled minibatches without a permutation array (replaces the HuggingFace
// `dataset.shuffle().iter(batch_size, drop_last_batch=True)` generator of optimization.py:31-58).
// Row j of an epoch is pi(j), pi a keyed bijection of [0, n): a balanced Feistel network on the smallest even number
// of bits covering n, cycle-walked back into range (at most 4 expected walks).  O(1) state per epoch (the round keys),
// no sort, no n-sized index buffer; the batch gather reads each selected row once and writes a dense [batch, row] slab.
// ------------------------------------------------------------------------------------------------
struct FeistelKeys { unsigned int k[GPFY_SHUFFLE_ROUNDS]; };

__host__ __device__ __forceinline__ unsigned int feistel_round(unsigned int r, unsigned int key) {
  unsigned int h = r * 0x9E3779B1u + key;
  h ^= h >> 16; h *= 0x85EBCA6Bu;
  h ^= h >> 13; h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}

__host__ __device__ __forceinline__ int64_t feistel_permute(int64_t x, int64_t n, int half_bits, const FeistelKeys& fk) {
  const unsigned long long mask = (1ull << half_bits) - 1ull;
  unsigned long long v = (unsigned long long)x;
  do {
    unsigned long long l = v >> half_bits, r = v & mask;
#pragma unroll
    for (int i = 0; i < GPFY_SHUFFLE_ROUNDS; ++i) {
      const unsigned long long t = l ^ ((unsigned long long)feistel_round((unsigned int)r, fk.k[i]) & mask);
      l = r;
      r = t;
    }
    v = (l << half_bits) | r;
  } while (v >= (unsigned long long)n);
  return (int64_t)v;
}

__host__ __device__ static inline int feistel_half_bits(int64_t n) {
  int bits = 2;
  while (bits < 62 && (1ll << bits) < n) ++bits;
  return (bits + 1) / 2;
}

__host__ __device__ static inline FeistelKeys feistel_keys(uint64_t seed, uint64_t epoch) {
  FeistelKeys fk;
  uint64_t z = seed * 0x9E3779B97F4A7C15ull + epoch * 0xD1B54A32D192ED03ull + 0x2545F4914F6CDD1Dull;
  for (int i = 0; i < GPFY_SHUFFLE_ROUNDS; ++i) {   // splitmix64 stream
    z += 0x9E3779B97F4A7C15ull;
    uint64_t y = z;
    y = (y ^ (y >> 30)) * 0xBF58476D1CE4E5B9ull;
    y = (y ^ (y >> 27)) * 0x94D049BB133111EBull;
    y ^= y >> 31;
    fk.k[i] = (unsigned int)(y >> 32);
  }
  return fk;
}

// A row is moved as 16-byte pieces of X (when dx is even and both X buffers are 16-byte aligned: VEC = 2) followed by its
// dy doubles of Y; consecutive threads take consecutive pieces of one row, so a row is one or two full 32-byte sectors
// per request and the dense output is written coalesced.  Every thread evaluates pi for its own row (about 60 integer
// operations, hidden behind the load).
template <int VEC>
__device__ __forceinline__ void gather_piece(const FeistelKeys& fk, int64_t n_rows, int half_bits, int64_t start, int64_t batch,
                                             const double* __restrict__ X, int dx, const double* __restrict__ Y, int dy,
                                             double* __restrict__ Xb, double* __restrict__ Yb, int64_t* __restrict__ idx_out, int64_t e) {
  const int px = dx / VEC;              // pieces of X per row
  const int row_len = px + dy;
  if (e >= batch * row_len) return;
  const int64_t j = e / row_len;
  const int c = (int)(e - j * row_len);
  const int64_t src = feistel_permute(start + j, n_rows, half_bits, fk);
  if (c < px) {
    if (VEC == 2) {
      *reinterpret_cast<double2*>(Xb + j * dx + 2 * c) = __ldg(reinterpret_cast<const double2*>(X + src * dx + 2 * c));
    } else {
      Xb[j * dx + c] = __ldg(X + src * dx + c);
    }
  } else {
    Yb[j * dy + (c - px)] = __ldg(Y + src * dy + (c - px));
  }
  if (c == 0 && idx_out) idx_out[j] = src;
}

template <int VEC>
__global__ void minibatch_gather_kernel(FeistelKeys fk, int64_t n_rows, int half_bits, int64_t start, int64_t batch,
                                        const double* __restrict__ X, int dx, const double* __restrict__ Y, int dy,
                                        double* __restrict__ Xb, double* __restrict__ Yb, int64_t* __restrict__ idx_out) {
  gather_piece<VEC>(fk, n_rows, half_bits, start, batch, X, dx, Y, dy, Xb, Yb, idx_out, (int64_t)blockIdx.x * blockDim.x + threadIdx.x);
}

// ------------------------------------------------------------------------------------------------
// Device-resident step state, so that a whole training step has no per-step host argument and can be replayed from
// a CUDA graph:  state[0] = steps taken, state[1] = epoch, state[2] = position inside the epoch.
// ------------------------------------------------------------------------------------------------
__global__ void step_adam_state_kernel(int64_t n, double* __restrict__ theta, const double* __restrict__ grad, double* __restrict__ m1,
                                       double* __restrict__ m2, const unsigned char* __restrict__ mask, double lr, double b1, double b2,
                                       double eps, const int64_t* __restrict__ state) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double t = (double)(state[0] + 1);
  const double bc1 = 1.0 - pow(b1, t), bc2 = 1.0 - pow(b2, t);
  const double g = grad[i];
  const double a = b1 * m1[i] + (1.0 - b1) * g;
  const double v = b2 * m2[i] + (1.0 - b2) * g * g;
  m1[i] = a;
  m2[i] = v;
  if (mask[i]) theta[i] += -lr * (a / bc1) / (sqrt(v / bc2) + eps);
}

template <int VEC>
__global__ void minibatch_gather_state_kernel(uint64_t seed, int64_t n_rows, int64_t batch, const int64_t* __restrict__ state,
                                              const double* __restrict__ X, int dx, const double* __restrict__ Y, int dy,
                                              double* __restrict__ Xb, double* __restrict__ Yb, int64_t* __restrict__ idx_out) {
  const FeistelKeys fk = feistel_keys(seed, (uint64_t)state[1]);
  gather_piece<VEC>(fk, n_rows, feistel_half_bits(n_rows), state[2], batch, X, dx, Y, dy, Xb, Yb, idx_out,
                    (int64_t)blockIdx.x * blockDim.x + threadIdx.x);
}

// end of a step: log the loss, count the step, move to the next batch (drop the last partial batch, then reshuffle)
__global__ void step_advance_kernel(int64_t* state, int64_t n_rows, int64_t batch, const double* __restrict__ loss,
                                    double* __restrict__ loss_trace, int64_t trace_capacity) {
  if (blockIdx.x || threadIdx.x) return;
  if (loss_trace && trace_capacity > 0) loss_trace[state[0] % trace_capacity] = loss[0];
  state[0] += 1;
  if (batch > 0) {
    int64_t pos = state[2] + batch;
    if (pos + batch > n_rows) { state[1] += 1; pos = 0; }
    state[2] = pos;
  }
}

}  // namespace gpfy

using namespace gpfy;

#define STEP_K(kernel, grid, block, st, ...)                 \
  do {                                                       \
    kernel<<<grid, block, 0, st>>>(__VA_ARGS__);             \
    GPFY_COUNT_LAUNCH();                                     \
    GPFY_LAUNCH_OK();                                        \
  } while (0)

// 16-byte pieces when every row of X and of the batch buffer starts 16-byte aligned
static inline int gather_vec(const double* X, const double* Xb, int dx) {
  return (dx % 2 == 0 && ((uintptr_t)X & 15) == 0 && ((uintptr_t)Xb & 15) == 0) ? 2 : 1;
}

extern "C" {

int gpfy_step_constrain(void* stream, const GpfyDesc* desc, const GpfyStepLayout* lay, const double* theta, const int* tri_map,
                        const double* cf, double* cons) {
  Shape s;
  GPFY_TRY(make_shape(desc, &s));
  if (!lay || !theta || !cons) { set_error("null argument"); return GPFY_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  const int fullcov = desc->q_kind == GPFY_Q_FULLCOV;
  STEP_K(step_constrain_kernel, 64, 256, st, *lay, theta, tri_map, cf, cons, s.M, s.P, s.F, fullcov);
  if (!fullcov) {
    const int64_t n = (int64_t)s.P * s.M * (s.M + 1) / 2;
    STEP_K(step_fill_tril_kernel, (unsigned)((n + 255) / 256), 256, st, *lay, theta, tri_map, cons, s.M, s.P);
  }
  return GPFY_OK;
}

int gpfy_step_grad(void* stream, const GpfyDesc* desc, const GpfyStepLayout* lay, const double* cons, const double* packed,
                   const double* kl, const double* dV, const int* tri_map, double scale_over_n, double* grad, double* loss_out) {
  Shape s;
  GPFY_TRY(make_shape(desc, &s));
  GpfyElboLayout E;
  GPFY_TRY(gpfy_elbo_packed_layout(desc, &E));
  STEP_K(step_grad_kernel, 128, 256, (cudaStream_t)stream, *lay, E, cons, packed, kl, dV, tri_map, scale_over_n, grad, loss_out, s.M,
         s.P, s.F, desc->q_kind == GPFY_Q_FULLCOV);
  return GPFY_OK;
}

int gpfy_step_adam(void* stream, int64_t n, double* theta, const double* grad, double* m1, double* m2, const unsigned char* mask,
                   double lr, double b1, double b2, double eps, int64_t step_count) {
  if (n <= 0) return GPFY_OK;
  const double bc1 = 1.0 - pow(b1, (double)step_count), bc2 = 1.0 - pow(b2, (double)step_count);
  STEP_K(step_adam_kernel, (unsigned)((n + 255) / 256), 256, (cudaStream_t)stream, n, theta, grad, m1, m2, mask, lr, b1, b2, eps, bc1, bc2);
  return GPFY_OK;
}

int gpfy_shuffle_index(uint64_t seed, uint64_t epoch, int64_t n_rows, int64_t position, int64_t* out) {
  if (!out || n_rows <= 0 || position < 0 || position >= n_rows) { set_error("shuffle index: position outside [0, n_rows)"); return GPFY_EINVAL; }
  *out = feistel_permute(position, n_rows, feistel_half_bits(n_rows), feistel_keys(seed, epoch));
  return GPFY_OK;
}

int gpfy_minibatch_gather(void* stream, uint64_t seed, uint64_t epoch, int64_t n_rows, int64_t start, int64_t batch, const double* X,
                          int dx, const double* Y, int dy, double* Xb, double* Yb, int64_t* idx_out) {
  if (n_rows <= 0 || batch < 0 || start < 0 || start + batch > n_rows || dx <= 0 || dy < 0) {
    set_error("minibatch gather: rows [%lld, %lld) outside the epoch of %lld rows", (long long)start, (long long)(start + batch), (long long)n_rows);
    return GPFY_EINVAL;
  }
  if (batch == 0) return GPFY_OK;
  if (gather_vec(X, Xb, dx) == 2) {
    const int64_t total = batch * (dx / 2 + dy);
    STEP_K(minibatch_gather_kernel<2>, (unsigned)((total + 255) / 256), 256, (cudaStream_t)stream, feistel_keys(seed, epoch), n_rows,
           feistel_half_bits(n_rows), start, batch, X, dx, Y, dy, Xb, Yb, idx_out);
  } else {
    const int64_t total = batch * (dx + dy);
    STEP_K(minibatch_gather_kernel<1>, (unsigned)((total + 255) / 256), 256, (cudaStream_t)stream, feistel_keys(seed, epoch), n_rows,
           feistel_half_bits(n_rows), start, batch, X, dx, Y, dy, Xb, Yb, idx_out);
  }
  return GPFY_OK;
}

int gpfy_step_adam_state(void* stream, int64_t n, double* theta, const double* grad, double* m1, double* m2, const unsigned char* mask,
                         double lr, double b1, double b2, double eps, const int64_t* state) {
  if (n <= 0) return GPFY_OK;
  STEP_K(step_adam_state_kernel, (unsigned)((n + 255) / 256), 256, (cudaStream_t)stream, n, theta, grad, m1, m2, mask, lr, b1, b2, eps, state);
  return GPFY_OK;
}

int gpfy_minibatch_gather_state(void* stream, uint64_t seed, int64_t n_rows, int64_t batch, const int64_t* state, const double* X, int dx,
                                const double* Y, int dy, double* Xb, double* Yb, int64_t* idx_out) {
  if (n_rows <= 0 || batch < 0 || batch > n_rows || dx <= 0 || dy < 0 || !state) { set_error("minibatch gather: bad argument"); return GPFY_EINVAL; }
  if (batch == 0) return GPFY_OK;
  if (gather_vec(X, Xb, dx) == 2) {
    const int64_t total = batch * (dx / 2 + dy);
    STEP_K(minibatch_gather_state_kernel<2>, (unsigned)((total + 255) / 256), 256, (cudaStream_t)stream, seed, n_rows, batch, state, X,
           dx, Y, dy, Xb, Yb, idx_out);
  } else {
    const int64_t total = batch * (dx + dy);
    STEP_K(minibatch_gather_state_kernel<1>, (unsigned)((total + 255) / 256), 256, (cudaStream_t)stream, seed, n_rows, batch, state, X,
           dx, Y, dy, Xb, Yb, idx_out);
  }
  return GPFY_OK;
}

int gpfy_step_advance(void* stream, int64_t* state, int64_t n_rows, int64_t batch, const double* loss, double* loss_trace,
                      int64_t trace_capacity) {
  if (!state || (batch > 0 && batch > n_rows)) { set_error("step advance: bad argument"); return GPFY_EINVAL; }
  STEP_K(step_advance_kernel, 1, 32, (cudaStream_t)stream, state, n_rows, batch, loss, loss_trace, trace_capacity);
  return GPFY_OK;
}

}  // extern "C"
