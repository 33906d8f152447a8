/* gpfy_b200 — C-ABI of the B200-native GPfY hot path (sm_100a).
 *
 * Drop-in boundary for ONE path of stefanosele/GPfY: spherical-harmonic features Phi(X) and the
 * sparse variational ELBO (+ gradient, + predict_diag) built on them.  The reference has no FFI of
 * its own (pure Python/JAX); these are the entry points its `jax.ffi` custom calls bind (see
 * INTEGRATION.md and gpfy_b200/csrc/xla_ffi_shim.cc).  Citations: /root/reference/src/gpfy/<file>:<line>.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless it says "host".
 *   - all arithmetic is IEEE float64 (the reference forces x64: utils.py:26, param.py:151) unless the descriptor asks for
 *     GPFY_PREC_TF32X3 (the f32 instantiation of the dense contraction); buffers are float64 either way.
 *   - stream-ordered: every call only enqueues work on `stream`; no allocation, no host sync.
 *     Scratch comes from the caller (`gpfy_workspace_bytes`).
 *   - return 0 on success, negative GPFY_E* otherwise; `gpfy_last_error_string()` explains
 *     (thread-local).  No exceptions cross the boundary.
 *   - layouts follow the reference: X [N, d] row-major; Phi / Kuf / A [M, N] with N contiguous, rows
 *     ordered by degree then by row of V_l (spherical_harmonics.py:320); mu [M, P]; Sigma [P, M, M]
 *     dense; fmu / fvar [N, P]; Y [N, P].
 *   - "Vhat" is the concatenation over degrees of the fundamental points actually used
 *     (`SphericalHarmonics.Vs`, spherical_harmonics.py:194-217) as [M, dim] row-major; "Lcat" is the
 *     concatenation of the per-degree lower Cholesky factors (`SphericalHarmonics.Ls`, :219-239),
 *     each [m_l, m_l] row-major.  Normalising V and the Cholesky of the Gram matrix (O(sum m_l^3),
 *     N-independent) stay on the host framework's side, like the bijectors.
 */
#ifndef GPFY_B200_H_
#define GPFY_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define GPFY_API __attribute__((visibility("default")))
#else
#define GPFY_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define GPFY_ABI_VERSION 2
#define GPFY_MAX_DEGREES 32

enum {
  GPFY_OK = 0,
  GPFY_EINVAL = -1,       /* bad descriptor / argument */
  GPFY_EUNSUPPORTED = -2, /* shape outside what the kernels are built for */
  GPFY_EWORKSPACE = -3,   /* workspace too small or misaligned */
  GPFY_ECUDA = -4         /* a CUDA runtime call failed (launch error) */
};

enum { GPFY_LIK_GAUSSIAN = 0, GPFY_LIK_BERNOULLI = 1 };
enum { GPFY_Q_TRIL = 0, GPFY_Q_FULLCOV = 1 };
enum { GPFY_W_ARD = 0, GPFY_W_SCALAR = 1, GPFY_W_PROJECTION = 2 };
/* Arithmetic of the dense contraction C = W2 Z (variational.py:323-328 folded; 55 % of the step's flops):
 *   GPFY_PREC_F64    : IEEE float64 on the FP64 tensor cores (DMMA) - the reference's precision (utils.py:26, param.py:151);
 *   GPFY_PREC_TF32X3 : the f32 instantiation - operands split into two TF32 planes, three tcgen05 kind::tf32 MMAs per
 *                      product into an fp32 TMEM accumulator (~2^-21 relative per product); everything else (zonal
 *                      recurrences, likelihood, reductions over N) stays float64.  Results agree with GPFY_PREC_F64 to 1e-5
 *                      (BASELINE.json north_star).  Built for num_outputs = 1, M <= 512, no projection weights.
 *   GPFY_PREC_F64_I8 : float64 RESULTS from the int8 tensor cores - an Ozaki-style error-free split of W2 and Z into six signed
 *                      radix-256 digits each, the 21 digit products with i + j <= 7 as tcgen05 kind::i8 MMAs into exact int32
 *                      TMEM accumulators, recombined in float64 (csrc/gemm_i8.cuh).  Normwise error of C ~1e-14 (DMMA: ~1e-15),
 *                      inside the 1e-10 contract of the float64 path; shapes it is not built for (num_outputs > 1, M > 864,
 *                      the one-point-per-lane backward kernels) run the DMMA contraction - same results, same contract. */
enum { GPFY_PREC_F64 = 0, GPFY_PREC_TF32X3 = 1, GPFY_PREC_F64_I8 = 2 };

/* POD, versioned.  Mirrors the static (non-pytree) fields the reference keeps on its dataclasses:
 * SphericalHarmonics.num_frequencies / phase_truncation (spherical_harmonics.py:75-76, via m[]),
 * Spherical.order / ard (spherical.py:77-80), constants["sphere"] (spherical_harmonics.py:149-153). */
typedef struct GpfyDesc {
  uint32_t struct_size; /* sizeof(GpfyDesc) */
  uint32_t abi_version; /* GPFY_ABI_VERSION */
  int32_t input_dim;    /* d = X.shape[-1]; ambient dim = sphere_dim + 1 (bias always appended, spherical.py:243-249),
                           sphere_dim = projection_dim if > 0 else d */
  int32_t num_degrees;  /* F; degrees are 0..F-1 (spherical_harmonics.py:78-80) */
  int32_t m[GPFY_MAX_DEGREES]; /* rows of V_l = min(phase_truncation, N(dim, l)); M = sum m[l] */
  int32_t num_outputs;  /* P = num_independent_processes (variational.py:86) */
  int32_t kernel_order; /* Spherical.order: K_diag = variance * r^(2*order) (spherical.py:400) */
  int32_t likelihood;   /* GPFY_LIK_* (likelihoods.py:188-220 / :254-278) */
  int32_t q_kind;       /* GPFY_Q_TRIL: Sigma holds L, q = N(mu, L L^T) (variational.py:243-351);
                           GPFY_Q_FULLCOV: Sigma is the covariance (variational.py:354-464) */
  int32_t weight_mode;  /* GPFY_W_ARD: weight_variances [d]; GPFY_W_SCALAR: one value (spherical.py:110-114);
                           GPFY_W_PROJECTION: a [d, p] linear projection X @ W, identity bijector (spherical.py:110-117,225-226) */
  int32_t projection_dim; /* p for GPFY_W_PROJECTION, else 0 */
  int32_t precision;    /* GPFY_PREC_*: arithmetic of the dense contraction (gpfy_svgp_elbo_*, gpfy_predict_diag) */
  double alpha;         /* (ambient dim - 2) / 2 (spherical_harmonics.py:149) */
} GpfyDesc;

/* Which scratch size to ask for. */
enum { GPFY_OP_FEATURES = 0, GPFY_OP_ELBO = 1, GPFY_OP_PREDICT = 2, GPFY_OP_FEATURES_VJP = 3 };

/* Layout of the packed ELBO output (doubles); offsets are filled by gpfy_elbo_packed_layout.
 * Everything is a SUM over the N points passed in (no 1/N, no dataset_size scale, no KL), so ranks
 * that own disjoint rows of (X, Y) can add their buffers with one allreduce. */
typedef struct GpfyElboLayout {
  int64_t total;     /* number of doubles */
  int64_t data_term; /* [1]  sum_n var_exp_n                         (optimization.py:153,156) */
  int64_t d_mu;      /* [M, P]                                        (variational.py:305) */
  int64_t d_sigma;   /* [P, M, M] dense, w.r.t. the constrained Sigma (variational.py:326 / :438) */
  int64_t d_eig;     /* [F]   w.r.t. kernel.eigenvalues lambda_l      (spherical_harmonics.py:335,385) */
  int64_t d_vhat;    /* [M, dim]                                      (spherical_harmonics.py:308) */
  int64_t d_lcat;    /* [sum m_l^2], lower triangles valid            (spherical_harmonics.py:316) */
  int64_t d_w;       /* [d], [1] or [d, p]                            (spherical.py:225-228) */
  int64_t d_b;       /* [1] bias_variance                             (spherical.py:245-246) */
  int64_t d_v;       /* [1] kernel variance                           (spherical_harmonics.py:362, spherical.py:400) */
  int64_t d_sigma2;  /* [1] Gaussian noise variance; 0 for Bernoulli  (likelihoods.py:210-218) */
} GpfyElboLayout;

GPFY_API int gpfy_abi_version(void);
GPFY_API const char* gpfy_last_error_string(void);

/* M and sum m_l^2 of a descriptor; negative error code if the descriptor is invalid. */
GPFY_API int64_t gpfy_num_inducing(const GpfyDesc* desc);
GPFY_API int gpfy_elbo_packed_layout(const GpfyDesc* desc, GpfyElboLayout* out /* host */);

/* Bytes of scratch `op` needs for up to n_points rows (host call, no GPU work). */
GPFY_API int gpfy_workspace_bytes(const GpfyDesc* desc, int64_t n_points, int op, size_t* bytes /* host */);

/* Rows the kernels process per internal chunk for an input of n_points rows (a whole number of GEMM tile rounds per SM);
 * callers that stream slabs through gpfy_svgp_elbo_rows size their slabs as multiples of it. */
GPFY_API int gpfy_chunk_rows(const GpfyDesc* desc, int64_t n_points, int64_t* rows /* host */);

/* Which arithmetic the dense contraction of `op` (GPFY_OP_ELBO / GPFY_OP_PREDICT) runs in for this descriptor on the current
 * device: GPFY_PREC_F64_I8 is a request - shapes the int8 kernel is not built for run the DMMA contraction (GPFY_PREC_F64), with
 * the same float64 contract.  Host call, no GPU work. */
GPFY_API int gpfy_contraction_path(const GpfyDesc* desc, int64_t n_points, int op, int32_t* path /* host: GPFY_PREC_* */);

/* Elementwise Gegenbauer polynomial and its derivative 2*alpha*C_{level-1}^{alpha+1}
 * (replaces gegenbauer.py:27-85: `gegenbauer`, `gegenbauer_fwd`).  `dout` may be NULL. */
GPFY_API int gpfy_gegenbauer(void* stream, int level, double alpha, const double* x, int64_t n, double* out, double* dout);

/* Spherical-harmonic feature matrix (replaces spherical_harmonics.py:290-320 `polynomial_expansion`,
 * :344-362 `Kuf`, :385-387 `project_fun`):
 *     out[i, n] = degree_scale[l(i)] * (r_n if mul_r) * (L_l^{-1} C_l^alpha(V_l xhat_n))[i]
 * apply_to_sphere != 0: X is [N, d] raw input, xhat/r come from to_sphere (spherical.py:230-249) with w, b.
 * apply_to_sphere == 0: X is [N, dim] already on the sphere (what `polynomial_expansion` receives);
 *                       w, b are ignored and r = 1.
 * degree_scale: [F] device array or NULL (= 1).  Kuf: sqrt(variance) each, mul_r = 1;
 * projection A: sqrt(lambda_l * variance), mul_r = 1.  r_out [N] may be NULL. */
GPFY_API int gpfy_sh_features_fwd(void* stream, const GpfyDesc* desc, int64_t n, const double* X, int apply_to_sphere,
                         const double* w, const double* b, const double* vhat, const double* lcat,
                         const double* degree_scale, int mul_r, double* out, double* r_out, void* workspace,
                         size_t workspace_bytes);

/* Reverse pass of gpfy_sh_features_fwd: what `jax.grad` builds for polynomial_expansion / Kuf / project_fun
 * (spherical_harmonics.py:290-320, 344-362, 385-387), so the feature matrix is differentiable on its own (jax.custom_vjp's bwd).
 * Same X / apply_to_sphere / w / b / vhat / lcat / degree_scale / mul_r as the forward call; dPhi [M, N] is the cotangent of `out`.
 * Outputs (each may be NULL): dvhat [M, dim] and dlcat [sum m_l^2] (lower triangles) w.r.t. the inputs `vhat`, `lcat`;
 * dscale [F] w.r.t. degree_scale; dw [d] or [1] and db [1] w.r.t. weight_variances / bias_variance (apply_to_sphere != 0 only).
 * Uses GPFY_OP_FEATURES_VJP scratch.  Not built for GPFY_W_PROJECTION weights or degree blocks wider than 112 rows. */
GPFY_API int gpfy_sh_features_vjp(void* stream, const GpfyDesc* desc, int64_t n, const double* X, int apply_to_sphere, const double* w,
                         const double* b, const double* vhat, const double* lcat, const double* degree_scale, int mul_r,
                         const double* dPhi, double* dvhat, double* dlcat, double* dscale, double* dw, double* db, void* workspace,
                         size_t workspace_bytes);

/* Fused ELBO data term and all its gradients (replaces optimization.py:125-156 `elbo` composed with
 * model.py:201-229 `predict_diag`, and the reverse pass `jax.value_and_grad` builds at
 * optimization.py:113-118), for the N rows given.  Inputs are the CONSTRAINED parameters:
 *   w [d] or [1], b [1], variance [1], eig [F] (kernel.eigenvalues at levels 0..F-1), vhat [M, dim],
 *   lcat [sum m_l^2], mu [M, P], sigma [P, M, M], sigma2 [1] (ignored for Bernoulli).
 * `packed` receives GpfyElboLayout.total doubles. */
GPFY_API int gpfy_svgp_elbo_valgrad(void* stream, const GpfyDesc* desc, int64_t n, const double* X, const double* Y,
                           const double* w, const double* b, const double* variance, const double* eig,
                           const double* vhat, const double* lcat, const double* mu, const double* sigma,
                           const double* sigma2, double* packed, void* workspace, size_t workspace_bytes);

/* The same step in three stages that share one workspace, for inputs that arrive slab by slab (host-
 * resident data streamed through the device, minibatches: optimization.py:84-104).  `n_capacity` is
 * the row count the workspace was sized for with gpfy_workspace_bytes(.., GPFY_OP_ELBO, ..) and must
 * be the same in all three calls; every `rows` call may pass any n >= 0.
 *   begin : parameter prologue, zeroed accumulators
 *   rows  : adds the terms of n rows of (X, Y)
 *   finish: writes `packed` (sums over all rows added since `begin`) */
GPFY_API int gpfy_svgp_elbo_begin(void* stream, const GpfyDesc* desc, int64_t n_capacity, const double* w, const double* b,
                         const double* variance, const double* eig, const double* vhat, const double* lcat,
                         const double* mu, const double* sigma, void* workspace, size_t workspace_bytes);
GPFY_API int gpfy_svgp_elbo_rows(void* stream, const GpfyDesc* desc, int64_t n_capacity, int64_t n, const double* X,
                        const double* Y, const double* variance, const double* sigma2, void* workspace,
                        size_t workspace_bytes);
GPFY_API int gpfy_svgp_elbo_finish(void* stream, const GpfyDesc* desc, int64_t n_capacity, const double* w, const double* b,
                          const double* variance, const double* eig, const double* mu, double* packed, void* workspace,
                          size_t workspace_bytes);

/* Inducing-basis algebra of the learned (phase-truncated) mode, N-independent, one call per step each way
 * (replaces spherical_harmonics.py:194-217 `Vs`, :219-239 `Ls`, :265-288 `orthogonalise_basis` and their reverse
 * pass).  V [M, dim] are the raw `V_l` parameters concatenated over degrees.
 *   fwd: vhat = rows of V normalised (normalise != 0), lcat = chol(alpha/(alpha+l) C_l^alpha(vhat vhat^T) + 1e-16 I)
 *   vjp: dV from the cotangents of (vhat, lcat) as gpfy_svgp_elbo_valgrad returns them (d_vhat, d_lcat).
 * Without scratch everything for a degree lives in shared memory: m_l <= 64, GPFY_EUNSUPPORTED otherwise.  The _ws forms take
 * 32 * sum_l m_l^2 bytes of 16-byte aligned scratch (the four m_l x m_l matrices of every degree) and serve m_l <= 256
 * (phase_truncation = 100 of the reference's sweeps included); with m_l <= 64 they ignore it. */
GPFY_API int gpfy_inducing_basis_fwd(void* stream, const GpfyDesc* desc, const double* V, int normalise, double* vhat, double* lcat);
GPFY_API int gpfy_inducing_basis_vjp(void* stream, const GpfyDesc* desc, const double* V, const double* lcat, int normalise,
                            const double* dvhat, const double* dlcat, double* dV);
GPFY_API int gpfy_inducing_basis_fwd_ws(void* stream, const GpfyDesc* desc, const double* V, int normalise, double* vhat, double* lcat,
                               void* workspace, size_t workspace_bytes);
GPFY_API int gpfy_inducing_basis_vjp_ws(void* stream, const GpfyDesc* desc, const double* V, const double* lcat, int normalise,
                               const double* dvhat, const double* dlcat, double* dV, void* workspace, size_t workspace_bytes);

/* KL(q || N(0, I)) and its gradients (replaces variational.py:225-240 with :262-286 / :375-399).
 * out: [1 + M*P + P*M*M] = KL, dKL/dmu, dKL/dSigma.  TriL only reads diag/all of L; full covariance
 * needs a Cholesky of Sigma, i.e. scratch: this entry returns GPFY_EUNSUPPORTED for it, gpfy_prior_kl_valgrad_ws serves both. */
GPFY_API int gpfy_prior_kl_valgrad(void* stream, const GpfyDesc* desc, const double* mu, const double* sigma, double* out);
/* The same for either q_kind, with scratch: the full-covariance KL (variational.py:375-399: logdet through
 * chol(Sigma), trace of Sigma; dKL/dSigma = 1/2 I - 1/2 Sigma^-1, the symmetrised cotangent jnp.linalg.cholesky gives) runs on
 * the device with 2 P M^2 doubles of workspace (Cholesky factor and its inverse).  TriL ignores the workspace. */
GPFY_API int gpfy_prior_kl_valgrad_ws(void* stream, const GpfyDesc* desc, const double* mu, const double* sigma, double* out,
                             void* workspace, size_t workspace_bytes);

/* Predictive mean and diagonal variance (replaces model.py:201-229 `GP.predict_diag` with the
 * conditional of :103-145): fmu, fvar are [N, P]. */
GPFY_API int gpfy_predict_diag(void* stream, const GpfyDesc* desc, int64_t n, const double* X, const double* w,
                      const double* b, const double* variance, const double* eig, const double* vhat,
                      const double* lcat, const double* mu, const double* sigma, double* fmu, double* fvar,
                      void* workspace, size_t workspace_bytes);

/* to_sphere (replaces spherical.py:214-249 `_scale_X` + `to_sphere`): xs [N, dim] unit vectors, r [N] radii. */
GPFY_API int gpfy_to_sphere(void* stream, const GpfyDesc* desc, int64_t n, const double* X, const double* w, const double* b,
                   double* xs, double* r);

/* Expected log-likelihood per point (replaces likelihoods.py:188-220 `Gaussian.variational_expectations`
 * and :254-278 `Bernoulli.variational_expectations` with quadrature.py:30-53): out [N] = sum over the P
 * outputs; fmu, fvar, Y are [N, P]; sigma2 [1] is read for the Gaussian only. */
GPFY_API int gpfy_variational_expectations(void* stream, const GpfyDesc* desc, int64_t n, const double* fmu, const double* fvar,
                                  const double* Y, const double* sigma2, double* out);

/* q-projections of a given projection matrix A [M, N] (replaces variational.py:288-305 `project_mean`,
 * :307-328 / :420-441 `project_diag_variance`): mean [N, P] = A^T mu, var [N, P] = diag(A^T S A) with
 * S = L L^T (TriL) or Sigma (full covariance).  Uses GPFY_OP_PREDICT scratch. */
GPFY_API int gpfy_project_q(void* stream, const GpfyDesc* desc, int64_t n, const double* A, const double* mu,
                   const double* sigma, double* mean, double* var, void* workspace, size_t workspace_bytes);

/* ---- device-side step driver (replaces the host-JAX work around the hot path) ----
 * theta is the flat UNCONSTRAINED parameter vector (what optax updates, model.py:285-295), `cons` the flat constrained
 * buffer the hot-path calls read from; both are device memory, offsets in doubles. */
typedef struct GpfyStepLayout {
  int32_t kernel_kind;       /* 0: eigenvalues are constants (NTK / ArcCosine: cf = lambda_l [F]);
                                1: PolynomialDecay (cf = alpha / (n + alpha) / C_n^alpha(1) [truncation_level], spherical.py:526-561) */
  int32_t truncation_level;
  int32_t w_identity;        /* 1: projection weights (Identity bijector), 0: Exp */
  int32_t n_w;               /* number of weight entries */
  int64_t th_w, th_b, th_v, th_beta, th_s2, th_mu, th_sig, th_V, th_total; /* theta; th_beta / th_s2 = -1 when absent */
  int64_t c_w, c_b, c_v, c_beta, c_eig, c_deig, c_s2, c_mu, c_sig, c_V, c_total; /* cons; c_deig: d lambda_l / d beta */
} GpfyStepLayout;

/* cons = bijector.forward(theta) (param.py:255: Exp / Identity / FillTriangular through `tri_map`, the flat [M, M]
 * position of every vector entry) plus kernel.eigenvalues (spherical.py:563-587) and its beta-derivative. */
GPFY_API int gpfy_step_constrain(void* stream, const GpfyDesc* desc, const GpfyStepLayout* lay, const double* theta,
                        const int* tri_map, const double* cf, double* cons);
/* grad = d(-ELBO)/d theta from the all-reduced packed buffer, the KL triple of gpfy_prior_kl_valgrad and dV of
 * gpfy_inducing_basis_vjp (un-scaled); scale_over_n = dataset_size / N (optimization.py:154-156); loss_out[0] = -ELBO. */
GPFY_API int gpfy_step_grad(void* stream, const GpfyDesc* desc, const GpfyStepLayout* lay, const double* cons, const double* packed,
                   const double* kl, const double* dV, const int* tri_map, double scale_over_n, double* grad, double* loss_out);
/* optax.adam with the frozen leaves masked to zero updates (model.py:288-289, training.py:78-86); step_count >= 1. */
GPFY_API int gpfy_step_adam(void* stream, int64_t n, double* theta, const double* grad, double* m1, double* m2,
                   const unsigned char* mask, double lr, double b1, double b2, double eps, int64_t step_count);

/* ---- shuffled minibatches (replaces batched_data_generator, optimization.py:31-58) ----
 * An epoch visits the rows in the order pi(0), pi(1), ..., pi(n_rows - 1), pi a bijection of [0, n_rows) keyed by
 * (seed, epoch) - a cycle-walked Feistel network, so there is no permutation array and no sort.  A batch is `batch`
 * consecutive epoch positions from `start` (the caller drops the last partial batch and bumps `epoch`, like
 * `dataset.shuffle().iter(batch_size, drop_last_batch=True)`). */
#define GPFY_SHUFFLE_ROUNDS 6
/* Xb[j, :] = X[pi(start + j), :], Yb likewise; X [n_rows, dx], Y [n_rows, dy] (dy may be 0) are device pointers or
 * pinned host pointers (read through unified addressing when the data set exceeds HBM); idx_out [batch] may be NULL. */
GPFY_API int gpfy_minibatch_gather(void* stream, uint64_t seed, uint64_t epoch, int64_t n_rows, int64_t start, int64_t batch,
                          const double* X, int dx, const double* Y, int dy, double* Xb, double* Yb, int64_t* idx_out);
/* pi(position) on the host (same function the kernel evaluates; for tests and host-side bookkeeping). */
GPFY_API int gpfy_shuffle_index(uint64_t seed, uint64_t epoch, int64_t n_rows, int64_t position, int64_t* out /* host */);

/* ---- device-resident step state: no per-step host argument, so a training step replays from a CUDA graph ----
 * state [3] int64 on the device: steps taken (Adam's bias-correction count, model.py:297-298 scan index), epoch, position
 * inside the epoch.  `_state` variants read it; gpfy_step_advance ends a step: loss_trace[steps % trace_capacity] = loss[0]
 * (the loss history `GP.fit` returns, model.py:298-300), steps += 1 and, for batch > 0, position += batch with the
 * drop-last / reshuffle rule of optimization.py:48-57. */
GPFY_API int gpfy_step_adam_state(void* stream, int64_t n, double* theta, const double* grad, double* m1, double* m2,
                         const unsigned char* mask, double lr, double b1, double b2, double eps, const int64_t* state);
GPFY_API int gpfy_minibatch_gather_state(void* stream, uint64_t seed, int64_t n_rows, int64_t batch, const int64_t* state,
                                const double* X, int dx, const double* Y, int dy, double* Xb, double* Yb, int64_t* idx_out);
GPFY_API int gpfy_step_advance(void* stream, int64_t* state, int64_t n_rows, int64_t batch, const double* loss, double* loss_trace,
                      int64_t trace_capacity);

/* ---- kernel-matrix side (dense N x N2 outputs) ---- */
#define GPFY_MAX_SHAPE_LEVELS 64
enum { GPFY_SHAPE_ARCCOSINE = 0, GPFY_SHAPE_NTK = 1, GPFY_SHAPE_POLYDECAY = 2 };
/* The shape function of a Spherical kernel (host struct, POD):
 *   ARCCOSINE: kappa_order nested `depth` times (spherical.py:438-455); NTK: spherical.py:479-500;
 *   POLYDECAY: sum_n coef[n] C_n^alpha(x), C_n interpolated linearly on linspace(-1, 1, grid_resolution) like
 *   GegenbauerLookupTable (gegenbauer.py:111-119, 152-167); coef[n] = kernel.eigenvalues[n] (n + alpha) / alpha for
 *   n < num_levels = truncation_level (spherical.py:602-611) - N-independent, computed by the caller. */
typedef struct GpfyShapeFn {
  uint32_t struct_size; /* sizeof(GpfyShapeFn) */
  int32_t kind;         /* GPFY_SHAPE_* */
  int32_t depth;        /* ArcCosine / NTK */
  int32_t order;        /* ArcCosine: 0, 1, 2 */
  int32_t num_levels;   /* PolynomialDecay truncation_level */
  int32_t grid_resolution; /* PolynomialDecay lookup grid (reference default 100000) */
  double coef[GPFY_MAX_SHAPE_LEVELS];
} GpfyShapeFn;

/* Scratch for gpfy_kernel_matrix (n1, n2 rows; n2 = 0 when X2 is NULL) and gpfy_project_variance (n1 = N, n2 = 0). */
GPFY_API int gpfy_kmat_workspace_bytes(const GpfyDesc* desc, int64_t n1, int64_t n2, size_t* bytes /* host */);
/* out[i] = shape_function(x[i]), input squashed to [-1, 1] first (replaces Spherical.shape_function of the three kernels). */
GPFY_API int gpfy_shape_function(void* stream, const GpfyShapeFn* fn /* host */, double alpha, const double* x, int64_t n, double* out);
/* K[n1, n2] = variance * shape(<xhat_i, xhat2_j>) * (r_i r2_j)^order  (replaces Spherical.K, spherical.py:318-352, and
 * Spherical.K2, :354-382).  X2 == NULL: pairs of X1; `unit_diagonal` then forces the diagonal cosine to exactly 1 as K
 * does (:340-341) - K2 passes 0. */
GPFY_API int gpfy_kernel_matrix(void* stream, const GpfyDesc* desc, const GpfyShapeFn* fn /* host */, int64_t n1, const double* X1,
                       int64_t n2, const double* X2, const double* w, const double* b, const double* variance,
                       int unit_diagonal, double* K, void* workspace, size_t workspace_bytes);
/* out[p] = base + A^T (S_p - subtract_identity * I) A, [P, N, N], with S_p = L_p L_p^T (GPFY_Q_TRIL) or Sigma_p
 * (replaces VariationalDistribution*.project_variance, variational.py:330-351 / :443-464; with base = K(X) and
 * subtract_identity = 1 it is the conditional covariance of GP.predict / GP.cov, model.py:129-131, 231-259 and
 * spherical_harmonics.py:391-393).  A is the [M, N] projection; base [N, N] may be NULL. */
GPFY_API int gpfy_project_variance(void* stream, const GpfyDesc* desc, int64_t n, const double* A, const double* sigma,
                          int subtract_identity, const double* base, double* out, void* workspace, size_t workspace_bytes);

/* Sum of the packed buffers of all ranks, in place, on `stream` (the single exchange step).
 * `nccl_comm` is an ncclComm_t owned by the caller; NCCL is resolved at run time from the process
 * (the host framework's copy), so this library has no link-time NCCL dependency. */
GPFY_API int gpfy_allreduce_packed(void* nccl_comm, void* stream, double* packed, int64_t count);

#ifdef __cplusplus
}
#endif
#endif /* GPFY_B200_H_ */
