// XLA typed-FFI handlers over the C-ABI in include/gpfy_b200.h - the `jax.ffi` layer BASELINE.json's north_star prescribes.
//
// The reference is a JAX program; its drop-in replacement reaches CUDA through `jax.ffi` (register_ffi_target + ffi_call).
// The XLA FFI headers ship inside jaxlib (jaxlib/include/xla/ffi/api/ffi.h); jaxlib is not installable in this image, so this
// translation unit compiles to an empty object here and to the real handlers wherever the header exists
// (pass -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())")).  tests/test_ffi_shim.py type-checks every handler
// against tests/ffi_stub/xla/ffi/api/ffi.h, a minimal restatement of the public FFI C++ API (binding chain <-> handler
// signature, buffer accessors), so the file is at least compiled in this image; it is EXPERIMENTAL until it has been built
// against a real jaxlib.  INTEGRATION.md holds the Python side (gpfy/_b200.py).
//
// Every handler: unpack buffers / static attributes -> check ranks -> rebuild the POD GpfyDesc -> ONE C-ABI call on XLA's
// stream.  No allocation (scratch is an extra U8 result sized with gpfy_workspace_bytes on the Python side), no
// synchronisation, no exception; errors return as ffi::Error carrying gpfy_last_error_string().  The library keeps its per-
// device state keyed by device ordinal under a mutex (common.cuh), so handlers may run concurrently for different devices.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define GPFY_HAVE_XLA_FFI 1
#endif
#endif

#ifdef GPFY_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include <string>

#include "../../include/gpfy_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using F64 = ffi::Buffer<ffi::F64>;
using RF64 = ffi::ResultBuffer<ffi::F64>;
using RU8 = ffi::ResultBuffer<ffi::U8>;
using I32s = ffi::Span<const int32_t>;

namespace {

ffi::Error status(int rc) {
  if (rc == GPFY_OK) return ffi::Error::Success();
  const ffi::ErrorCode code = rc == GPFY_EINVAL ? ffi::ErrorCode::kInvalidArgument
                              : rc == GPFY_EUNSUPPORTED ? ffi::ErrorCode::kUnimplemented
                              : rc == GPFY_EWORKSPACE ? ffi::ErrorCode::kResourceExhausted
                                                      : ffi::ErrorCode::kInternal;
  return ffi::Error(code, gpfy_last_error_string());
}
ffi::Error invalid(const std::string& what) { return ffi::Error(ffi::ErrorCode::kInvalidArgument, what); }

// The static (non-pytree) configuration of (SphericalHarmonics, kernel, q, likelihood): passed as attributes of every call.
struct Static {
  I32s m;
  int32_t input_dim, num_outputs, kernel_order, likelihood, q_kind, weight_mode, projection_dim, precision;
  double alpha;
};
bool make_desc(const Static& s, GpfyDesc* d, ffi::Error* err) {
  if (s.m.size() == 0 || s.m.size() > GPFY_MAX_DEGREES) {
    *err = invalid("attribute m must hold 1.." + std::to_string(GPFY_MAX_DEGREES) + " degrees");
    return false;
  }
  *d = GpfyDesc{};
  d->struct_size = sizeof(GpfyDesc);
  d->abi_version = GPFY_ABI_VERSION;
  d->input_dim = s.input_dim;
  d->num_degrees = (int32_t)s.m.size();
  for (size_t i = 0; i < s.m.size(); ++i) d->m[i] = s.m[i];
  d->num_outputs = s.num_outputs;
  d->kernel_order = s.kernel_order;
  d->likelihood = s.likelihood;
  d->q_kind = s.q_kind;
  d->weight_mode = s.weight_mode;
  d->projection_dim = s.weight_mode == GPFY_W_PROJECTION ? s.projection_dim : 0;
  d->precision = s.precision;
  d->alpha = s.alpha;
  return true;
}
bool rank_is(const F64& b, size_t rank) { return b.dimensions().size() == rank; }

}  // namespace

// The attribute block shared by all handlers (the Python side passes the same **static dict everywhere).
#define GPFY_STATIC_ATTRS(B)                                                                                          \
  B.Attr<I32s>("m").Attr<int32_t>("input_dim").Attr<int32_t>("num_outputs").Attr<int32_t>("kernel_order")             \
      .Attr<int32_t>("likelihood").Attr<int32_t>("q_kind").Attr<int32_t>("weight_mode").Attr<int32_t>("projection_dim") \
      .Attr<int32_t>("precision").Attr<double>("alpha")
#define GPFY_STATIC_PARAMS                                                                                             \
  I32s m, int32_t input_dim, int32_t num_outputs, int32_t kernel_order, int32_t likelihood, int32_t q_kind,            \
      int32_t weight_mode, int32_t projection_dim, int32_t precision, double alpha
#define GPFY_STATIC_INIT                                                                                               \
  Static st{m, input_dim, num_outputs, kernel_order, likelihood, q_kind, weight_mode, projection_dim, precision, alpha}; \
  GpfyDesc d;                                                                                                          \
  ffi::Error err = ffi::Error::Success();                                                                              \
  if (!make_desc(st, &d, &err)) return err

// ---- polynomial_expansion / Kuf / project_fun  (spherical_harmonics.py:290-320, 344-362, 385-387) ----
static ffi::Error FeaturesFwdImpl(cudaStream_t stream, F64 X, F64 w, F64 b, F64 vhat, F64 lcat, F64 degree_scale, RF64 out, RF64 r,
                                  RU8 scratch, int32_t apply_to_sphere, int32_t mul_r, GPFY_STATIC_PARAMS) {
  GPFY_STATIC_INIT;
  if (!rank_is(X, 2) || !rank_is(vhat, 2)) return invalid("sh_features_fwd: X and vhat must be rank 2");
  return status(gpfy_sh_features_fwd(stream, &d, X.dimensions()[0], X.typed_data(), apply_to_sphere, w.typed_data(), b.typed_data(),
                                     vhat.typed_data(), lcat.typed_data(), degree_scale.typed_data(), mul_r, out->typed_data(),
                                     r->typed_data(), scratch->untyped_data(), scratch->size_bytes()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(gpfy_ffi_sh_features_fwd, FeaturesFwdImpl,
                              GPFY_STATIC_ATTRS(ffi::Ffi::Bind()
                                                    .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                                    .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                                    .Ret<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::U8>>()
                                                    .Attr<int32_t>("apply_to_sphere").Attr<int32_t>("mul_r")));

// ---- its reverse pass (jax.custom_vjp bwd of the three functions above) ----
static ffi::Error FeaturesVjpImpl(cudaStream_t stream, F64 X, F64 w, F64 b, F64 vhat, F64 lcat, F64 degree_scale, F64 dPhi, RF64 dvhat,
                                  RF64 dlcat, RF64 dscale, RF64 dw, RF64 db, RU8 scratch, int32_t apply_to_sphere, int32_t mul_r,
                                  GPFY_STATIC_PARAMS) {
  GPFY_STATIC_INIT;
  if (!rank_is(X, 2) || !rank_is(dPhi, 2)) return invalid("sh_features_vjp: X and dPhi must be rank 2");
  return status(gpfy_sh_features_vjp(stream, &d, X.dimensions()[0], X.typed_data(), apply_to_sphere, w.typed_data(), b.typed_data(),
                                     vhat.typed_data(), lcat.typed_data(), degree_scale.typed_data(), mul_r, dPhi.typed_data(),
                                     dvhat->typed_data(), dlcat->typed_data(), dscale->typed_data(), dw->typed_data(), db->typed_data(),
                                     scratch->untyped_data(), scratch->size_bytes()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(gpfy_ffi_sh_features_vjp, FeaturesVjpImpl,
                              GPFY_STATIC_ATTRS(ffi::Ffi::Bind()
                                                    .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                                    .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                                    .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::U8>>()
                                                    .Attr<int32_t>("apply_to_sphere").Attr<int32_t>("mul_r")));

// ---- elbo o predict_diag + value_and_grad  (optimization.py:113-118, 125-156) ----
static ffi::Error ElboValGradImpl(cudaStream_t stream, F64 X, F64 Y, F64 w, F64 b, F64 variance, F64 eig, F64 vhat, F64 lcat, F64 mu,
                                  F64 sigma, F64 sigma2, RF64 packed, RU8 scratch, GPFY_STATIC_PARAMS) {
  GPFY_STATIC_INIT;
  if (!rank_is(X, 2) || !rank_is(Y, 2) || !rank_is(mu, 2) || !rank_is(sigma, 3)) return invalid("svgp_elbo_valgrad: X [N, d], Y [N, P], mu [M, P], Sigma [P, M, M]");
  if (Y.dimensions()[0] != X.dimensions()[0] || Y.dimensions()[1] != num_outputs) return invalid("svgp_elbo_valgrad: Y must be [N, num_outputs]");
  return status(gpfy_svgp_elbo_valgrad(stream, &d, X.dimensions()[0], X.typed_data(), Y.typed_data(), w.typed_data(), b.typed_data(),
                                       variance.typed_data(), eig.typed_data(), vhat.typed_data(), lcat.typed_data(), mu.typed_data(),
                                       sigma.typed_data(), sigma2.typed_data(), packed->typed_data(), scratch->untyped_data(),
                                       scratch->size_bytes()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(gpfy_ffi_svgp_elbo_valgrad, ElboValGradImpl,
                              GPFY_STATIC_ATTRS(ffi::Ffi::Bind()
                                                    .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                                    .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                                    .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                                    .Ret<F64>().Ret<ffi::Buffer<ffi::U8>>()));

// ---- GP.predict_diag  (model.py:201-229) ----
static ffi::Error PredictDiagImpl(cudaStream_t stream, F64 X, F64 w, F64 b, F64 variance, F64 eig, F64 vhat, F64 lcat, F64 mu, F64 sigma,
                                  RF64 fmu, RF64 fvar, RU8 scratch, GPFY_STATIC_PARAMS) {
  GPFY_STATIC_INIT;
  if (!rank_is(X, 2) || !rank_is(mu, 2) || !rank_is(sigma, 3)) return invalid("predict_diag: X [N, d], mu [M, P], Sigma [P, M, M]");
  return status(gpfy_predict_diag(stream, &d, X.dimensions()[0], X.typed_data(), w.typed_data(), b.typed_data(), variance.typed_data(),
                                  eig.typed_data(), vhat.typed_data(), lcat.typed_data(), mu.typed_data(), sigma.typed_data(),
                                  fmu->typed_data(), fvar->typed_data(), scratch->untyped_data(), scratch->size_bytes()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(gpfy_ffi_predict_diag, PredictDiagImpl,
                              GPFY_STATIC_ATTRS(ffi::Ffi::Bind()
                                                    .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                                    .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                                    .Arg<F64>().Arg<F64>()
                                                    .Ret<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::U8>>()));

// ---- gegenbauer / gegenbauer_fwd  (gegenbauer.py:27-85) ----
static ffi::Error GegenbauerImpl(cudaStream_t stream, F64 x, RF64 out, RF64 dout, int32_t level, double alpha) {
  return status(gpfy_gegenbauer(stream, level, alpha, x.typed_data(), (int64_t)x.element_count(), out->typed_data(), dout->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(gpfy_ffi_gegenbauer, GegenbauerImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>().Ret<F64>().Ret<F64>()
                                  .Attr<int32_t>("level").Attr<double>("alpha"));

// ---- _scale_X + to_sphere  (spherical.py:214-249) ----
static ffi::Error ToSphereImpl(cudaStream_t stream, F64 X, F64 w, F64 b, RF64 xs, RF64 r, GPFY_STATIC_PARAMS) {
  GPFY_STATIC_INIT;
  if (!rank_is(X, 2)) return invalid("to_sphere: X must be [N, d]");
  return status(gpfy_to_sphere(stream, &d, X.dimensions()[0], X.typed_data(), w.typed_data(), b.typed_data(), xs->typed_data(), r->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(gpfy_ffi_to_sphere, ToSphereImpl,
                              GPFY_STATIC_ATTRS(ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>().Arg<F64>().Arg<F64>()
                                                    .Ret<F64>().Ret<F64>()));

// ---- Gaussian / Bernoulli variational_expectations  (likelihoods.py:188-220, 254-278) ----
static ffi::Error VarExpImpl(cudaStream_t stream, F64 fmu, F64 fvar, F64 Y, F64 sigma2, RF64 out, GPFY_STATIC_PARAMS) {
  GPFY_STATIC_INIT;
  if (!rank_is(fmu, 2)) return invalid("variational_expectations: fmu must be [N, P]");
  return status(gpfy_variational_expectations(stream, &d, fmu.dimensions()[0], fmu.typed_data(), fvar.typed_data(), Y.typed_data(),
                                              sigma2.typed_data(), out->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(gpfy_ffi_variational_expectations, VarExpImpl,
                              GPFY_STATIC_ATTRS(ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>().Arg<F64>().Arg<F64>()
                                                    .Arg<F64>().Ret<F64>()));

// ---- project_mean / project_diag_variance  (variational.py:288-328, 420-441) ----
static ffi::Error ProjectQImpl(cudaStream_t stream, F64 A, F64 mu, F64 sigma, RF64 mean, RF64 var, RU8 scratch, GPFY_STATIC_PARAMS) {
  GPFY_STATIC_INIT;
  if (!rank_is(A, 2)) return invalid("project_q: A must be [M, N]");
  return status(gpfy_project_q(stream, &d, A.dimensions()[1], A.typed_data(), mu.typed_data(), sigma.typed_data(), mean->typed_data(),
                               var->typed_data(), scratch->untyped_data(), scratch->size_bytes()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(gpfy_ffi_project_q, ProjectQImpl,
                              GPFY_STATIC_ATTRS(ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>().Arg<F64>().Arg<F64>()
                                                    .Ret<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::U8>>()));

// ---- prior_KL with logdet / trace, either q  (variational.py:225-240, 262-286, 375-399) ----
static ffi::Error PriorKlImpl(cudaStream_t stream, F64 mu, F64 sigma, RF64 out, RU8 scratch, GPFY_STATIC_PARAMS) {
  GPFY_STATIC_INIT;
  return status(gpfy_prior_kl_valgrad_ws(stream, &d, mu.typed_data(), sigma.typed_data(), out->typed_data(), scratch->untyped_data(),
                                         scratch->size_bytes()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(gpfy_ffi_prior_kl_valgrad, PriorKlImpl,
                              GPFY_STATIC_ATTRS(ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>().Arg<F64>()
                                                    .Ret<F64>().Ret<ffi::Buffer<ffi::U8>>()));

// ---- Vs / Ls / orthogonalise_basis and their reverse pass  (spherical_harmonics.py:194-239, 265-288) ----
// scratch: 32 * sum m_l^2 bytes when some m_l > 64 (the per-degree matrices then live in global memory), any size otherwise
static ffi::Error BasisFwdImpl(cudaStream_t stream, F64 V, RF64 vhat, RF64 lcat, RU8 scratch, int32_t normalise, GPFY_STATIC_PARAMS) {
  GPFY_STATIC_INIT;
  return status(gpfy_inducing_basis_fwd_ws(stream, &d, V.typed_data(), normalise, vhat->typed_data(), lcat->typed_data(),
                                           scratch->untyped_data(), scratch->size_bytes()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(gpfy_ffi_inducing_basis_fwd, BasisFwdImpl,
                              GPFY_STATIC_ATTRS(ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>().Ret<F64>().Ret<F64>()
                                                    .Ret<ffi::Buffer<ffi::U8>>().Attr<int32_t>("normalise")));
static ffi::Error BasisVjpImpl(cudaStream_t stream, F64 V, F64 lcat, F64 dvhat, F64 dlcat, RF64 dV, RU8 scratch, int32_t normalise,
                               GPFY_STATIC_PARAMS) {
  GPFY_STATIC_INIT;
  return status(gpfy_inducing_basis_vjp_ws(stream, &d, V.typed_data(), lcat.typed_data(), normalise, dvhat.typed_data(), dlcat.typed_data(),
                                           dV->typed_data(), scratch->untyped_data(), scratch->size_bytes()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(gpfy_ffi_inducing_basis_vjp, BasisVjpImpl,
                              GPFY_STATIC_ATTRS(ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>().Arg<F64>().Arg<F64>()
                                                    .Arg<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::U8>>().Attr<int32_t>("normalise")));

// ---- project_variance / the conditional covariance of GP.predict  (variational.py:330-351, 443-464; model.py:231-259) ----
static ffi::Error ProjectVarianceImpl(cudaStream_t stream, F64 A, F64 sigma, F64 base, RF64 out, RU8 scratch, int32_t subtract_identity,
                                      int32_t has_base, GPFY_STATIC_PARAMS) {
  GPFY_STATIC_INIT;
  if (!rank_is(A, 2)) return invalid("project_variance: A must be [M, N]");
  return status(gpfy_project_variance(stream, &d, A.dimensions()[1], A.typed_data(), sigma.typed_data(), subtract_identity,
                                      has_base ? base.typed_data() : nullptr, out->typed_data(), scratch->untyped_data(),
                                      scratch->size_bytes()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(gpfy_ffi_project_variance, ProjectVarianceImpl,
                              GPFY_STATIC_ATTRS(ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>().Arg<F64>().Arg<F64>()
                                                    .Ret<F64>().Ret<ffi::Buffer<ffi::U8>>()
                                                    .Attr<int32_t>("subtract_identity").Attr<int32_t>("has_base")));
#endif  // GPFY_HAVE_XLA_FFI
