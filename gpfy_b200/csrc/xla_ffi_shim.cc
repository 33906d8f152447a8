// XLA typed-FFI handlers over the C-ABI in include/gpfy_b200.h.
//
// The reference is a JAX program; its drop-in replacement reaches CUDA through `jax.ffi`
// (register_ffi_target + ffi_call).  The XLA FFI headers ship inside jaxlib
// (jaxlib/include/xla/ffi/api/ffi.h); jaxlib is not installable in this image, so this translation
// unit compiles to an empty object here and to the real handlers wherever the header exists
// (pass -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())")).  INTEGRATION.md shows the
// Python side.  Each handler only unpacks buffers/attributes and forwards to one C-ABI call.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define GPFY_HAVE_XLA_FFI 1
#endif
#endif

#ifdef GPFY_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include "../../include/gpfy_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static GpfyDesc make_desc(int32_t input_dim, ffi::Span<const int32_t> m, int32_t num_outputs, int32_t kernel_order,
                          int32_t likelihood, int32_t q_kind, int32_t weight_mode, double alpha) {
  GpfyDesc d{};
  d.struct_size = sizeof(GpfyDesc);
  d.abi_version = GPFY_ABI_VERSION;
  d.input_dim = input_dim;
  d.num_degrees = (int32_t)m.size();
  for (size_t i = 0; i < m.size() && i < GPFY_MAX_DEGREES; ++i) d.m[i] = m[i];
  d.num_outputs = num_outputs;
  d.kernel_order = kernel_order;
  d.likelihood = likelihood;
  d.q_kind = q_kind;
  d.weight_mode = weight_mode;
  d.alpha = alpha;
  return d;
}

static ffi::Error status(int rc) {
  if (rc == GPFY_OK) return ffi::Error::Success();
  return ffi::Error(rc == GPFY_EINVAL ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                    gpfy_last_error_string());
}

// packed = gpfy_svgp_elbo_valgrad(X, Y, w, b, variance, eig, vhat, lcat, mu, sigma, sigma2); scratch is an
// extra result buffer sized by gpfy_workspace_bytes on the Python side.
static ffi::Error ElboValGradImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> X, ffi::Buffer<ffi::F64> Y,
                                  ffi::Buffer<ffi::F64> w, ffi::Buffer<ffi::F64> b, ffi::Buffer<ffi::F64> variance,
                                  ffi::Buffer<ffi::F64> eig, ffi::Buffer<ffi::F64> vhat, ffi::Buffer<ffi::F64> lcat,
                                  ffi::Buffer<ffi::F64> mu, ffi::Buffer<ffi::F64> sigma, ffi::Buffer<ffi::F64> sigma2,
                                  ffi::ResultBuffer<ffi::F64> packed, ffi::ResultBuffer<ffi::U8> scratch,
                                  ffi::Span<const int32_t> m, int32_t kernel_order, int32_t likelihood, int32_t q_kind,
                                  int32_t weight_mode, double alpha) {
  auto dims = X.dimensions();
  GpfyDesc d = make_desc((int32_t)dims[1], m, (int32_t)Y.dimensions()[1], kernel_order, likelihood, q_kind, weight_mode, alpha);
  return status(gpfy_svgp_elbo_valgrad(stream, &d, dims[0], X.typed_data(), Y.typed_data(), w.typed_data(), b.typed_data(),
                                       variance.typed_data(), eig.typed_data(), vhat.typed_data(), lcat.typed_data(),
                                       mu.typed_data(), sigma.typed_data(), sigma2.typed_data(), packed->typed_data(),
                                       scratch->untyped_data(), scratch->size_bytes()));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(gpfy_ffi_svgp_elbo_valgrad, ElboValGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<ffi::Span<const int32_t>>("m").Attr<int32_t>("kernel_order")
                                  .Attr<int32_t>("likelihood").Attr<int32_t>("q_kind").Attr<int32_t>("weight_mode")
                                  .Attr<double>("alpha"));

static ffi::Error PredictDiagImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> X, ffi::Buffer<ffi::F64> w,
                                  ffi::Buffer<ffi::F64> b, ffi::Buffer<ffi::F64> variance, ffi::Buffer<ffi::F64> eig,
                                  ffi::Buffer<ffi::F64> vhat, ffi::Buffer<ffi::F64> lcat, ffi::Buffer<ffi::F64> mu,
                                  ffi::Buffer<ffi::F64> sigma, ffi::ResultBuffer<ffi::F64> fmu,
                                  ffi::ResultBuffer<ffi::F64> fvar, ffi::ResultBuffer<ffi::U8> scratch,
                                  ffi::Span<const int32_t> m, int32_t kernel_order, int32_t q_kind, int32_t weight_mode,
                                  double alpha) {
  auto dims = X.dimensions();
  GpfyDesc d = make_desc((int32_t)dims[1], m, (int32_t)mu.dimensions()[1], kernel_order, GPFY_LIK_GAUSSIAN, q_kind,
                         weight_mode, alpha);
  return status(gpfy_predict_diag(stream, &d, dims[0], X.typed_data(), w.typed_data(), b.typed_data(), variance.typed_data(),
                                  eig.typed_data(), vhat.typed_data(), lcat.typed_data(), mu.typed_data(), sigma.typed_data(),
                                  fmu->typed_data(), fvar->typed_data(), scratch->untyped_data(), scratch->size_bytes()));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(gpfy_ffi_predict_diag, PredictDiagImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<ffi::Span<const int32_t>>("m").Attr<int32_t>("kernel_order").Attr<int32_t>("q_kind")
                                  .Attr<int32_t>("weight_mode").Attr<double>("alpha"));

static ffi::Error FeaturesImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> X, ffi::Buffer<ffi::F64> w, ffi::Buffer<ffi::F64> b,
                               ffi::Buffer<ffi::F64> vhat, ffi::Buffer<ffi::F64> lcat, ffi::Buffer<ffi::F64> degree_scale,
                               ffi::ResultBuffer<ffi::F64> out, ffi::ResultBuffer<ffi::F64> r,
                               ffi::ResultBuffer<ffi::U8> scratch, ffi::Span<const int32_t> m, int32_t input_dim,
                               int32_t apply_to_sphere, int32_t mul_r, int32_t weight_mode, double alpha) {
  GpfyDesc d = make_desc(input_dim, m, 1, 1, GPFY_LIK_GAUSSIAN, GPFY_Q_TRIL, weight_mode, alpha);
  return status(gpfy_sh_features_fwd(stream, &d, X.dimensions()[0], X.typed_data(), apply_to_sphere, w.typed_data(),
                                     b.typed_data(), vhat.typed_data(), lcat.typed_data(), degree_scale.typed_data(), mul_r,
                                     out->typed_data(), r->typed_data(), scratch->untyped_data(), scratch->size_bytes()));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(gpfy_ffi_sh_features_fwd, FeaturesImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<ffi::Span<const int32_t>>("m").Attr<int32_t>("input_dim")
                                  .Attr<int32_t>("apply_to_sphere").Attr<int32_t>("mul_r").Attr<int32_t>("weight_mode")
                                  .Attr<double>("alpha"));
#endif  // GPFY_HAVE_XLA_FFI
