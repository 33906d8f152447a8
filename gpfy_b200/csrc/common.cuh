// Shared host/device helpers for the gpfy_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gpfy_b200.h"

namespace gpfy {

// ------------------------------------------------------------------------------------------------
// host side: errors
// ------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define GPFY_CUDA_OK(expr)                                                                  \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      ::gpfy::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return GPFY_ECUDA;                                                                    \
    }                                                                                       \
  } while (0)

#define GPFY_LAUNCH_OK()                                                                    \
  do {                                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                   \
    if (e__ != cudaSuccess) {                                                               \
      ::gpfy::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return GPFY_ECUDA;                                                                    \
    }                                                                                       \
  } while (0)

#define GPFY_TRY(expr)              \
  do {                              \
    int rc__ = (expr);              \
    if (rc__ != GPFY_OK) return rc__; \
  } while (0)

static inline int64_t round_up64(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Launch counter (bench.py reports "gpu_launches" from it).
void count_launch();
#define GPFY_COUNT_LAUNCH() (::gpfy::count_launch())

// Per-device state (a host framework such as XLA drives several GPUs from one process, possibly from several
// threads).  Everything that CUDA keeps per device - the SM count, cudaFuncSetAttribute(MaxDynamicSharedMemorySize) of the
// non-templated kernels - is cached here per device ordinal under a mutex, keyed by cudaGetDevice() of the calling thread.
enum { DEV_CFG_GEMM = 1u << 0, DEV_CFG_SYRK = 1u << 1, DEV_CFG_GEMM32 = 1u << 2, DEV_CFG_SYRK32 = 1u << 3 };
int device_sms();                                             // SM count of the calling thread's current device
int device_configure_once(unsigned bit, const void* kernel, int smem_bytes);  // cudaFuncSetAttribute once per device

// Development switches (gpfy_debug_set_option; never read from the environment, default 0 = production path).
enum { DBG_CHUNK_MULT = 0, DBG_OLD_FWD = 1, DBG_NO_FUSED_BWD = 2, DBG_OLD_FEATURES = 3, DBG_NO_BWD4 = 4, DBG_NO_SYRK32 = 5, DBG_SYRK32_KF = 6, DBG_NO_I8 = 7, DBG_NO_SYRK_SLICE = 8, DBG_NO_FEATURES3 = 9, DBG_NO_FWD3 = 10, DBG_COUNT = 12 };
int debug_option(int which);

// ------------------------------------------------------------------------------------------------
// Gegenbauer three-term recurrence coefficients, passed by value as a kernel parameter so the
// uniform operands sit in the constant bank (gegenbauer.py:27-60):
//   C_k = (2 x (k + a - 1) C_{k-1} - (k + 2a - 2) C_{k-2}) / k  =  c1[k] x C_{k-1} - c2[k] C_{k-2}
// d*/two_alpha_d: the same for alpha + 1, used by d/dx C_n^a = 2 a C_{n-1}^{a+1} (gegenbauer.py:73).
// ------------------------------------------------------------------------------------------------
struct RecCoef {
  double c1[GPFY_MAX_DEGREES + 1];
  double c2[GPFY_MAX_DEGREES + 1];
  double d1[GPFY_MAX_DEGREES + 1];
  double d2[GPFY_MAX_DEGREES + 1];
  double two_alpha;    // 2 a
  double two_alpha_p1; // 2 (a + 1)
};

static inline RecCoef make_rec_coef(double alpha) {
  RecCoef rc;
  for (int k = 0; k <= GPFY_MAX_DEGREES; ++k) {
    double kk = k < 1 ? 1.0 : (double)k;
    rc.c1[k] = 2.0 * (kk + alpha - 1.0) / kk;
    rc.c2[k] = (kk + 2.0 * alpha - 2.0) / kk;
    rc.d1[k] = 2.0 * (kk + (alpha + 1.0) - 1.0) / kk;
    rc.d2[k] = (kk + 2.0 * (alpha + 1.0) - 2.0) / kk;
  }
  rc.two_alpha = 2.0 * alpha;
  rc.two_alpha_p1 = 2.0 * (alpha + 1.0);
  return rc;
}

#ifdef __CUDACC__
// C_level^alpha(t)
__device__ __forceinline__ double geg_eval(const RecCoef& rc, int level, double t) {
  if (level == 0) return 1.0;
  double cp = 1.0, c = rc.two_alpha * t;
  for (int k = 2; k <= level; ++k) {
    double cn = fma(rc.c1[k] * t, c, -rc.c2[k] * cp);
    cp = c;
    c = cn;
  }
  return c;
}

// d/dt C_level^alpha(t) = 2 alpha C_{level-1}^{alpha+1}(t); 0 at level 0
__device__ __forceinline__ double geg_deriv(const RecCoef& rc, int level, double t) {
  if (level == 0) return 0.0;
  if (level == 1) return rc.two_alpha;
  double cp = 1.0, c = rc.two_alpha_p1 * t;
  for (int k = 2; k <= level - 1; ++k) {
    double cn = fma(rc.d1[k] * t, c, -rc.d2[k] * cp);
    cp = c;
    c = cn;
  }
  return rc.two_alpha * c;
}

// Degree known at compile time: fully unrolled, coefficients become constant-bank operands.
template <int L>
__device__ __forceinline__ double geg_eval_s(const RecCoef& rc, double t) {
  if constexpr (L == 0) {
    return 1.0;
  } else {
    double cp = 1.0, c = rc.two_alpha * t;
#pragma unroll
    for (int k = 2; k <= L; ++k) {
      double cn = fma(rc.c1[k] * t, c, -rc.c2[k] * cp);
      cp = c;
      c = cn;
    }
    return c;
  }
}
template <int L>
__device__ __forceinline__ double geg_deriv_s(const RecCoef& rc, double t) {
  if constexpr (L == 0) {
    return 0.0;
  } else if constexpr (L == 1) {
    return rc.two_alpha;
  } else {
    double cp = 1.0, c = rc.two_alpha_p1 * t;
#pragma unroll
    for (int k = 2; k <= L - 1; ++k) {
      double cn = fma(rc.d1[k] * t, c, -rc.d2[k] * cp);
      cp = c;
      c = cn;
    }
    return rc.two_alpha * c;
  }
}
// The legacy one-point-per-lane zonal kernels (fallback for P > 1 / large M / large dim) once carried a fully
// unrolled code path per degree; the tensor-core kernels made that obsolete, so the switch now only selects the
// runtime-degree loop (LS = -1), which also keeps the 72 template instantiations quick to compile.
#define GPFY_DEGREE_SWITCH(l, ...) { constexpr int LS = -1; __VA_ARGS__; }
// LS = -1 selects the runtime-degree path
template <int LS>
__device__ __forceinline__ double geg_eval_d(const RecCoef& rc, int l, double t) {
  if constexpr (LS >= 0) return geg_eval_s<LS>(rc, t);
  else return geg_eval(rc, l, t);
}
template <int LS>
__device__ __forceinline__ double geg_deriv_d(const RecCoef& rc, int l, double t) {
  if constexpr (LS >= 0) return geg_deriv_s<LS>(rc, t);
  else return geg_deriv(rc, l, t);
}

// ------------------------------------------------------------------------------------------------
// FP64 tensor-core MMA: D(8x8) += A(8x4) B(4x8).  SASS: DMMA.8x8x4 (the only f64 MMA shape sm_100a
// has; tcgen05 has no f64 kind).  lane l: a = A[l/4][l%4], b = B[l%4][l/4], c0/c1 = C[l/4][2(l%4)+{0,1}].
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------------
// cp.async (LDGSTS) 16-byte copies with zero fill, and group control
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ------------------------------------------------------------------------------------------------
// TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP + SYNCS).
// Used to stage the replicated parameter panels (fundamental points, mu2) into shared memory.
// bytes must be a multiple of 16, both addresses 16-byte aligned.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(b), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(bytes));
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem, const void* gmem, unsigned bytes, uint64_t* bar) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(s),
               "l"(gmem), "r"(bytes), "r"(b)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
  unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(b), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// cp.async completion tracked by an mbarrier: the barrier gets one arrival from this thread once all
// of its prior cp.async copies have landed (".noinc": the expected count must include it).
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
  unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(b) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(b) : "memory");
}

// Stage `bytes` from global to shared with one elected thread issuing TMA bulk copies (<= 64 KiB
// each is not a limit of the instruction, but we chunk to keep expect_tx below 2^20).  All threads
// of the CTA must call this; on return the data is visible to all of them.
__device__ __forceinline__ void stage_params_tma(void* smem, const void* gmem, unsigned bytes, uint64_t* bar) {
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, bytes);
    unsigned off = 0;
    while (off < bytes) {
      unsigned n = bytes - off < 32768u ? bytes - off : 32768u;
      tma_bulk_g2s((char*)smem + off, (const char*)gmem + off, n, bar);
      off += n;
    }
  }
  mbar_wait(bar, 0);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif  // __CUDACC__

// ------------------------------------------------------------------------------------------------
// Shape bookkeeping derived from a descriptor (host).
// ------------------------------------------------------------------------------------------------
struct Shape {
  int d, dim, dimp;  // input dim, ambient dim, ambient dim rounded up to a multiple of 4
  int sd, proj, nw;  // sphere dim (= proj or d), projection dim (0: none), number of weight entries (d, 1 or d * p)
  int F, P, M, Mp;   // degrees, outputs, inducing features, M rounded up to a multiple of 32
  int off[GPFY_MAX_DEGREES + 1];   // row offset of each degree
  int64_t loff[GPFY_MAX_DEGREES + 1];  // offset of L_l inside lcat
  int64_t lsize;     // sum m_l^2
  int mmax;
};
int make_shape(const GpfyDesc* desc, Shape* s);

}  // namespace gpfy
