#include "gemm.cuh"

namespace gpfy {

__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_dmma_kernel(GemmArgs a) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;                                           // [STAGES][BM][AS]
  double* Bs = smem + GEMM_STAGES * GEMM_BM * GEMM_AS;         // [STAGES][BK][BS]
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int wm = warp & 3, wn = warp >> 2;
  const int g = lane >> 2, t4 = lane & 3;

  const int m_tiles = (a.m + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = (a.n + GEMM_BN - 1) / GEMM_BN;
  const int tiles = m_tiles * n_tiles;
  const int kt = a.k / GEMM_BK;
  if ((int)blockIdx.x >= tiles) return;
  const int n_my = (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = n_my * kt;

  // producer state: iteration `pit` of the flattened (tile, kb) stream
  int p_tile = blockIdx.x, p_kb = 0, pit = 0;
  auto load_next = [&]() {
    if (pit < total) {
      const int st = pit % GEMM_STAGES;
      const int row0 = (p_tile % m_tiles) * GEMM_BM, col0 = (p_tile / m_tiles) * GEMM_BN;
      double* as = As + st * GEMM_BM * GEMM_AS;
#pragma unroll
      for (int i = 0; i < 2; ++i) {  // A: 96 rows x 8 chunks of 16 B
        int c = tid + i * GEMM_THREADS;
        int r = c >> 3, kc = (c & 7) * 2;
        bool p = (row0 + r) < a.m;
        const double* src = p ? a.A + (int64_t)(row0 + r) * a.lda + p_kb * GEMM_BK + kc : a.A;
        cp_async16(as + r * GEMM_AS + kc, src, p);
      }
      double* bs = Bs + st * GEMM_BK * GEMM_BS;
#pragma unroll
      for (int i = 0; i < 4; ++i) {  // B: 16 rows x 96 chunks
        int c = tid + i * GEMM_THREADS;
        int r = c / 96, nc = (c % 96) * 2;
        bool p = (col0 + nc) < a.n;
        const double* src = p ? a.B + (int64_t)(p_kb * GEMM_BK + r) * a.ldb + col0 + nc : a.B;
        cp_async16(bs + r * GEMM_BS + nc, src, p);
      }
      ++pit;
      if (++p_kb == kt) { p_kb = 0; p_tile += gridDim.x; }
    }
    cp_async_commit();
  };

  double acc[3][8][2];
#pragma unroll
  for (int s = 0; s < GEMM_STAGES - 1; ++s) load_next();

  int c_tile = blockIdx.x, c_kb = 0;
  for (int it = 0; it < total; ++it) {
    cp_async_wait<GEMM_STAGES - 2>();
    __syncthreads();
    load_next();
    if (c_kb == 0) {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;
    }
    const int row0 = (c_tile % m_tiles) * GEMM_BM, col0 = (c_tile / m_tiles) * GEMM_BN;
    const bool active = (row0 + wm * 24) < a.m && (col0 + wn * 64) < a.n;
    if (active) {
      const int st = it % GEMM_STAGES;
      const double* as = As + st * GEMM_BM * GEMM_AS + (wm * 24 + g) * GEMM_AS + t4;
      const double* bs = Bs + st * GEMM_BK * GEMM_BS + t4 * GEMM_BS + wn * 64 + g;
#pragma unroll
      for (int k4 = 0; k4 < GEMM_BK / 4; ++k4) {
        double af[3], bf[8];
#pragma unroll
        for (int r = 0; r < 3; ++r) af[r] = as[r * 8 * GEMM_AS + k4 * 4];
#pragma unroll
        for (int c = 0; c < 8; ++c) bf[c] = bs[k4 * 4 * GEMM_BS + c * 8];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 8; ++c) dmma884(acc[r][c][0], acc[r][c][1], af[r], bf[c]);
      }
    }
    if (++c_kb == kt) {
      double ps[8][2];
#pragma unroll
      for (int c = 0; c < 8; ++c) ps[c][0] = ps[c][1] = 0.0;
      if (active) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int row = row0 + wm * 24 + r * 8 + g;
          if (row < a.m) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const int col = col0 + wn * 64 + c * 8 + 2 * t4;
              if (col < a.n) {
                double2* dst = reinterpret_cast<double2*>(a.C + (int64_t)row * a.ldc + col);
                double2 v;
                v.x = a.alpha * acc[r][c][0];
                v.y = a.alpha * acc[r][c][1];
                if (a.beta != 0.0) {
                  double2 o = *dst;
                  v.x += a.beta * o.x;
                  v.y += a.beta * o.y;
                }
                *dst = v;
                if (a.psi_out) {
                  const double2 z = *reinterpret_cast<const double2*>(a.B + (int64_t)row * a.ldb + col);
                  ps[c][0] = fma(z.x, v.x, ps[c][0]);
                  ps[c][1] = fma(z.y, v.y, ps[c][1]);
                }
              }
            }
          }
        }
      }
      if (a.psi_out && (col0 + wn * 64) < a.n) {
        double* dst = a.psi_out + (int64_t)((c_tile % m_tiles) * 4 + wm) * a.ld_psi;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
#pragma unroll
          for (int o = 4; o <= 16; o <<= 1) {
            ps[c][0] += __shfl_xor_sync(0xffffffffu, ps[c][0], o);
            ps[c][1] += __shfl_xor_sync(0xffffffffu, ps[c][1], o);
          }
          const int col = col0 + wn * 64 + c * 8 + 2 * t4;
          if (g == 0 && col < a.n) *reinterpret_cast<double2*>(dst + col) = make_double2(ps[c][0], ps[c][1]);
        }
      }
      c_kb = 0;
      c_tile += gridDim.x;
    }
  }
  cp_async_wait<0>();
}

int launch_gemm(cudaStream_t stream, const GemmArgs& a, int num_sms) {
  if (a.k % GEMM_BK != 0 || (a.n & 1) || (a.lda & 1) || (a.ldb & 1) || (a.ldc & 1)) {
    set_error("launch_gemm: k=%d must be a multiple of 16 and n=%d, lda, ldb, ldc even", a.k, a.n);
    return GPFY_EINVAL;
  }
  if (a.psi_out && (a.m != a.k || a.alpha != 1.0 || a.beta != 0.0 || (a.ld_psi & 1))) {
    set_error("launch_gemm: the psi epilogue needs m == k, alpha = 1, beta = 0 and an even ld_psi");
    return GPFY_EINVAL;
  }
  static bool configured = false;
  if (!configured) {
    GPFY_CUDA_OK(cudaFuncSetAttribute(gemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    configured = true;
  }
  int64_t tiles = (int64_t)((a.m + GEMM_BM - 1) / GEMM_BM) * ((a.n + GEMM_BN - 1) / GEMM_BN);
  if (tiles == 0) return GPFY_OK;
  int grid = (int)(tiles < num_sms ? tiles : num_sms);
  gemm_dmma_kernel<<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(a);
  GPFY_COUNT_LAUNCH();
  GPFY_LAUNCH_OK();
  return GPFY_OK;
}

}  // namespace gpfy
