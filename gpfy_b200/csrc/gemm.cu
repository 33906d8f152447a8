#include "gemm.cuh"

namespace gpfy {

__global__ void __launch_bounds__(GEMM_THREADS, 2) gemm_dmma_kernel(GemmArgs a) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;                                           // [STAGES][BM][AS]
  double* Bs = smem + GEMM_STAGES * GEMM_BM * GEMM_AS;         // [STAGES][BK][BS]
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int wm = warp % 3, wn = warp / 3;
  const int g = lane >> 2, t4 = lane & 3;

  const int m_tiles = (a.m + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = (a.n + GEMM_BN - 1) / GEMM_BN;
  const int kt = a.k / GEMM_BK;

  for (int tile = blockIdx.x; tile < m_tiles * n_tiles; tile += gridDim.x) {
    const int tm = tile % m_tiles, tn = tile / m_tiles;
    const int row0 = tm * GEMM_BM, col0 = tn * GEMM_BN;

    auto load_stage = [&](int st, int kb) {
      // A: 96 rows x 16 doubles = 768 16-byte chunks
      double* as = As + st * GEMM_BM * GEMM_AS;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int c = tid + i * GEMM_THREADS;
        int r = c >> 3, kc = (c & 7) * 2;
        bool p = (row0 + r) < a.m;
        const double* src = p ? a.A + (int64_t)(row0 + r) * a.lda + kb * GEMM_BK + kc : a.A;
        cp_async16(as + r * GEMM_AS + kc, src, p);
      }
      // B: 16 rows x 128 doubles = 1024 chunks
      double* bs = Bs + st * GEMM_BK * GEMM_BS;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        int c = tid + i * GEMM_THREADS;
        if (c < 1024) {
          int r = c >> 6, nc = (c & 63) * 2;
          bool p = (col0 + nc) < a.n;
          const double* src = p ? a.B + (int64_t)(kb * GEMM_BK + r) * a.ldb + col0 + nc : a.B;
          cp_async16(bs + r * GEMM_BS + nc, src, p);
        }
      }
    };

    double acc[4][8][2];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;

#pragma unroll
    for (int s = 0; s < GEMM_STAGES - 1; ++s) {
      if (s < kt) load_stage(s, s);
      cp_async_commit();
    }
    const bool active = (row0 + wm * 32) < a.m;
    for (int kb = 0; kb < kt; ++kb) {
      cp_async_wait<GEMM_STAGES - 2>();
      __syncthreads();
      if (kb + GEMM_STAGES - 1 < kt) load_stage((kb + GEMM_STAGES - 1) % GEMM_STAGES, kb + GEMM_STAGES - 1);
      cp_async_commit();
      if (active) {
        const double* as = As + (kb % GEMM_STAGES) * GEMM_BM * GEMM_AS + (wm * 32 + g) * GEMM_AS + t4;
        const double* bs = Bs + (kb % GEMM_STAGES) * GEMM_BK * GEMM_BS + t4 * GEMM_BS + wn * 64 + g;
#pragma unroll
        for (int k4 = 0; k4 < GEMM_BK / 4; ++k4) {
          double af[4], bf[8];
#pragma unroll
          for (int r = 0; r < 4; ++r) af[r] = as[r * 8 * GEMM_AS + k4 * 4];
#pragma unroll
          for (int c = 0; c < 8; ++c) bf[c] = bs[k4 * 4 * GEMM_BS + c * 8];
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) dmma884(acc[r][c][0], acc[r][c][1], af[r], bf[c]);
        }
      }
    }
    cp_async_wait<0>();
    __syncthreads();  // all warps done with smem before the next tile's prologue overwrites it

    if (active) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        int row = row0 + wm * 32 + r * 8 + g;
        if (row < a.m) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            int col = col0 + wn * 64 + c * 8 + 2 * t4;
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
            }
          }
        }
      }
    }
  }
}

int launch_gemm(cudaStream_t stream, const GemmArgs& a, int num_sms) {
  if (a.k % GEMM_BK != 0 || (a.n & 1) || (a.lda & 1) || (a.ldb & 1) || (a.ldc & 1)) {
    set_error("launch_gemm: k=%d must be a multiple of 16 and n=%d, lda, ldb, ldc even", a.k, a.n);
    return GPFY_EINVAL;
  }
  static bool configured = false;
  if (!configured) {
    GPFY_CUDA_OK(cudaFuncSetAttribute(gemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM));
    configured = true;
  }
  int64_t tiles = (int64_t)((a.m + GEMM_BM - 1) / GEMM_BM) * ((a.n + GEMM_BN - 1) / GEMM_BN);
  if (tiles == 0) return GPFY_OK;
  int grid = (int)(tiles < 2 * num_sms ? tiles : 2 * num_sms);
  gemm_dmma_kernel<<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(a);
  GPFY_COUNT_LAUNCH();
  GPFY_LAUNCH_OK();
  return GPFY_OK;
}

}  // namespace gpfy
