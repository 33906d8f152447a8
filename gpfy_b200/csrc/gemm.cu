#include "gemm.cuh"

namespace gpfy {

// Producer/consumer pipeline without CTA-wide barriers: every thread copies its share of a stage with
// cp.async and signals the stage's FULL mbarrier when its copies land; a warp waits on FULL, issues
// its DMMAs, and arrives on the stage's EMPTY mbarrier; a stage is refilled only after all 12 warps
// released it.  Warps may therefore drift by up to one k-step and keep the FP64 tensor pipe fed.
// VAR is 0 in production; tools/gemm_bench.cu instantiates knock-out variants to attribute time:
//   bit 0: no global loads (stages are signalled full without copying)   bit 1: no epilogue at all
//   bit 2: epilogue without the psi reduction                          bit 3: epilogue without the C stores
template <int VAR>
__device__ __forceinline__ void gemm_body(const GemmArgs& a) {
  extern __shared__ __align__(16) double smem[];
  __shared__ uint64_t bar_full[GEMM_STAGES], bar_empty[GEMM_STAGES];
  double* As = smem;                                           // [STAGES][BM][AS]
  double* Bs = smem + GEMM_STAGES * GEMM_BM * GEMM_AS;         // [STAGES][BK][BS]
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int wm = warp & 3, wn = warp >> 2;
  const int g = lane >> 2, t4 = lane & 3;

  const int m_tiles = (a.m + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = (a.n + GEMM_BN - 1) / GEMM_BN;
  const int tiles = m_tiles * n_tiles;
  const int m_eff = a.m_eff > 0 ? a.m_eff : a.m;
  const int kt_full = (a.k_eff > 0 ? a.k_eff : a.k) / GEMM_BK;
  if ((int)blockIdx.x >= tiles) return;
  // lower_k: A is lower triangular in 96-row blocks, so row tile t only needs k < 96 (t + 1)
  auto kt_of = [&](int tile) {
    const int lim = ((tile % m_tiles) + 1) * (GEMM_BM / GEMM_BK);
    return (a.lower_k && lim < kt_full) ? lim : kt_full;
  };
  int total = 0;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) total += kt_of(t);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < GEMM_STAGES; ++s) {
      mbar_init(&bar_full[s], GEMM_THREADS);
      mbar_init(&bar_empty[s], GEMM_THREADS / 32);
    }
  }
  __syncthreads();

  // ---- producer state: this thread's 2 A chunks and 4 B chunks (16 bytes each) of every stage ----
  const int a_r0 = tid >> 3, a_kc = (tid & 7) * 2;          // rows a_r0 and a_r0 + 48
  const int b_r0 = tid / 96, b_nc = (tid % 96) * 2;         // rows b_r0 + {0, 4, 8, 12}
  const double* pA[2];
  const double* pB[4];
  bool okA[2], okB;
  int p_tile = blockIdx.x, p_kb = 0, pit = 0, p_kt = kt_of(blockIdx.x);
  auto producer_tile = [&]() {
    const int row0 = (p_tile % m_tiles) * GEMM_BM, col0 = (p_tile / m_tiles) * GEMM_BN;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      okA[i] = (row0 + a_r0 + i * 48) < a.m;
      pA[i] = okA[i] ? a.A + (int64_t)(row0 + a_r0 + i * 48) * a.lda + a_kc : a.A;
    }
    okB = (col0 + b_nc) < a.n;
#pragma unroll
    for (int i = 0; i < 4; ++i) pB[i] = okB ? a.B + (int64_t)(b_r0 + i * 4) * a.ldb + col0 + b_nc : a.B;
  };
  producer_tile();
  auto produce = [&]() {
    if (pit < total) {
      const int st = pit % GEMM_STAGES;
      if (pit >= GEMM_STAGES) mbar_wait(&bar_empty[st], ((pit / GEMM_STAGES) - 1) & 1);
      double* as = As + st * GEMM_BM * GEMM_AS + a_r0 * GEMM_AS + a_kc;
      double* bs = Bs + st * GEMM_BK * GEMM_BS + b_r0 * GEMM_BS + b_nc;
      if (!(VAR & 1)) {
#pragma unroll
        for (int i = 0; i < 2; ++i) cp_async16(as + i * 48 * GEMM_AS, pA[i], okA[i]);
#pragma unroll
        for (int i = 0; i < 4; ++i) cp_async16(bs + i * 4 * GEMM_BS, pB[i], okB);
      }
      cp_async_mbar_arrive(&bar_full[st]);
      ++pit;
      if (++p_kb == p_kt) {
        p_kb = 0;
        p_tile += gridDim.x;
        if (pit < total) {
          producer_tile();
          p_kt = kt_of(p_tile);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 2; ++i) if (okA[i]) pA[i] += GEMM_BK;
        if (okB) {
#pragma unroll
          for (int i = 0; i < 4; ++i) pB[i] += (int64_t)GEMM_BK * a.ldb;
        }
      }
    }
  };
#pragma unroll
  for (int s = 0; s < GEMM_STAGES - 1; ++s) produce();

  // A warp's three 8-row sub-tiles are rows r * 32 + wm * 8 (+ g) of the CTA tile - interleaved, so that a partial last row
  // tile (m = 352 or 256: 64 of 96 rows) costs every warp two of its three sub-tiles instead of idling a warp row.
  double acc[3][8][2];
  int c_tile = blockIdx.x, c_kb = 0, c_kt = kt_of(blockIdx.x);
  int row0 = (c_tile % m_tiles) * GEMM_BM, col0 = (c_tile / m_tiles) * GEMM_BN;
  bool active = (row0 + wm * 8) < a.m && (col0 + wn * 64) < a.n;
  bool ract[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) ract[r] = (row0 + r * 32 + wm * 8) < m_eff;
  for (int it = 0; it < total; ++it) {
    const int st = it % GEMM_STAGES;
    if (c_kb == 0) {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;
    }
    mbar_wait(&bar_full[st], (it / GEMM_STAGES) & 1);
    const double* as = As + st * GEMM_BM * GEMM_AS + (wm * 8 + g) * GEMM_AS + t4;
    const double* bs = Bs + st * GEMM_BK * GEMM_BS + t4 * GEMM_BS + wn * 64 + g;
#pragma unroll
    for (int k4 = 0; k4 < GEMM_BK / 4; ++k4) {
      if (active) {
        double af[3], bf[8];
#pragma unroll
        for (int r = 0; r < 3; ++r) af[r] = as[r * 32 * GEMM_AS + k4 * 4];
#pragma unroll
        for (int c = 0; c < 8; ++c) bf[c] = bs[k4 * 4 * GEMM_BS + c * 8];
#pragma unroll
        for (int r = 0; r < 3; ++r)
          if (ract[r]) {
#pragma unroll
            for (int c = 0; c < 8; ++c) dmma884(acc[r][c][0], acc[r][c][1], af[r], bf[c]);
          }
      }
      if (k4 == 0) produce();  // the copy issue hides behind the DMMAs just queued
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar_empty[st]);

    if (++c_kb == c_kt) {
      double ps[8][2];
#pragma unroll
      for (int c = 0; c < 8; ++c) ps[c][0] = ps[c][1] = 0.0;
      if (active && !(VAR & 2)) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int row = row0 + r * 32 + wm * 8 + g;
          if (row < a.m) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const int col = col0 + wn * 64 + c * 8 + 2 * t4;
              if (col < a.n) {
                double2* dst = a.c_blocked
                                   ? reinterpret_cast<double2*>(a.C + ((int64_t)(col >> 5) * a.m + row) * 32 + ((col & 31) ^ ((row & 3) << 2)))
                                   : reinterpret_cast<double2*>(a.C + (int64_t)row * a.ldc + col);
                double2 v;
                v.x = a.alpha * acc[r][c][0];
                v.y = a.alpha * acc[r][c][1];
                if (a.beta != 0.0) {
                  double2 o = *dst;
                  v.x += a.beta * o.x;
                  v.y += a.beta * o.y;
                }
                if ((!(VAR & 8) || v.x == 1.2345e300) && a.C) *dst = v;  // VAR bit 3: keep the MMAs, drop the stores; C == null: psi only
                if (a.psi_out && !(VAR & 4)) {
                  const double2 z = *reinterpret_cast<const double2*>((a.psi_B ? a.psi_B : a.B) + (int64_t)row * a.ldb + col);
                  ps[c][0] = fma(z.x, v.x, ps[c][0]);
                  ps[c][1] = fma(z.y, v.y, ps[c][1]);
                }
              }
            }
          }
        }
      }
      if (a.psi_out && !(VAR & 6) && (col0 + wn * 64) < a.n) {
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
      c_kt = kt_of(c_tile);
      row0 = (c_tile % m_tiles) * GEMM_BM;
      col0 = (c_tile / m_tiles) * GEMM_BN;
      active = (row0 + wm * 8) < a.m && (col0 + wn * 64) < a.n;
#pragma unroll
      for (int r = 0; r < 3; ++r) ract[r] = (row0 + r * 32 + wm * 8) < m_eff;
    }
  }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_dmma_kernel(GemmArgs a) { gemm_body<0>(a); }

int launch_gemm(cudaStream_t stream, const GemmArgs& a, int num_sms) {
  if ((a.m_eff & 7) || (a.k_eff % GEMM_BK) || a.m_eff > a.m || a.k_eff > a.k) { set_error("launch_gemm: m_eff must be a multiple of 8, k_eff of 16"); return GPFY_EINVAL; }
  if (a.k % GEMM_BK != 0 || (a.n & 1) || (a.lda & 1) || (a.ldb & 1) || (a.ldc & 1)) {
    set_error("launch_gemm: k=%d must be a multiple of 16 and n=%d, lda, ldb, ldc even", a.k, a.n);
    return GPFY_EINVAL;
  }
  if (a.c_blocked && a.beta != 0.0) { set_error("launch_gemm: the blocked C layout needs beta = 0"); return GPFY_EINVAL; }
  if (a.psi_out && (a.m != a.k || a.alpha != 1.0 || a.beta != 0.0 || (a.ld_psi & 1))) {
    set_error("launch_gemm: the psi epilogue needs m == k, alpha = 1, beta = 0 and an even ld_psi");
    return GPFY_EINVAL;
  }
  GPFY_TRY(device_configure_once(DEV_CFG_GEMM, (const void*)gemm_dmma_kernel, GEMM_SMEM));
  int64_t tiles = (int64_t)((a.m + GEMM_BM - 1) / GEMM_BM) * ((a.n + GEMM_BN - 1) / GEMM_BN);
  if (tiles == 0) return GPFY_OK;
  int grid = (int)(tiles < num_sms ? tiles : num_sms);
  gemm_dmma_kernel<<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(a);
  GPFY_COUNT_LAUNCH();
  GPFY_LAUNCH_OK();
  return GPFY_OK;
}

}  // namespace gpfy
