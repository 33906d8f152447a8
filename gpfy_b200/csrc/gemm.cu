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
  const int m_eff = a.m_eff > 0 ? a.m_eff : a.m;
  const int kt_full = (a.k_eff > 0 ? a.k_eff : a.k) / GEMM_BK;
  if ((int)blockIdx.x >= n_tiles) return;
  // Tile order: a CTA owns column tiles blockIdx, blockIdx + grid, ... and walks ALL row tiles of a column tile before it
  // moves on, so the [k x 192] panel of B comes from HBM once and from L2 for the other row tiles (and for the psi epilogue),
  // and every CTA carries the same mix of full and partial row tiles.  (Round 1 let row tiles vary fastest ACROSS CTAs: with
  // four row tiles of which one is partial the CTAs drifted apart and B was read 3.3 x from HBM - ncu r02m, M_p = 352.)
  // The "tile" below is col_tile * m_tiles + row_tile.
  const int my_cols = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int my_tiles = my_cols * m_tiles;
  auto tile_at = [&](int i) { return ((int)blockIdx.x + (i / m_tiles) * (int)gridDim.x) * m_tiles + (i % m_tiles); };
  // lower_k: A is lower triangular in 96-row blocks, so row tile t only needs k < 96 (t + 1)
  auto kt_of = [&](int tile) {
    const int lim = ((tile % m_tiles) + 1) * (GEMM_BM / GEMM_BK);
    return (a.lower_k && lim < kt_full) ? lim : kt_full;
  };
  int total = 0;
  for (int i = 0; i < my_tiles; ++i) total += kt_of(tile_at(i));

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < GEMM_STAGES; ++s) {
      mbar_init(&bar_full[s], GEMM_THREADS);
      mbar_init(&bar_empty[s], GEMM_THREADS / 32);
    }
  }
  __syncthreads();

  // ---- producer state: this thread's 2 A chunks and 4 B chunks (16 bytes each) of every stage.  Only the two base pointers
  // are carried (the other chunks are fixed strides away): six live pointers spilled into the k-loop once the epilogue grew ----
  const int a_r0 = tid >> 3, a_kc = (tid & 7) * 2;          // rows a_r0 and a_r0 + 48
  const int b_r0 = tid / 96, b_nc = (tid % 96) * 2;         // rows b_r0 + {0, 4, 8, 12}
  const int64_t a_step = 48 * a.lda, b_step = 4 * a.ldb;
  const double* pA0;
  const double* pB0;
  bool okA0, okA1, okB;
  int p_idx = 0, p_tile = tile_at(0), p_kb = 0, pit = 0, p_kt = kt_of(p_tile);
  auto producer_tile = [&]() {
    const int row0 = (p_tile % m_tiles) * GEMM_BM, col0 = (p_tile / m_tiles) * GEMM_BN;
    okA0 = (row0 + a_r0) < a.m;
    okA1 = (row0 + a_r0 + 48) < a.m;
    pA0 = a.A + (okA0 ? (int64_t)(row0 + a_r0) * a.lda : 0) + a_kc;   // a row past the end is never read (size 0): keep the address valid
    okB = (col0 + b_nc) < a.n;
    pB0 = a.B + (int64_t)b_r0 * a.ldb + (okB ? col0 + b_nc : 0);
  };
  producer_tile();
  auto produce = [&]() {
    if (pit < total) {
      const int st = pit % GEMM_STAGES;
      if (pit >= GEMM_STAGES) mbar_wait(&bar_empty[st], ((pit / GEMM_STAGES) - 1) & 1);
      double* as = As + st * GEMM_BM * GEMM_AS + a_r0 * GEMM_AS + a_kc;
      double* bs = Bs + st * GEMM_BK * GEMM_BS + b_r0 * GEMM_BS + b_nc;
      if (!(VAR & 1)) {
        cp_async16(as, pA0, okA0);
        cp_async16(as + 48 * GEMM_AS, okA1 ? pA0 + a_step : pA0, okA1);
#pragma unroll
        for (int i = 0; i < 4; ++i) cp_async16(bs + i * 4 * GEMM_BS, pB0 + i * b_step, okB);
      }
      cp_async_mbar_arrive(&bar_full[st]);
      ++pit;
      if (++p_kb == p_kt) {
        p_kb = 0;
        p_tile = tile_at(++p_idx);
        if (pit < total) {
          producer_tile();
          p_kt = kt_of(p_tile);
        }
      } else {
        pA0 += GEMM_BK;
        pB0 += (int64_t)GEMM_BK * a.ldb;
      }
    }
  };
#pragma unroll
  for (int s = 0; s < GEMM_STAGES - 1; ++s) produce();

  // A warp's three 8-row sub-tiles are rows r * 32 + wm * 8 (+ g) of the CTA tile - interleaved, so that a partial last row
  // tile (m = 352 or 256: 64 of 96 rows) costs every warp two of its three sub-tiles instead of idling a warp row.
  double acc[3][8][2];
  int c_idx = 0, c_tile = tile_at(0), c_kb = 0, c_kt = kt_of(c_tile);
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
      if (!(VAR & 2) && active) {
        // No per-element bounds checks: the output panels are allocated for whole 192-column tiles (launch_gemm checks it) and
        // columns at and beyond n hold zeros (their B columns are zero-filled by the producer).  Addresses are one base per
        // 8-row sub-tile plus compile-time offsets.  The round-1 epilogue tested every element and rebuilt every address:
        // ~2500 instructions per warp and tile, a third of the main loop's count (ncu r02m).
        const int swz = (g & 3) << 2;                       // (row & 3) << 2: row0, r * 32 and wm * 8 are multiples of 4
        const int64_t blk = (int64_t)a.m * 32;              // one 32-column block of the blocked layout
        const double* zsrc = a.psi_B ? a.psi_B : a.B;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int rb = row0 + r * 32 + wm * 8;
          if (rb < a.m) {                                   // uniform over the warp (m is a multiple of 8)
            const int row = rb + g;
            if (a.C) {
              if (a.c_blocked) {
                double* base = a.C + ((int64_t)((col0 + wn * 64) >> 5) * a.m + row) * 32;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                  *reinterpret_cast<double2*>(base + (c >> 2) * blk + ((((c & 3) << 3) + 2 * t4) ^ swz)) = make_double2(acc[r][c][0], acc[r][c][1]);
              } else {
                double* base = a.C + (int64_t)row * a.ldc + col0 + wn * 64 + 2 * t4;
#pragma unroll
                for (int c = 0; c < 8; ++c) *reinterpret_cast<double2*>(base + c * 8) = make_double2(acc[r][c][0], acc[r][c][1]);
              }
            }
            if (a.psi_out) {
              const double* zb = zsrc + (int64_t)row * a.ldb + col0 + wn * 64 + 2 * t4;
#pragma unroll
              for (int c0 = 0; c0 < 8; c0 += 4) {           // four loads in flight at a time (register budget)
                double2 z[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) z[c] = *reinterpret_cast<const double2*>(zb + (c0 + c) * 8);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  ps[c0 + c][0] = fma(z[c].x, acc[r][c0 + c][0], ps[c0 + c][0]);
                  ps[c0 + c][1] = fma(z[c].y, acc[r][c0 + c][1], ps[c0 + c][1]);
                }
              }
            }
          }
        }
      }
      if (a.psi_out && !(VAR & 2) && active) {
        double* dst = a.psi_out + (int64_t)((c_tile % m_tiles) * 4 + wm) * a.ld_psi;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
#pragma unroll
          for (int o = 4; o <= 16; o <<= 1) {
            ps[c][0] += __shfl_xor_sync(0xffffffffu, ps[c][0], o);
            ps[c][1] += __shfl_xor_sync(0xffffffffu, ps[c][1], o);
          }
          const int col = col0 + wn * 64 + c * 8 + 2 * t4;
          if (g == 0) *reinterpret_cast<double2*>(dst + col) = make_double2(ps[c][0], ps[c][1]);
        }
      }
      c_kb = 0;
      c_tile = tile_at(++c_idx);
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
  if (a.alpha != 1.0 || a.beta != 0.0) { set_error("launch_gemm: only alpha = 1, beta = 0 is built"); return GPFY_EINVAL; }
  if (a.psi_out && (a.m != a.k || (a.ld_psi & 1))) {
    set_error("launch_gemm: the psi epilogue needs m == k and an even ld_psi");
    return GPFY_EINVAL;
  }
  // the epilogue writes whole 192-column tiles: the output panels must have that many columns
  const int64_t cols_out = (int64_t)((a.n + GEMM_BN - 1) / GEMM_BN) * GEMM_BN;
  if ((a.C && !a.c_blocked && a.ldc < cols_out) || (a.psi_out && a.ld_psi < cols_out)) {
    set_error("launch_gemm: output panels need ld >= %lld (whole 192-column tiles)", (long long)cols_out);
    return GPFY_EINVAL;
  }
  GPFY_TRY(device_configure_once(DEV_CFG_GEMM, (const void*)gemm_dmma_kernel, GEMM_SMEM));
  const int64_t col_tiles = (a.n + GEMM_BN - 1) / GEMM_BN;   // a CTA walks all row tiles of its column tiles
  if (col_tiles == 0 || a.m == 0) return GPFY_OK;
  int grid = (int)(col_tiles < num_sms ? col_tiles : num_sms);
  gemm_dmma_kernel<<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(a);
  GPFY_COUNT_LAUNCH();
  GPFY_LAUNCH_OK();
  return GPFY_OK;
}

}  // namespace gpfy
