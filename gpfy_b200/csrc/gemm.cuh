// FP64 tensor-core GEMM (DMMA.8x8x4) for the dense contractions of the SVGP path:
//   C[m, n] = alpha * A[m, k] * B[k, n] + beta * C          (all row-major)
// Used for  C = W2 * Z  (the L_q^T A / L_q (L_q^T A) contractions of variational.py:323-328 and
// their VJP, folded — see DESIGN.md) and for the O(M^3) per-step parameter algebra.
//
// CTA tile 96 x 192, BK = 16, 4-stage cp.async pipeline, 12 warps (4 x 3, three per SM sub-partition
// so the four FP64 pipes carry equal work), warp tile 24 x 64: per k4 step a warp loads 3 A + 8 B
// fragments (11 LDS.64) for 24 DMMA.  Shared-memory rows are padded by 4 doubles so every fragment
// load is bank-conflict-free (row stride = 4 mod 16 doubles).  Persistent: one CTA per SM walks its
// tiles as ONE flattened stream of k-steps, so the loads of the next tile are in flight while the
// current tile finishes; row tiles vary fastest so CTAs running together share the B panel in L2.
#pragma once
#include "common.cuh"

namespace gpfy {

struct GemmArgs {
  const double* A;
  int64_t lda;
  const double* B;
  int64_t ldb;
  double* C;    // may be null when only the psi epilogue is wanted (predict_diag): the product is not stored
  int64_t ldc;
  int m, n, k;  // k must be a multiple of 16 (operands zero padded); n even; lda, ldb, ldc even.  C / psi_out are written for
                // whole 192-column tiles (zeros beyond n): their panels must hold ceil(n / 192) * 192 columns
  double alpha, beta;
  // optional fused epilogue (needs m == k, alpha = 1, beta = 0): per-column partial sums of
  // B[i, n] * C[i, n] over the rows of each (row tile, warp row) slab,
  //   psi_out[(tile_m * 4 + warp_m) * ld_psi + n] = sum_{i in slab} B[i, n] C[i, n]
  // i.e. the pieces of z_n^T W2 z_n (variational.py:326-327 folded); fixed layout => deterministic.
  double* psi_out;
  int64_t ld_psi;
  const double* psi_B = nullptr;   // panel whose rows are dotted with C in the psi epilogue (same layout / ld as B); null = B
  // c_blocked != 0: C is written as [n / 32][m][32] (each 32-column block of all m rows contiguous, so the
  // fused backward kernel fetches a tile with a few TMA bulk copies) with the column inside a block XOR-
  // swizzled by the row, (col & 31) ^ ((row & 3) << 2), which makes the tensor-core fragment reads of that
  // tile bank-conflict-free without padding.  ldc is ignored; needs beta = 0.
  int c_blocked = 0;
  // lower_k != 0: A is lower triangular at the granularity of the 96-row tiles (everything right of the diagonal
  // block is zero), so row tile t stops its k loop at 96 (t + 1): a third fewer k-steps for three row tiles.  Used by
  // predict_diag, where only psi_n = z_n^T W z_n is wanted and W can be folded onto its lower triangle.
  int lower_k = 0;
  // Rows of A at and beyond m_eff and rows of B (k) at and beyond k_eff are known to be zero (the padding of M up to the
  // next multiple of 32): their 8-row sub-tiles / 16-deep k-steps are skipped (C still receives the zeros).  0 = m / k.
  int m_eff = 0, k_eff = 0;
};

static inline int gemm_psi_slabs(int m) { return ((m + 95) / 96) * 4; }

constexpr int GEMM_BM = 96, GEMM_BN = 192, GEMM_BK = 16, GEMM_STAGES = 4, GEMM_THREADS = 384;
constexpr int GEMM_AS = GEMM_BK + 4;   // 20 doubles: A-fragment loads conflict-free
constexpr int GEMM_BS = GEMM_BN + 4;   // 196 doubles: B-fragment loads conflict-free
constexpr int GEMM_SMEM = GEMM_STAGES * (GEMM_BM * GEMM_AS + GEMM_BK * GEMM_BS) * 8;

int launch_gemm(cudaStream_t stream, const GemmArgs& a, int num_sms);

}  // namespace gpfy
