// FP64 tensor-core GEMM (DMMA.8x8x4) for the dense contractions of the SVGP path:
//   C[m, n] = alpha * A[m, k] * B[k, n] + beta * C          (all row-major)
// Used for  C = W2 * Z  (the L_q^T A / L_q (L_q^T A) contractions of variational.py:323-328 and
// their VJP, folded — see DESIGN.md) and for the O(M^3) per-step parameter algebra.
//
// CTA tile 96 x 128, BK = 16, 3-stage cp.async pipeline, 6 warps (3 x 2), warp tile 32 x 64:
// per k4 step a warp loads 4 A + 8 B fragments (12 LDS.64) for 32 DMMA.  Shared-memory rows are
// padded by 4 doubles so every fragment load is bank-conflict-free (row stride = 4 mod 16 doubles).
// Persistent over tiles: grid = min(#tiles, 2 CTAs x #SM); row tiles vary fastest so that CTAs
// running together read the same B panel from L2.
#pragma once
#include "common.cuh"

namespace gpfy {

struct GemmArgs {
  const double* A;
  int64_t lda;
  const double* B;
  int64_t ldb;
  double* C;
  int64_t ldc;
  int m, n, k;  // k must be a multiple of 16 (operands zero padded); n even; lda, ldb, ldc even
  double alpha, beta;
};

constexpr int GEMM_BM = 96, GEMM_BN = 128, GEMM_BK = 16, GEMM_STAGES = 3, GEMM_THREADS = 192;
constexpr int GEMM_AS = GEMM_BK + 4;   // 20 doubles: A-fragment loads conflict-free
constexpr int GEMM_BS = GEMM_BN + 4;   // 132 doubles: B-fragment loads conflict-free
constexpr int GEMM_SMEM = GEMM_STAGES * (GEMM_BM * GEMM_AS + GEMM_BK * GEMM_BS) * 8;

int launch_gemm(cudaStream_t stream, const GemmArgs& a, int num_sms);

}  // namespace gpfy
