// Times the production GEMM body (gpfy_b200/csrc/gemm.cu) and knock-out variants on the headline
// shape (m = k = 288, n = 113664) to attribute the gap to the DMMA peak.  Development aid.
#include <cstdio>
#include <cstdarg>
#include <vector>
#include "../gpfy_b200/csrc/gemm.cu"
namespace gpfy { void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); }
unsigned long long g_launch_count = 0; }
using namespace gpfy;
template <int VAR> __global__ void __launch_bounds__(GEMM_THREADS, 1) bench_kernel(GemmArgs a) { gemm_body<VAR>(a); }
template <int VAR> double run(const GemmArgs& a, int reps) {
  cudaFuncSetAttribute(bench_kernel<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) bench_kernel<VAR><<<148, GEMM_THREADS, GEMM_SMEM>>>(a);
  cudaDeviceSynchronize(); cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) bench_kernel<VAR><<<148, GEMM_THREADS, GEMM_SMEM>>>(a);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
  return ms / reps * 1e3;
}
int main() {
  const int M = 288; const int64_t N = 113664;
  double *A, *B, *C, *psi;
  cudaMalloc(&A, M * M * 8); cudaMalloc(&B, (M + 32) * N * 8); cudaMalloc(&C, M * N * 8); cudaMalloc(&psi, 12 * N * 8);
  cudaMemset(A, 0, M * M * 8); cudaMemset(B, 0, (M + 32) * N * 8);
  GemmArgs a; a.A = A; a.lda = M; a.B = B; a.ldb = N; a.C = C; a.ldc = N; a.m = M; a.k = M; a.n = (int)N; a.alpha = 1; a.beta = 0;
  a.psi_out = psi; a.ld_psi = N;
  const double fl = 2.0 * M * M * N;
  double t;
  t = run<0>(a, 5); printf("full              %8.1f us  %5.2f TF\n", t, fl / t * 1e-6);
  t = run<4>(a, 5); printf("no psi            %8.1f us  %5.2f TF\n", t, fl / t * 1e-6);
  t = run<12>(a, 5); printf("no psi, no stores %8.1f us  %5.2f TF\n", t, fl / t * 1e-6);
  a.c_blocked = 1; t = run<4>(a, 5); printf("no psi, blocked C %8.1f us  %5.2f TF\n", t, fl / t * 1e-6); a.c_blocked = 0;
  t = run<1>(a, 5); printf("no loads          %8.1f us  %5.2f TF\n", t, fl / t * 1e-6);
  t = run<13>(a, 5); printf("no loads no stores%8.1f us  %5.2f TF\n", t, fl / t * 1e-6);
  return 0;
}
