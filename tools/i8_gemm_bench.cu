// Development harness (not part of the library): times gemm_i8_kernel alone on random digit planes and, built with
// -DGI8_PROFILE, prints where the MMA and epilogue warps spend their cycles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo [-DGI8_PROFILE] -I gpfy_b200/csrc -I include -o tools/i8_gemm_bench tools/i8_gemm_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
namespace gpfy { void set_error(const char*, ...) {} void count_launch() {} int debug_option(int) { return 0; } }
#include "gemm_i8.cuh"
using namespace gpfy;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void fill(uint32_t* p, size_t n, uint32_t seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t x = (uint32_t)i * 2654435761u + seed; x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
    p[i] = x;
  }
}

int main(int argc, char** argv) {
  const int Mp = argc > 1 ? atoi(argv[1]) : 288;
  const int64_t npts = argc > 2 ? atoll(argv[2]) : 909312;
  const int with_c = argc > 3 ? atoi(argv[3]) : 1;
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  GemmI8Args a{};
  int nsa;
  if (!gi8_plan(Mp, sms, &a.blk, &nsa)) { printf("no plan\n"); return 1; }
  if (argc > 4) nsa = atoi(argv[4]);
  const size_t zb = gi8_z_image_bytes(Mp, npts), wb = gi8_w_image_bytes(Mp);
  int8_t *Zs, *Ws; double *rs, *C;
  CK(cudaMalloc(&Zs, zb)); CK(cudaMalloc(&Ws, wb)); CK(cudaMalloc(&rs, Mp * 8)); CK(cudaMalloc(&C, (size_t)Mp * npts * 8));
  fill<<<1024, 256>>>((uint32_t*)Zs, zb / 4, 1); fill<<<64, 256>>>((uint32_t*)Ws, wb / 4, 2);
  std::vector<double> h(Mp, 1.0); CK(cudaMemcpy(rs, h.data(), Mp * 8, cudaMemcpyHostToDevice));
  a.Zs = Zs; a.Ws = Ws; a.rs = rs; a.Z = nullptr; a.ldz = npts; a.C = with_c ? C : nullptr; a.psi = nullptr; a.ld_psi = npts; a.Mp = Mp; a.npts = npts; a.nsa = nsa;
#ifdef GI8_PROFILE
  long long* prof; CK(cudaMalloc(&prof, sms * 8 * 8)); CK(cudaMemset(prof, 0, sms * 8 * 8)); a.prof = prof;
#endif
  const int smem = gi8_smem_bytes(Mp, a.blk, nsa);
  CK(cudaFuncSetAttribute(gemm_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = a.blk.cta0[a.blk.nblk];
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int r = 0; r < 4; ++r) {
#ifdef GI8_PROFILE
    CK(cudaMemset(prof, 0, sms * 8 * 8));
#endif
    cudaEventRecord(e0);
    gemm_i8_kernel<<<grid, GI8_THREADS, smem>>>(a);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double macs = 21.0 * Mp * (double)Mp * npts;
    printf("Mp %d npts %lld nsa %d C %d: %.3f ms  %.0f MAC/clk/SM at 1.965 GHz  (%.1f ms per 10 M points)\n", Mp, (long long)npts, nsa, with_c, ms,
           macs / (ms * 1e-3 * 1.965e9 * sms), ms * 1e7 / npts);
  }
#ifdef GI8_PROFILE
  std::vector<long long> p(sms * 8); CK(cudaMemcpy(p.data(), prof, sms * 8 * 8, cudaMemcpyDeviceToHost));
  for (int b : {0, a.blk.cta0[1], a.blk.cta0[a.blk.nblk - 1], grid - 1}) {
    const long long* q = &p[b * 8];
    printf("cta %3d: tiles %lld  per tile: total %.0f  wait A %.0f  wait D %.0f | epilogue warp: wait d_full %.0f  drain+store %.0f\n", b, q[3],
           (double)q[0] / q[3], (double)q[1] / q[3], (double)q[2] / q[3], (double)q[4] / q[3], (double)q[5] / q[3]);
  }
#endif
  return 0;
}
