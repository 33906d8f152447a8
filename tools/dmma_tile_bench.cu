// Does DMMA.8x8x4 sustain peak with a realistic register tile (3 A x 8 B fragments, 24 accumulators)
// and with fragments re-loaded from shared memory every k4 step?  Development microbenchmark.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
constexpr int ITERS = 2048;
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int RA, int RB>
__global__ void __launch_bounds__(384, 1) reg_tile(double* out, double s) {
  double acc[RA][RB][2], af[RA], bf[RB];
  for (int r = 0; r < RA; ++r) af[r] = s * (r + 1) + threadIdx.x * 1e-9;
  for (int c = 0; c < RB; ++c) bf[c] = s * (c + 2);
  for (int r = 0; r < RA; ++r) for (int c = 0; c < RB; ++c) acc[r][c][0] = acc[r][c][1] = 0;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < RA; ++r)
#pragma unroll
      for (int c = 0; c < RB; ++c) dmma(acc[r][c][0], acc[r][c][1], af[r], bf[c]);
  }
  double t = 0;
  for (int r = 0; r < RA; ++r) for (int c = 0; c < RB; ++c) t += acc[r][c][0] + acc[r][c][1];
  if (t == 123.456) out[0] = t;
}
// fragments from shared memory each k4 (layout like the GEMM: A [96][20], B [16][196])
template <int RA, int RB, int WARPS_M>
__global__ void __launch_bounds__(384, 1) smem_tile(double* out, double s) {
  __shared__ double As[96 * 20];
  __shared__ double Bs[16 * 196];
  for (int i = threadIdx.x; i < 96 * 20; i += blockDim.x) As[i] = s * i;
  for (int i = threadIdx.x; i < 16 * 196; i += blockDim.x) Bs[i] = s * (i + 1);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
  const int wm = warp % WARPS_M, wn = warp / WARPS_M;
  double acc[RA][RB][2];
  for (int r = 0; r < RA; ++r) for (int c = 0; c < RB; ++c) acc[r][c][0] = acc[r][c][1] = 0;
  const double* as = As + (wm * RA * 8 + g) * 20 + t4;
  const double* bs = Bs + t4 * 196 + wn * RB * 8 + g;
  for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
      double af[RA], bf[RB];
#pragma unroll
      for (int r = 0; r < RA; ++r) af[r] = as[r * 8 * 20 + k4 * 4];
#pragma unroll
      for (int c = 0; c < RB; ++c) bf[c] = bs[k4 * 4 * 196 + c * 8];
#pragma unroll
      for (int r = 0; r < RA; ++r)
#pragma unroll
        for (int c = 0; c < RB; ++c) dmma(acc[r][c][0], acc[r][c][1], af[r], bf[c]);
    }
  }
  double t = 0;
  for (int r = 0; r < RA; ++r) for (int c = 0; c < RB; ++c) t += acc[r][c][0] + acc[r][c][1];
  if (t == 123.456) out[0] = t;
}
template <typename F> double timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaDeviceSynchronize(); cudaEventRecord(a);
  for (int i = 0; i < 10; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms / 10;
}
int main() {
  double* out; CK(cudaMalloc(&out, 64));
  const double fl = 148.0 * 12 * ITERS * 512.0;  // per MMA-per-warp unit
  for (int threads : {128, 256, 384}) {
    double w = threads / 32.0 / 12.0;
    double ms = timeit([&] { reg_tile<3, 8><<<148, threads>>>(out, 1e-3); });
    printf("reg_tile<3,8> %d thr: %.2f TF\n", threads, fl * w * 24 / ms * 1e-9);
    ms = timeit([&] { reg_tile<4, 8><<<148, threads>>>(out, 1e-3); });
    printf("reg_tile<4,8> %d thr: %.2f TF\n", threads, fl * w * 32 / ms * 1e-9);
    ms = timeit([&] { reg_tile<3, 6><<<148, threads>>>(out, 1e-3); });
    printf("reg_tile<3,6> %d thr: %.2f TF\n", threads, fl * w * 18 / ms * 1e-9);
  }
  double ms = timeit([&] { smem_tile<3, 8, 4><<<148, 384>>>(out, 1e-3); });
  printf("smem_tile<3,8> 384 thr: %.2f TF\n", fl * 24 / ms * 1e-9);
  ms = timeit([&] { smem_tile<3, 8, 4><<<148, 256>>>(out, 1e-3); });
  printf("smem_tile<3,8> 256 thr: %.2f TF\n", fl * (8.0 / 12) * 24 / ms * 1e-9);
  CK(cudaDeviceSynchronize());
  return 0;
}
