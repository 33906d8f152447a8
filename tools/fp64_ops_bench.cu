// Development micro-benchmark: issue rate of DFMA / DMUL / DADD / I2F.F64 on one B200 SM sub-partition (cycles per warp instruction)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/fp64_ops_bench tools/fp64_ops_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
template <int OP>
__global__ void k(double* out, long long* cyc, double a, double b, int iters) {
  double x[8];
  for (int i = 0; i < 8; ++i) x[i] = a + i + threadIdx.x;
  int xi[8];
  for (int i = 0; i < 8; ++i) xi[i] = (int)threadIdx.x + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) x[i] = fma(x[i], a, b);
      if (OP == 1) x[i] = x[i] * a;
      if (OP == 2) x[i] = x[i] + b;
      if (OP == 3) { x[i] += (double)xi[i]; xi[i] += it; }
    }
  }
  const long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 8; ++i) s += x[i] + xi[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
  const char* names[4] = {"DFMA", "DMUL", "DADD", "I2F.F64 + DADD"};
  for (int warps : {4, 8, 16}) {
    for (int op = 0; op < 4; ++op) {
      const int iters = 4000;
      if (op == 0) k<0><<<148, warps * 32>>>(out, cyc, 1.0000001, 1e-9, iters);
      if (op == 1) k<1><<<148, warps * 32>>>(out, cyc, 1.0000001, 1e-9, iters);
      if (op == 2) k<2><<<148, warps * 32>>>(out, cyc, 1.0000001, 1e-9, iters);
      if (op == 3) k<3><<<148, warps * 32>>>(out, cyc, 1.0000001, 1e-9, iters);
      cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      // warp instructions per SM sub-partition = warps / 4 * 8 * iters
      printf("%-16s %2d warps/SM: %.2f cycles per warp instruction per sub-partition\n", names[op], warps, (double)c / (warps / 4.0 * 8 * iters));
    }
  }
  return 0;
}
