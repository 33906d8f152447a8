// FP64 peak microbenchmark for B200 (sm_100a): DFMA vector pipe vs DMMA (mma.sync f64) shapes.
// The SVGP hot path is FP64-compute-bound, and MEASURED_PEAKS.json has no FP64
// entry, so this program measures the roofline denominator.  Prints one JSON line.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

template <int NCHAIN>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, double a, double b) {
  double acc[NCHAIN];
#pragma unroll
  for (int i = 0; i < NCHAIN; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NCHAIN; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;  // never true; keeps the chain alive
}

// m8n8k4: A 8x4 (1 reg), B 4x8 (1 reg), C 8x8 (2 regs) per thread
template <int NACC>
__global__ void __launch_bounds__(256) dmma884_kernel(double* out, double a, double b) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c0[i] = i; c1[i] = -i; }
  double fa = a + threadIdx.x * 1e-12, fb = b;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(fa), "d"(fb));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;
}

// m16n8k4: A 16x4 (2 regs), B 4x8 (1 reg), C 16x8 (4 regs)
template <int NACC>
__global__ void __launch_bounds__(256) dmma1684_kernel(double* out, double a, double b) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  double fa0 = a + threadIdx.x * 1e-12, fa1 = a, fb = b;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(fa0), "d"(fa1), "d"(fb));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[0] = s;
}

// m16n8k8: A 16x8 (4 regs), B 8x8 (2 regs), C 16x8 (4 regs)
template <int NACC>
__global__ void __launch_bounds__(256) dmma1688_kernel(double* out, double a, double b) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  double fa0 = a + threadIdx.x * 1e-12, fa1 = a, fa2 = a * 0.5, fa3 = a * 0.25, fb0 = b, fb1 = b * 0.5;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(fa0), "d"(fa1), "d"(fa2), "d"(fa3), "d"(fb0), "d"(fb1));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[0] = s;
}

// m16n8k16: A 16x16 (8 regs), B 16x8 (4 regs), C 16x8 (4 regs)
template <int NACC>
__global__ void __launch_bounds__(256) dmma16816_kernel(double* out, double a, double b) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  double fa[8], fb[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) fa[i] = a * (i + 1) + threadIdx.x * 1e-12;
#pragma unroll
  for (int i = 0; i < 4; ++i) fb[i] = b * (i + 1);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
                   "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(fa[0]), "d"(fa[1]), "d"(fa[2]), "d"(fa[3]), "d"(fa[4]), "d"(fa[5]), "d"(fa[6]), "d"(fa[7]),
                     "d"(fb[0]), "d"(fb[1]), "d"(fb[2]), "d"(fb[3]));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[0] = s;
}

// DFMA and DMMA interleaved: do the two pipes overlap?
__global__ void __launch_bounds__(256) mixed_kernel(double* out, double a, double b) {
  double acc[8];
  double c0[4], c1[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-9 + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) { c0[i] = i; c1[i] = -i; }
  double fa = a + threadIdx.x * 1e-12, fb = b;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(fa), "d"(fb));
      acc[2 * i] = fma(acc[2 * i], a, b);
      acc[2 * i + 1] = fma(acc[2 * i + 1], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) launch();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, 64));
  const int reps = 20;
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
  // CTAs per SM sweep: 2, 4, 8 blocks of 256 threads
  for (int bps : {1, 2, 4, 8}) {
    int grid = sms * bps;
    double thr = (double)grid * 256;
    {
      double ms = time_ms([&] { dfma_kernel<16><<<grid, 256>>>(out, 1.0000001, 1e-9); }, reps);
      printf(", \"dfma16_bps%d_tflops\": %.3f", bps, thr * 16 * ITERS * 2 / ms * 1e-9);
    }
    {
      double ms = time_ms([&] { dmma884_kernel<8><<<grid, 256>>>(out, 1.0000001, 1e-9); }, reps);
      // per warp-instr: 8*8*4 FMA = 512 flop
      printf(", \"dmma884_bps%d_tflops\": %.3f", bps, (thr / 32) * 8 * ITERS * 512.0 / ms * 1e-9);
    }
    {
      double ms = time_ms([&] { dmma1684_kernel<4><<<grid, 256>>>(out, 1.0000001, 1e-9); }, reps);
      printf(", \"dmma1684_bps%d_tflops\": %.3f", bps, (thr / 32) * 4 * ITERS * 1024.0 / ms * 1e-9);
    }
    {
      double ms = time_ms([&] { dmma1688_kernel<4><<<grid, 256>>>(out, 1.0000001, 1e-9); }, reps);
      printf(", \"dmma1688_bps%d_tflops\": %.3f", bps, (thr / 32) * 4 * ITERS * 2048.0 / ms * 1e-9);
    }
    {
      double ms = time_ms([&] { dmma16816_kernel<4><<<grid, 256>>>(out, 1.0000001, 1e-9); }, reps);
      printf(", \"dmma16816_bps%d_tflops\": %.3f", bps, (thr / 32) * 4 * ITERS * 4096.0 / ms * 1e-9);
    }
    {
      double ms = time_ms([&] { mixed_kernel<<<grid, 256>>>(out, 1.0000001, 1e-9); }, reps);
      double flop = (thr / 32) * 4 * ITERS * 512.0 + thr * 8 * ITERS * 2;
      printf(", \"mixed_bps%d_tflops\": %.3f", bps, flop / ms * 1e-9);
    }
  }
  // sustained DFMA for ~3 s to see the clock under the power cap
  {
    int grid = sms * 4;
    double thr = (double)grid * 256;
    double ms = time_ms([&] { dfma_kernel<16><<<grid, 256>>>(out, 1.0000001, 1e-9); }, 3000);
    printf(", \"dfma16_sustained_tflops\": %.3f", thr * 16 * ITERS * 2 / ms * 1e-9);
  }
  printf("}\n");
  return 0;
}
