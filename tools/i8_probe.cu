// Development probe (not part of the library): tcgen05.mma kind::i8 (s8 x s8 -> s32 in TMEM) with both operands K-major in the
// un-swizzled canonical layout (8 rows x 16 bytes core matrices), to pin the descriptors of the Ozaki-split f64 GEMM
// (csrc/gemm_i8.cuh) and to measure how fast one thread can issue small-N MMAs:
//   part 1: D_g[128 x N] = sum over pairs (i, j), i + j = g, of A_i[128 x 32] B_j[N x 32]^T  against a CPU product
//   part 2: cycles per MMA for N = 32 .. 256, the 21-product sequence of a 6-slice split issued back to back
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/i8_probe tools/i8_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int S = 6;        // slices per operand
constexpr int M = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr, uint32_t lbo) { return ((addr >> 4) & 0x3FFF) | (((lbo >> 4) & 0x3FFF) << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo) { return ((sbo >> 4) & 0x3FFF) | (1u << 14); }   // version 1, layout 0 (no swizzle)
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\nsetp.ne.b32 p, %6, 0;\n"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %5, p;\n}\n"
               ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                 "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                 "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// smem image of one slice of an operand tile [R rows x 32 k]: [k16 chunk (2)][row group (R / 8)][8 rows][16 bytes]
//   -> LBO (k direction) = R * 16 bytes, SBO (row-group direction) = 128 bytes
// Aimg: S slices x 4096 B, Bimg: S slices x N * 32 B (host-prepared in exactly this order)
template <int NISSUE>
__global__ void __launch_bounds__(192, 1) probe(const int8_t* Aimg, const int8_t* Bimg, int N, int G, int reps, int32_t* D, long long* cycles,
                                                int swap_lbo_sbo) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* sA = smem;
  uint8_t* sB = smem + S * 4096;
  for (int i = tid; i < S * 4096 / 16; i += blockDim.x) ((int4*)sA)[i] = ((const int4*)Aimg)[i];
  for (int i = tid; i < S * N * 32 / 16; i += blockDim.x) ((int4*)sB)[i] = ((const int4*)Bimg)[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  // D s32 (2 << 4), A s8 (1 << 7), B s8 (1 << 10), both K-major, N >> 3 << 17, M >> 4 << 24
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  const uint32_t a_lbo = swap_lbo_sbo ? 128 : M * 16, a_sbo = swap_lbo_sbo ? M * 16 : 128;
  const uint32_t b_lbo = swap_lbo_sbo ? 128 : N * 16, b_sbo = swap_lbo_sbo ? N * 16 : 128;
  const uint32_t hiA = desc_hi(a_sbo), hiB = desc_hi(b_sbo);
  const uint32_t a0 = desc_lo(smem_u32(sA), a_lbo), b0 = desc_lo(smem_u32(sB), b_lbo);
  const uint32_t bstep = (uint32_t)(N * 32) >> 4;   // slice stride of B in descriptor units

  // issuing warps 0 .. NISSUE-1: issuer w takes the groups g with g % NISSUE == w (distinct accumulators per issuer)
  if (warp < NISSUE && lane == 0) {
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int i = 0; i < S; ++i)
#pragma unroll
        for (int j = 0; j < S; ++j) {
          if (i + j > S - 1) continue;                      // i + j + 2 <= S + 1
          const int g = (i + j) % G;                        // accumulator (group) index
          if (g % NISSUE != warp) continue;
          // first product of a group in the first repetition overwrites: the pair with i = 0 (j = g) when G == S
          const uint32_t acc = (r == 0 && i == 0 && (i + j) < G) ? 0u : 1u;
          mma_i8(tmem + (uint32_t)(g * N), a0 + 256u * i, hiA, b0 + bstep * j, hiB, idesc, acc);
        }
    }
    commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
    const long long t1 = clock64();
    cycles[blockIdx.x * 4 + warp] = t1 - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp < 4 && D && blockIdx.x == 0) {
    for (int c0 = 0; c0 < G * N; c0 += 32) {
      uint32_t v[32];
      ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      for (int j = 0; j < 32; ++j) D[(size_t)(warp * 32 + lane) * (G * N) + c0 + j] = (int32_t)v[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// ---- TS variant: the A planes are copied once per k-block from shared memory into TMEM (tcgen05.cp 128x256b) and the MMAs read A
// from there, so an MMA's shared-memory traffic is B only (32 N bytes instead of 4096 + 32 N) ----
__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint32_t lo, uint32_t hi) {
  asm volatile("{\n.reg .b64 d;\nmov.b64 d, {%1, %2};\ntcgen05.cp.cta_group::1.128x256b [%0], d;\n}\n" ::"r"(taddr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ void mma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\n.reg .b64 db;\nmov.b64 db, {%2, %3};\nsetp.ne.b32 p, %5, 0;\n"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], db, %4, p;\n}\n"
               ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\nselp.b32 %0, 1, 0, px;\n}\n" : "=r"(pred));
  return pred != 0;
}
__global__ void __launch_bounds__(192, 1) probe_ts(const int8_t* Aimg, const int8_t* Bimg, int N, int reps, int32_t* D, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  uint8_t* sA = smem;
  uint8_t* sB = smem + S * 4096;
  for (int i = tid; i < S * 4096 / 16; i += blockDim.x) ((int4*)sA)[i] = ((const int4*)Aimg)[i];
  for (int i = tid; i < S * N * 32 / 16; i += blockDim.x) ((int4*)sB)[i] = ((const int4*)Bimg)[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_base_s, 0);
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  const uint32_t hiA = desc_hi(128), hiB = desc_hi(128);
  const uint32_t a0 = desc_lo(smem_u32(sA), M * 16), b0 = desc_lo(smem_u32(sB), N * 16);
  const uint32_t bstep = (uint32_t)(N * 32) >> 4;
  const uint32_t ta = tmem + (uint32_t)(S * N);        // A planes: 8 columns each, two buffers of S planes
  if (warp == 0) {
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (elect_one()) {
        const uint32_t tab = ta + (uint32_t)((r & 1) * S * 8);
#pragma unroll
        for (int i = 0; i < S; ++i) cp_128x256b(tab + 8u * i, a0 + 256u * i, hiA);
#pragma unroll
        for (int i = 0; i < S; ++i)
#pragma unroll
          for (int j = 0; j < S - i; ++j)
            mma_i8_ts(tmem + (uint32_t)((i + j) * N), tab + 8u * i, b0 + bstep * j, hiB, idesc, (r | i) ? 1u : 0u);
      }
      __syncwarp();
    }
    if (elect_one()) commit(&bar[0]);
    __syncwarp();
    mbar_wait(&bar[0], 0);
    if (lane == 0) cycles[blockIdx.x * 4] = clock64() - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp < 4 && D && blockIdx.x == 0) {
    for (int c0 = 0; c0 < S * N; c0 += 32) {
      uint32_t v[32];
      ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      for (int j = 0; j < 32; ++j) D[(size_t)(warp * 32 + lane) * (S * N) + c0 + j] = (int32_t)v[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

static void make_image(const std::vector<int8_t>& T, int R, std::vector<int8_t>& img) {   // T[slice][row][32] -> image
  img.assign((size_t)S * R * 32, 0);
  for (int s = 0; s < S; ++s)
    for (int r = 0; r < R; ++r)
      for (int k = 0; k < 32; ++k)
        img[(size_t)s * R * 32 + (size_t)(k / 16) * R * 16 + (size_t)(r / 8) * 128 + (r % 8) * 16 + (k % 16)] = T[((size_t)s * R + r) * 32 + k];
}

int main() {
  srand(7);
  int32_t* dD; long long* dC;
  CK(cudaMalloc(&dD, 128 * 512 * 4));
  CK(cudaMalloc(&dC, 148 * 4 * 8));
  for (int N : {64, 80}) {
    const int G = S;
    std::vector<int8_t> A((size_t)S * M * 32), B((size_t)S * N * 32), Ai, Bi;
    for (auto& v : A) v = (int8_t)(rand() % 256 - 128);
    for (auto& v : B) v = (int8_t)(rand() % 256 - 128);
    make_image(A, M, Ai);
    make_image(B, N, Bi);
    int8_t *dA, *dB;
    CK(cudaMalloc(&dA, Ai.size())); CK(cudaMalloc(&dB, Bi.size()));
    CK(cudaMemcpy(dA, Ai.data(), Ai.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bi.data(), Bi.size(), cudaMemcpyHostToDevice));
    const int smem = S * 4096 + S * N * 32 + 1024;
    CK(cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int swap = 0; swap < 1; ++swap) {   // swapped LBO / SBO reads outside the tile (illegal address): not a candidate
      CK(cudaMemset(dD, 0, 128 * 512 * 4));
      probe<1><<<1, 192, smem>>>(dA, dB, N, G, 1, dD, dC, swap);
      CK(cudaDeviceSynchronize());
      std::vector<int32_t> D((size_t)128 * G * N);
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      long bad = 0;
      for (int g = 0; g < G; ++g)
        for (int m = 0; m < M; ++m)
          for (int n = 0; n < N; ++n) {
            int32_t ref = 0;
            for (int i = 0; i <= g; ++i) {
              const int j = g - i;
              for (int k = 0; k < 32; ++k) ref += (int32_t)A[((size_t)i * M + m) * 32 + k] * (int32_t)B[((size_t)j * N + n) * 32 + k];
            }
            if (ref != D[(size_t)m * (G * N) + g * N + n]) ++bad;
          }
      printf("N = %d  lbo/sbo %s: %ld of %d mismatches\n", N, swap ? "swapped" : "k-stride in LBO", bad, G * M * N);
    }
    CK(cudaFree(dA)); CK(cudaFree(dB));
  }
  // ---- TS variant: correctness (N = 64, one k-block) and rate ----
  {
    const int N = 64;
    std::vector<int8_t> A((size_t)S * M * 32), B((size_t)S * N * 32), Ai, Bi;
    for (auto& v : A) v = (int8_t)(rand() % 256 - 128);
    for (auto& v : B) v = (int8_t)(rand() % 256 - 128);
    make_image(A, M, Ai);
    make_image(B, N, Bi);
    int8_t *dA, *dB;
    CK(cudaMalloc(&dA, Ai.size())); CK(cudaMalloc(&dB, Bi.size()));
    CK(cudaMemcpy(dA, Ai.data(), Ai.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bi.data(), Bi.size(), cudaMemcpyHostToDevice));
    const int smem = S * 4096 + S * N * 32 + 1024;
    CK(cudaFuncSetAttribute(probe_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaMemset(dD, 0, 128 * 512 * 4));
    probe_ts<<<1, 192, smem>>>(dA, dB, N, 1, dD, dC);
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> D((size_t)128 * S * N);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int g = 0; g < S; ++g)
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
          int32_t ref = 0;
          for (int i = 0; i <= g; ++i)
            for (int k = 0; k < 32; ++k) ref += (int32_t)A[((size_t)i * M + m) * 32 + k] * (int32_t)B[((size_t)(g - i) * N + n) * 32 + k];
          if (ref != D[(size_t)m * (S * N) + g * N + n]) ++bad;
        }
    printf("TS (A planes via tcgen05.cp 128x256b into TMEM) N = %d: %ld of %d mismatches\n", N, bad, S * M * N);
    for (int grid : {1, 148}) {
      const int reps = 2000;
      probe_ts<<<grid, 192, smem>>>(dA, dB, N, reps, nullptr, dC);
      CK(cudaDeviceSynchronize());
      long long c[148 * 4];
      CK(cudaMemcpy(c, dC, sizeof(long long) * grid * 4, cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (int b = 0; b < grid; ++b) mx = c[b * 4] > mx ? c[b * 4] : mx;
      printf("TS N = %d grid %3d: %.1f cycles per MMA incl. 6 copies per 21 MMAs (math floor %d), %.0f MAC/clk/SM\n", N, grid, (double)mx / (21.0 * reps), N / 2,
             21.0 * reps * 128 * N * 32 / mx);
    }
    CK(cudaFree(dA)); CK(cudaFree(dB));
  }
  // ---- issue-rate timing: all SMs, operands static in shared memory ----
  for (int N : {64, 80}) {
    const int G = 512 / N < S ? 512 / N : S;
    std::vector<int8_t> Ai((size_t)S * 4096, 1), Bi((size_t)S * N * 32, 1);
    int8_t *dA, *dB;
    CK(cudaMalloc(&dA, Ai.size())); CK(cudaMalloc(&dB, Bi.size()));
    CK(cudaMemcpy(dA, Ai.data(), Ai.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bi.data(), Bi.size(), cudaMemcpyHostToDevice));
    const int smem = S * 4096 + S * N * 32 + 1024;
    const int reps = 2000;
    const int nmma = 21 * reps;
    auto run = [&](auto kern, int nissue) {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      for (int grid : {1, 148}) {
        kern<<<grid, 192, smem>>>(dA, dB, N, G, reps, nullptr, dC, 0);
        CK(cudaDeviceSynchronize());
        long long c[148 * 4];
        CK(cudaMemcpy(c, dC, sizeof(long long) * grid * 4, cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (int b = 0; b < grid; ++b) for (int w = 0; w < nissue; ++w) mx = c[b * 4 + w] > mx ? c[b * 4 + w] : mx;
        printf("N = %3d  issuers %d  grid %3d: %.1f cycles per MMA (floor %d), %.0f MAC/clk/SM\n", N, nissue, grid, (double)mx / nmma, N / 2,
               (double)nmma * 128 * N * 32 / mx);
      }
    };
    run(probe<1>, 1);
    run(probe<2>, 2);
    run(probe<3>, 3);
    CK(cudaFree(dA)); CK(cudaFree(dB));
  }
  return 0;
}
