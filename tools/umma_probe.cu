// Development probe (not part of the library): one CTA computes D[128 x 288] = A[128 x 288] * B[288 x 288]^T in 3 x TF32
// on tcgen05 (TMEM accumulator, TMA tensor loads with 128 B swizzle), to pin down the shared-memory / instruction
// descriptors the f32 path of gpfy_b200 uses before they go into the real kernels:
//   variant 0: A K-major  (A given as [128 points][288 k], k contiguous),  B K-major ([288 n][288 k])
//   variant 1: A MN-major (A given as [288 k][128 points], points contiguous: the feature-major Z panel), B K-major
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_probe tools/umma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int M = 128, N = 288, NH = 144, K = 288, BK = 32, KB = K / BK;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46 | layout << 61
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t layout = 2 /* SWIZZLE_128B */) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
         (1ull << 46) | (layout << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct Maps { CUtensorMap a_hi, a_lo, b_hi, b_lo; };

// A_MN: the MN-major tf32 operand uses the SWIZZLE_128B_BASE32B layout (type 1; 32-byte swizzle chunks, atoms of 4 k-rows x
// 128 B; TMA mode CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) - "for mn-major tf32 operands, SW128_32B is the only available smem
// layout" (cutlass sm100_common.inl).  lbo / sbo are passed at run time to try the candidates in one run.
template <int A_MN>
__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ Maps maps, float* D, uint32_t a_lbo, uint32_t a_sbo) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* a_hi = (float*)smem;                       // [128 x 32] = 16 KB
  float* a_lo = (float*)(smem + 16384);
  float* b_hi = (float*)(smem + 32768);             // [288 x 32] = 36 KB
  float* b_lo = (float*)(smem + 32768 + 36864);
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  // instruction descriptor: D f32 (1 << 4), A tf32 (2 << 7), B tf32 (2 << 10), a_major bit 15, b_major bit 16, N >> 3 << 17, M >> 4 << 24
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)A_MN << 15) | (0u << 16) | ((uint32_t)(NH >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

  if (tid == 0) {
    for (int kb = 0; kb < KB; ++kb) {
      mbar_expect_tx(&bar_full, (2 * 16384 + 2 * 36864));
      if (A_MN) {
        for (int c = 0; c < 4; ++c) {   // 4 chunks of 32 points: box = 32 points x 32 k rows
          tma_load_2d((uint8_t*)a_hi + c * 4096, &maps.a_hi, c * 32, kb * BK, &bar_full);
          tma_load_2d((uint8_t*)a_lo + c * 4096, &maps.a_lo, c * 32, kb * BK, &bar_full);
        }
      } else {
        tma_load_2d(a_hi, &maps.a_hi, kb * BK, 0, &bar_full);   // box = 32 k x 128 rows
        tma_load_2d(a_lo, &maps.a_lo, kb * BK, 0, &bar_full);
      }
      for (int h = 0; h < 2; ++h) {     // B rows in two boxes of 144
        tma_load_2d((uint8_t*)b_hi + h * 144 * 128, &maps.b_hi, kb * BK, h * 144, &bar_full);
        tma_load_2d((uint8_t*)b_lo + h * 144 * 128, &maps.b_lo, kb * BK, h * 144, &bar_full);
      }
      mbar_wait(&bar_full, kb & 1);
      asm volatile("tcgen05.fence::after_thread_sync;");
      for (int ks = 0; ks < BK / 8; ++ks) {
        for (int h = 0; h < 2; ++h) {
          const uint32_t d = tmem + h * NH;
          uint64_t dah, dal;
          if (A_MN) {
            dah = smem_desc(smem_u32(a_hi) + ks * 1024, a_lbo, a_sbo, 1);
            dal = smem_desc(smem_u32(a_lo) + ks * 1024, a_lbo, a_sbo, 1);
          } else {
            dah = smem_desc(smem_u32(a_hi) + ks * 32, 16, 1024);
            dal = smem_desc(smem_u32(a_lo) + ks * 32, 16, 1024);
          }
          const uint64_t dbh = smem_desc(smem_u32(b_hi) + h * 144 * 128 + ks * 32, 16, 1024);
          const uint64_t dbl = smem_desc(smem_u32(b_lo) + h * 144 * 128 + ks * 32, 16, 1024);
          const uint32_t first = (kb == 0 && ks == 0) ? 0u : 1u;
          umma_tf32(d, dal, dbh, idesc, first);   // small terms first
          umma_tf32(d, dah, dbl, idesc, 1u);
          umma_tf32(d, dah, dbh, idesc, 1u);
        }
      }
      umma_commit(&bar_mma);
      mbar_wait(&bar_mma, kb & 1);
    }
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  // epilogue: warp w reads TMEM lanes 32 w .. 32 w + 31 (= rows of D), 32 columns at a time
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// variant 4: both operands K-major with the 64-byte swizzle (layout type 4, rows of 16 tf32 = 64 B, 8-row atoms of 512 B):
// k-blocks of 16 instead of 32, i.e. half the bytes per pipeline stage (the weighted-Gram kernel wants 4 stages).
__global__ void __launch_bounds__(128, 1) probe_sw64_kernel(const __grid_constant__ Maps maps, float* D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_hi = smem;                  // [128 x 16] = 8 KB
  uint8_t* a_lo = smem + 8192;
  uint8_t* b_hi = smem + 16384;          // [288 x 16] = 18 KB
  uint8_t* b_lo = smem + 16384 + 18432;
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&bar_full, 1); mbar_init(&bar_mma, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NH >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  if (tid == 0) {
    for (int kb = 0; kb < K / 16; ++kb) {
      mbar_expect_tx(&bar_full, 2 * 8192 + 2 * 18432);
      tma_load_2d(a_hi, &maps.a_hi, kb * 16, 0, &bar_full);
      tma_load_2d(a_lo, &maps.a_lo, kb * 16, 0, &bar_full);
      for (int h = 0; h < 2; ++h) {
        tma_load_2d(b_hi + h * 144 * 64, &maps.b_hi, kb * 16, h * 144, &bar_full);
        tma_load_2d(b_lo + h * 144 * 64, &maps.b_lo, kb * 16, h * 144, &bar_full);
      }
      mbar_wait(&bar_full, kb & 1);
      asm volatile("tcgen05.fence::after_thread_sync;");
      for (int ks = 0; ks < 2; ++ks)
        for (int h = 0; h < 2; ++h) {
          const uint32_t d = tmem + h * NH;
          const uint64_t dah = smem_desc(smem_u32(a_hi) + ks * 32, 16, 512, 4), dal = smem_desc(smem_u32(a_lo) + ks * 32, 16, 512, 4);
          const uint64_t dbh = smem_desc(smem_u32(b_hi) + h * 144 * 64 + ks * 32, 16, 512, 4);
          const uint64_t dbl = smem_desc(smem_u32(b_lo) + h * 144 * 64 + ks * 32, 16, 512, 4);
          umma_tf32(d, dal, dbh, idesc, (kb == 0 && ks == 0) ? 0u : 1u);
          umma_tf32(d, dah, dbl, idesc, 1u);
          umma_tf32(d, dah, dbh, idesc, 1u);
        }
      umma_commit(&bar_mma);
      mbar_wait(&bar_mma, kb & 1);
    }
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn enc, float* base, uint64_t inner, uint64_t outer, uint32_t box_inner, uint32_t box_outer,
                            CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  CUtensorMap m;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner * sizeof(float)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
  return m;
}

static void split(const std::vector<double>& x, std::vector<float>& hi, std::vector<float>& lo) {
  hi.resize(x.size()); lo.resize(x.size());
  for (size_t i = 0; i < x.size(); ++i) {
    float f = (float)x[i];
    uint32_t u; memcpy(&u, &f, 4); u &= 0xFFFFE000u;   // keep sign, exponent and 10 mantissa bits: exact as tf32
    float h; memcpy(&h, &u, 4);
    float l = (float)(x[i] - (double)h);
    uint32_t ul; memcpy(&ul, &l, 4); ul &= 0xFFFFE000u; memcpy(&l, &ul, 4);
    hi[i] = h; lo[i] = l;
  }
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fn;
  srand(1);
  std::vector<double> A(M * K), B(N * K);    // A[m][k], B[n][k]
  for (auto& v : A) v = (rand() / (double)RAND_MAX - 0.5) * 4.0;
  for (auto& v : B) v = (rand() / (double)RAND_MAX - 0.5) * 0.1;
  std::vector<double> ref(M * N, 0.0);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k]; ref[m * N + n] = s; }
  std::vector<float> Bh, Bl; split(B, Bh, Bl);
  float *dBh, *dBl, *dAh, *dAl, *dD;
  CK(cudaMalloc(&dBh, N * K * 4)); CK(cudaMalloc(&dBl, N * K * 4)); CK(cudaMalloc(&dAh, M * K * 4)); CK(cudaMalloc(&dAl, M * K * 4)); CK(cudaMalloc(&dD, M * N * 4));
  CK(cudaMemcpy(dBh, Bh.data(), N * K * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dBl, Bl.data(), N * K * 4, cudaMemcpyHostToDevice));
  const int smem = 32768 + 2 * 36864 + 1024;
  const uint32_t cand[5][2] = {{0, 0}, {4096, 512}, {512, 4096}, {4096, 1024}, {16, 512}};   // (LBO, SBO) candidates of the MN-major operand
  for (int variant = 0; variant < 5; ++variant) {
    std::vector<double> Ax(M * K);
    if (variant == 0 || variant == 4) Ax = A;
    else for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) Ax[k * M + m] = A[m * K + k];   // [k][m]: MN-major
    std::vector<float> Ah, Al; split(Ax, Ah, Al);
    CK(cudaMemcpy(dAh, Ah.data(), M * K * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dAl, Al.data(), M * K * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, M * N * 4));
    Maps maps;
    if (variant == 0) { maps.a_hi = make_map(enc, dAh, K, M, 32, 128); maps.a_lo = make_map(enc, dAl, K, M, 32, 128); }
    else if (variant == 4) { maps.a_hi = make_map(enc, dAh, K, M, 16, 128, CU_TENSOR_MAP_SWIZZLE_64B); maps.a_lo = make_map(enc, dAl, K, M, 16, 128, CU_TENSOR_MAP_SWIZZLE_64B); }
    else { maps.a_hi = make_map(enc, dAh, M, K, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B); maps.a_lo = make_map(enc, dAl, M, K, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B); }
    maps.b_hi = make_map(enc, dBh, K, N, 32, 144); maps.b_lo = make_map(enc, dBl, K, N, 32, 144);
    if (variant == 4) {
      maps.b_hi = make_map(enc, dBh, K, N, 16, 144, CU_TENSOR_MAP_SWIZZLE_64B); maps.b_lo = make_map(enc, dBl, K, N, 16, 144, CU_TENSOR_MAP_SWIZZLE_64B);
      CK(cudaFuncSetAttribute(probe_sw64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      probe_sw64_kernel<<<1, 128, smem>>>(maps, dD);
    } else if (variant == 0) {
      CK(cudaFuncSetAttribute(probe_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      probe_kernel<0><<<1, 128, smem>>>(maps, dD, 0, 0);
    } else {
      CK(cudaFuncSetAttribute(probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      probe_kernel<1><<<1, 128, smem>>>(maps, dD, cand[variant][0], cand[variant][1]);
    }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> D(M * N);
    CK(cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0; int bad = 0;
    for (int i = 0; i < M * N; ++i) { double e = fabs(D[i] - ref[i]); if (e > maxerr) maxerr = e; if (fabs(ref[i]) > maxref) maxref = fabs(ref[i]); if (e > 1e-4) ++bad; }
    printf("variant %d (%s, LBO %u SBO %u): max abs err %.3e, max |ref| %.3e, rel %.3e, entries off by > 1e-4: %d  D[0..3] = %g %g %g %g  ref = %g %g %g %g\n", variant,
           variant == 4 ? "A, B K-major 64 B swizzle" : (variant ? "A MN-major" : "A K-major"), cand[variant][0], cand[variant][1], maxerr, maxref, maxerr / maxref, bad, D[0], D[1], D[2], D[3], ref[0], ref[1], ref[2], ref[3]);
    fflush(stdout);
  }
  return 0;
}
