// K_G32  the dense contraction of the SVGP step on the 5th-generation tensor cores (tcgen05, TMEM accumulators, TMA
// tensor loads) in 3 x TF32 "split" arithmetic - the f32 instantiation of  C = W2 Z  (variational.py:323-328 folded,
// see DESIGN.md §2).
//
// Every operand is stored as two float planes hi + lo with hi = the value truncated to TF32 (10 mantissa bits) and
// lo = tf32(value - hi): the pair carries 22 mantissa bits, and
//     C ~= Ahi Bhi + Ahi Blo + Alo Bhi        (three kind::tf32 MMAs into one fp32 TMEM accumulator)
// has a relative error of ~2^-21 per product, i.e. fp32-level accuracy at tensor-core speed.
//
// Shape: per CTA tile D[128 points x Mp features] = Zt[128 x Mp] W2^T.
//   A = the tile's 128 points of the feature-major panels Zh / Zl [Mp][ld] (points contiguous = "MN-major" operand):
//       per 32-feature k-block four TMA boxes of 32 points x 32 features per plane.  An MN-major tf32 operand has exactly
//       one legal shared-memory layout, SWIZZLE_128B_BASE32B (32-byte swizzle chunks, atoms of 4 k-rows x 128 B; TMA mode
//       CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): LBO = 4096 B between the 32-point chunks, SBO = 512 B between 4-row groups
//       (tools/umma_probe.cu pins these against a CPU product: profiles/umma_probe_r02.log)
//   B = W2h / W2l [Mp][Mp] row-major (K-major operand), loaded in two halves of Mp / 2 rows per k-block
//   D in TMEM: lane = point, column = feature (Mp <= 512 columns), two MMAs of N = Mp / 2 per k-step
// Roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warps 2..5 = epilogue
// (tcgen05.ld -> psi_n = z_n . C_n in fp64 -> C widened to double and stored in the tile-contiguous swizzled layout the
// backward kernels zonal_bwd3 / zonal_bwd4 read, so everything downstream of the contraction is shared with the f64 path).
// Two independent rings feed the MMA warp: A (2 stages) and B halves (NSB stages), so Mp up to 384 fits in 227 KB.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace gpfy {

constexpr int G32_THREADS = 192, G32_BM = 128, G32_BK = 32, G32_NSA = 2;
constexpr int G32_A_STAGE = 2 * 4 * 4096;   // two planes x four 32-point chunks x (32 features x 128 B)

struct Gemm32Maps {
  CUtensorMap zh, zl;   // tile-contiguous planes [n / 32][Mp][32] viewed as [tiles * Mp rows][32 cols], box 32 points x 32 features
  CUtensorMap wh, wl;   // [Mp rows][Mp cols] float, box 32 (k) x Mp / 2 rows, 128 B swizzle
};

struct Gemm32Args {
  const float* Zh;      // for the psi epilogue (the same planes the TMA maps describe): element (row, n) at
  const float* Zl;      //   ((n >> 5) * Mp + row) * 32 + (n & 31)   - every 32-point block of all Mp rows is contiguous
  int64_t ldz;
  double* C;            // [tiles32][Mp][32] blocked + swizzled (the layout of GemmArgs::c_blocked), or null (psi only)
  double* psi;          // [ldz] z_n . C_n, or null
  int Mp;
  int64_t npts;         // points of this launch (tiles of 128; the panels are padded to a multiple of 128 columns)
  int nsb;              // B-ring stages
};

__host__ __device__ static inline int g32_b_stage_bytes(int Mp) { return 2 * (Mp / 2) * 128; }   // two planes x Mp / 2 rows x 128 B
static inline int g32_nsb(int Mp) {
  const int avail = 226 * 1024 - 2048 - G32_NSA * G32_A_STAGE;
  int n = avail / g32_b_stage_bytes(Mp);
  return n > 6 ? 6 : n;
}
static inline int g32_smem_bytes(int Mp) { return 2048 + G32_NSA * G32_A_STAGE + g32_nsb(Mp) * g32_b_stage_bytes(Mp); }

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t g32_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void g32_tma_2d(void* smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(g32_smem_u32(smem)), "l"((uint64_t)map), "r"(g32_smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// cute::UMMA::SmemDescriptor: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46 | layout type << 61
// (layout 2 = SWIZZLE_128B for the K-major operand, 1 = SWIZZLE_128B_BASE32B for the MN-major tf32 operand)
__device__ __forceinline__ uint64_t g32_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t layout) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
         (1ull << 46) | (layout << 61);
}
__device__ __forceinline__ void g32_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// The single issuing thread is the bottleneck of these kernels unless an MMA costs it only a handful of instructions
// (measured r02h: ~300 cycles of issue per MMA with the descriptors rebuilt each time, 72 cycles of tensor work): the
// descriptor is split into its constant high word and a low word = (address >> 4) | (LBO >> 4) << 16, advanced by adds.
__device__ __forceinline__ uint32_t g32_desc_hi(uint32_t sbo_bytes, uint32_t layout) { return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (layout << 29); }
__device__ __forceinline__ uint32_t g32_desc_lo(uint32_t addr, uint32_t lbo_bytes) { return ((addr >> 4) & 0x3FFF) | (((lbo_bytes >> 4) & 0x3FFF) << 16); }
__device__ __forceinline__ void g32_mma2(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\nsetp.ne.b32 p, %6, 0;\n"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n}\n"
               ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
// One lane of a converged warp (elect.sync).  Issuing the asynchronous instructions under `if (elect)` inside a warp-uniform loop lets
// the compiler keep descriptors in uniform registers; under `if (lane == 0)` it wrapped every tcgen05.mma / TMA instruction in an
// ELECT / R2UR / BRA.U.ANY uniformisation loop (~75 instead of ~50 cycles of issue per MMA, measured on the int8 kernel).
__device__ __forceinline__ bool g32_elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\nselp.b32 %0, 1, 0, px;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void g32_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(g32_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void g32_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                 "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                 "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(G32_THREADS, 1) gemm_tf32x3_kernel(const __grid_constant__ Gemm32Maps maps, const __grid_constant__ Gemm32Args a) {
  extern __shared__ __align__(1024) uint8_t g32_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)g32_smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t a_full[G32_NSA], a_empty[G32_NSA], b_full[6], b_empty[6], d_full, d_empty;
  __shared__ uint32_t tmem_base_s;
  const int Mp = a.Mp, NH = Mp / 2, KB = Mp / G32_BK, NSB = a.nsb;
  const int b_stage = g32_b_stage_bytes(Mp);
  uint8_t* sA = smem;                                   // [NSA][hi: 4 chunks x 4 KB | lo: 4 chunks x 4 KB]
  uint8_t* sB = smem + G32_NSA * G32_A_STAGE;           // [NSB][hi: NH rows x 128 B | lo: NH rows x 128 B]
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler
  const int tiles = (int)((a.npts + G32_BM - 1) / G32_BM);
  const int n_my = ((int)blockIdx.x < tiles) ? (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < G32_NSA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < NSB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    mbar_init(&d_full, 1);
    mbar_init(&d_empty, 4);
  }
  if (warp == 1) {   // TMEM: all 512 columns (one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(g32_smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_base_s, 0);

  if (warp == 0) {
    // ================= TMA producer (warp-uniform loop, one elected lane issues) =================
    {
      int ia = 0, ib = 0;   // running item counters of the two rings
      for (int it = 0; it < n_my; ++it) {
        const int p0 = ((int)blockIdx.x + it * (int)gridDim.x) * G32_BM;
        for (int kb = 0; kb < KB; ++kb) {
          {
            const int st = ia % G32_NSA;
            if (ia >= G32_NSA) mbar_wait(&a_empty[st], (unsigned)(((ia / G32_NSA) - 1) & 1));
            uint8_t* dst = sA + st * G32_A_STAGE;
            if (g32_elect_one()) {
              mbar_expect_tx(&a_full[st], G32_A_STAGE);
#pragma unroll
              for (int c = 0; c < 4; ++c) {   // chunk c = the 32-point block (p0 / 32 + c): 32 consecutive rows of it are 4 KB of HBM
                const int row = ((p0 >> 5) + c) * Mp + kb * G32_BK;
                g32_tma_2d(dst + c * 4096, &maps.zh, 0, row, &a_full[st]);
                g32_tma_2d(dst + 16384 + c * 4096, &maps.zl, 0, row, &a_full[st]);
              }
            }
            __syncwarp();
            ++ia;
          }
          for (int h = 0; h < 2; ++h) {
            const int st = ib % NSB;
            if (ib >= NSB) mbar_wait(&b_empty[st], (unsigned)(((ib / NSB) - 1) & 1));
            uint8_t* dst = sB + st * b_stage;
            if (g32_elect_one()) {
              mbar_expect_tx(&b_full[st], (unsigned)b_stage);
              g32_tma_2d(dst, &maps.wh, kb * G32_BK, h * NH, &b_full[st]);
              g32_tma_2d(dst + NH * 128, &maps.wl, kb * G32_BK, h * NH, &b_full[st]);
            }
            __syncwarp();
            ++ib;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (warp-uniform loop, one elected lane issues) =================
    {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D f32, A/B tf32, A MN-major, B K-major, N = NH, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (0u << 16) | ((uint32_t)(NH >> 3) << 17) | ((uint32_t)(G32_BM >> 4) << 24);
      const uint32_t hiA = g32_desc_hi(512, 1), hiB = g32_desc_hi(1024, 2);
      int ia = 0, ib = 0;
      for (int it = 0; it < n_my; ++it) {
        if (it > 0) mbar_wait(&d_empty, (unsigned)((it - 1) & 1));   // the epilogue has drained the accumulator
        asm volatile("tcgen05.fence::after_thread_sync;");
        for (int kb = 0; kb < KB; ++kb) {
          const int sa = ia % G32_NSA;
          mbar_wait(&a_full[sa], (unsigned)((ia / G32_NSA) & 1));
          const uint32_t ah = g32_smem_u32(sA + sa * G32_A_STAGE), al = ah + 16384;
          for (int h = 0; h < 2; ++h) {
            const int sb = ib % NSB;
            mbar_wait(&b_full[sb], (unsigned)((ib / NSB) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint32_t bh = g32_smem_u32(sB + sb * b_stage), bl = bh + NH * 128;
            const uint32_t d = tmem + (uint32_t)(h * NH);
            // A (MN-major, layout 1): 8 features = one 1 KB group of a chunk (+64 in the address field), chunks 4 KB apart;
            // B (K-major, layout 2): 8 k = 32 B inside the swizzled row (+2)
            const uint32_t ahl = g32_desc_lo(ah, 4096), all = g32_desc_lo(al, 4096), bhl = g32_desc_lo(bh, 16), bll = g32_desc_lo(bl, 16);
            if (g32_elect_one()) {
#pragma unroll
              for (int ks = 0; ks < G32_BK / 8; ++ks) {
                g32_mma2(d, all + 64 * ks, hiA, bhl + 2 * ks, hiB, idesc, (kb | ks) ? 1u : 0u);
                g32_mma2(d, ahl + 64 * ks, hiA, bll + 2 * ks, hiB, idesc, 1u);
                g32_mma2(d, ahl + 64 * ks, hiA, bhl + 2 * ks, hiB, idesc, 1u);
              }
              g32_commit(&b_empty[sb]);      // frees the B stage once the MMAs above have read it
            }
            __syncwarp();
            ++ib;
          }
          if (g32_elect_one()) g32_commit(&a_empty[sa]);
          __syncwarp();
          ++ia;
        }
        if (g32_elect_one()) g32_commit(&d_full);               // accumulator of this tile complete
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue: warps 2..5 own TMEM lanes 32 (warp % 4) .. + 31 =================
    const int q = warp & 3;
    for (int it = 0; it < n_my; ++it) {
      const int64_t p0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * G32_BM + 32 * q;   // this warp's 32 points
      const int64_t n = p0 + lane;
      mbar_wait(&d_full, (unsigned)(it & 1));
      asm volatile("tcgen05.fence::after_thread_sync;");
      double* ctile = a.C ? a.C + (p0 >> 5) * (int64_t)Mp * 32 : nullptr;
      double psi = 0.0;
      for (int c0 = 0; c0 < Mp; c0 += 32) {
        uint32_t v[32];
        g32_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)c0, v);
        const float* zh = a.Zh + ((p0 >> 5) * (int64_t)Mp + c0) * 32 + lane;
        const float* zl = a.Zl + ((p0 >> 5) * (int64_t)Mp + c0) * 32 + lane;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int row = c0 + j;
          const float cv = __uint_as_float(v[j]);
          if (ctile) ctile[row * 32 + (lane ^ ((row & 3) << 2))] = (double)cv;
          if (a.psi) psi = fma((double)zh[j * 32] + (double)zl[j * 32], (double)cv, psi);
        }
      }
      if (a.psi && n < a.npts) a.psi[n] = psi;
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncwarp();
      if (lane == 0) mbar_arrive(&d_empty);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// x -> (hi, lo) float planes, both exactly representable in TF32 (10 explicit mantissa bits): hi = x rounded to nearest,
// lo = (x - hi) rounded to nearest.  Rounding (not truncating) matters: a truncated lo is biased towards zero by up to
// 2^-22 |x|, the bias does not average out over the sum and the eigenvalue gradient - a difference of large terms - amplified
// it to 1.5e-5 (measured, r02h); rounded, the representation error is +-2^-24 |x| and unbiased.
__device__ __forceinline__ float tf32_round(float f) { return __uint_as_float((__float_as_uint(f) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ void tf32_split(double x, float& hi, float& lo) {
  hi = tf32_round((float)x);
  lo = tf32_round((float)(x - (double)hi));
}
__global__ void tf32_split_kernel(const double* __restrict__ x, float* hi, float* lo, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) tf32_split(x[i], hi[i], lo[i]);
}
#endif  // __CUDACC__

}  // namespace gpfy
