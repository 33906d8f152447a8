// K_S32  the weighted Gram matrix of the reverse pass on tcgen05 (f32 instantiation, GPFY_PREC_TF32X3):
//     G = sum_n wq_n z_n z_n^T  [Mp, Mp],     bz = sum_n wp_n z_n  [Mp]          (dSigma, dmu, dE come from these)
// Operands are the hi / lo TF32 planes Zh / Zl [Mp][ld] the forward kernel writes (points contiguous = K-major for this
// product, whose contraction index is the point).  A CTA (h, ks) owns the column part h (NC columns of G) and the point
// range ks of the chunk.  Per 16-point k-block (four of them in flight: the kernel is bound by the latency of the
// TMA -> scale -> MMA chain, not by bandwidth - with two 32-point stages ncu showed 13 % issue activity and 22 ms per step):
//   TMA    : all Mp rows x 16 points of both planes -> the A stage (64 B swizzled rows, SWIZZLE_64B, UTMALDG)
//   scale  : four warps read the part's NC rows back from the A stage, multiply by wq_n in fp64, split into TF32 hi / lo
//            again and write the B stage in the same swizzled layout (the weighted operand never exists in HBM); the
//            same pass accumulates bz for the part's rows in fp64 registers
//   MMA    : for every 128-row tile t of G: D_t[128 x NC] += A_t B^T as Ah Bh + Ah Bl + Al Bh (kind::tf32, fp32 in TMEM);
//            the last tile starts at Mp - 128, overlapping its predecessor, so every tile is a full M = 128 instruction
//   flush  : every S32_KF k-blocks (1024 points) the accumulators are drained (tcgen05.ld) and ADDED IN FP64 to the
//            CTA's slot of Gpart, so fp32 rounding never compounds over more than 1024 points
// No symmetry is exploited: at ~0.9 PFLOP/s of TF32 the full product costs less than the HBM stream of the planes.
#pragma once
#include "gemm_tf32.cuh"

namespace gpfy {

// S32_KP points per k-block: 32 (128 B rows, SWIZZLE_128B, 2 stages) or 16 (64 B rows, SWIZZLE_64B, 4 stages; measured slower:
// the fixed cost per k-block dominates, not the pipeline depth)
constexpr int S32_THREADS = 192, S32_KP = 32, S32_RB = S32_KP * 4, S32_KF = 1024 / S32_KP, S32_NS = 64 / S32_KP;
constexpr uint32_t S32_SBO = 8 * S32_RB, S32_LAYOUT = S32_KP == 32 ? 2u : 4u;   // 8-row atoms; layout type of the swizzle

struct Syrk32Maps { CUtensorMap zh, zl; };   // blocked planes viewed as [tiles * Mp rows][32 cols] float, box S32_KP x (Mp / nbox) rows

struct Syrk32Args {
  const double* wq;    // [ld]
  const double* wp;    // [ld]
  int Mp, NC, H;       // columns per part (multiple of 16), parts
  int nbox;            // TMA boxes per plane and k-block (rows per box = Mp / nbox <= 256)
  int ntile;           // 128-row tiles of G: row offsets 0, 128, ..., the last one at Mp - 128
  int64_t npts;        // valid points of this launch
  int64_t range;       // points per split (multiple of 32)
  int kf;              // k-blocks between two drains of the fp32 accumulators
  double* Gpart;       // [KS][ntile][Mp cols][128 rows of the tile] fp64, accumulated ("TMEM-native": for one column the 32
                       // lanes of a flushing warp are contiguous, so the read-modify-write is coalesced)
  double* bzpart;      // [KS][Mp]
};

__host__ __device__ static inline int s32_stage_bytes(int Mp, int NC) { return 2 * Mp * S32_RB + 2 * NC * S32_RB; }
static inline int s32_smem_bytes(int Mp, int NC) { return 2048 + S32_NS * s32_stage_bytes(Mp, NC); }
// column parts: the fewest H (<= 4) whose two stages fit in shared memory and whose tiles fit in the 512 TMEM columns
static inline bool s32_plan(int Mp, int* H_out, int* NC_out) {
  if (Mp < 128 || Mp > 512) return false;
  const int ntile = (Mp + 127) / 128;
  for (int H = 1; H <= 4; ++H) {
    const int NC = ((Mp + H - 1) / H + 15) / 16 * 16;
    if (NC > 256 || ntile * NC > 512) continue;
    if (s32_smem_bytes(Mp, NC) + 4096 > 227 * 1024) continue;   // + the kernel's static shared memory (barriers, weights)
    *H_out = H; *NC_out = NC;
    return true;
  }
  return false;
}

#ifdef __CUDACC__
__global__ void __launch_bounds__(S32_THREADS, 1) syrk_tf32x3_kernel(const __grid_constant__ Syrk32Maps maps, const __grid_constant__ Syrk32Args a) {
  extern __shared__ __align__(1024) uint8_t s32_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)s32_smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_a[S32_NS], full_b[S32_NS], empty[S32_NS], d_full, d_empty;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float wsm[4][2][32];    // per scale warp: wq / wp of the k-block's points, rounded to fp32
  const int Mp = a.Mp, NC = a.NC;
  const int stage_bytes = s32_stage_bytes(Mp, NC);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler (see g32_elect_one)
  const int h = (int)blockIdx.x % a.H, ks = (int)blockIdx.x / a.H;
  const int64_t p0 = (int64_t)ks * a.range;
  int64_t p1 = p0 + a.range;
  if (p1 > a.npts) p1 = a.npts;
  const int kt = p1 > p0 ? (int)((p1 - p0 + S32_KP - 1) / S32_KP) : 0;   // k-blocks of this CTA
  const int KF = a.kf;
  const int foff = (int)(((unsigned)blockIdx.x * 7u) % (unsigned)KF);   // this CTA's offset inside the drain period
  const int nskip = 0;
  const int nflush = (kt + KF - 1) / KF;

  if (tid == 0) {
    for (int s = 0; s < S32_NS; ++s) { mbar_init(&full_a[s], 1); mbar_init(&full_b[s], 4); mbar_init(&empty[s], 1); }
    mbar_init(&d_full, 1);
    mbar_init(&d_empty, 4);
  }
  if (warp == 1) {
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
      const int rows_box = Mp / a.nbox;
      for (int kb = 0; kb < kt; ++kb) {
        const int st = kb % S32_NS;
        if (kb >= S32_NS) mbar_wait(&empty[st], (unsigned)(((kb / S32_NS) - 1) & 1));
        uint8_t* dst = smem + st * stage_bytes;
        // planes are tile-contiguous ([n / 32][Mp][32]): the k-block's rows are ONE contiguous slab of HBM per plane
        const int64_t n0 = p0 + S32_KP * (int64_t)kb;
        const int col = (int)(n0 & 31), row0 = (int)((n0 >> 5) * Mp);
        if (g32_elect_one()) {
          mbar_expect_tx(&full_a[st], (unsigned)(2 * Mp * S32_RB));
          for (int b = 0; b < a.nbox; ++b) {
            g32_tma_2d(dst + b * rows_box * S32_RB, &maps.zh, col, row0 + b * rows_box, &full_a[st]);
            g32_tma_2d(dst + Mp * S32_RB + b * rows_box * S32_RB, &maps.zl, col, row0 + b * rows_box, &full_a[st]);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (warp-uniform loop, one elected lane issues) =================
    {
      // both operands K-major (the contraction runs over the points), D f32, M = 128, N = NC
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t hiD = g32_desc_hi(S32_SBO, S32_LAYOUT);
      uint32_t toff[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) toff[t] = (uint32_t)((((t == a.ntile - 1) ? Mp - 128 : 128 * t) * S32_RB) >> 4);
      for (int kb = 0; kb < kt; ++kb) {
        const int st = kb % S32_NS;
        // drains are staggered over the CTAs (foff): 148 CTAs draining at once turned the read-modify-write of Gpart into
        // an L2 storm (ncu r02k: 30 % of the kernel's samples on the drain's loads)
        const int inflush = (kb + foff) % KF;
        if (inflush == 0 && kb > 0) mbar_wait(&d_empty, (unsigned)((((kb + foff) / KF) - 1 - nskip) & 1));   // accumulators drained
        mbar_wait(&full_a[st], (unsigned)((kb / S32_NS) & 1));
        mbar_wait(&full_b[st], (unsigned)((kb / S32_NS) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t ah = g32_smem_u32(smem + st * stage_bytes), al = ah + Mp * S32_RB;
        const uint32_t bh = ah + 2 * Mp * S32_RB, bl = bh + NC * S32_RB;
        // K-major operands (swizzled rows of S32_RB bytes, 8-row atoms): 8 points = 32 B inside a row (+2 in the descriptor's
        // address field); tile t of the A operand starts r0 rows further down (+toff[t])
        const uint32_t ahl = g32_desc_lo(ah, 16), all = g32_desc_lo(al, 16), bhl = g32_desc_lo(bh, 16), bll = g32_desc_lo(bl, 16);
        if (g32_elect_one()) {
#pragma unroll
          for (int k8 = 0; k8 < S32_KP / 8; ++k8) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              if (t < a.ntile) {
                const uint32_t d = tmem + (uint32_t)(t * NC);
                g32_mma2(d, all + toff[t] + 2 * k8, hiD, bhl + 2 * k8, hiD, idesc, ((inflush && kb) | k8) ? 1u : 0u);
                g32_mma2(d, ahl + toff[t] + 2 * k8, hiD, bll + 2 * k8, hiD, idesc, 1u);
                g32_mma2(d, ahl + toff[t] + 2 * k8, hiD, bhl + 2 * k8, hiD, idesc, 1u);
              }
            }
          }
          g32_commit(&empty[st]);
          if (inflush == KF - 1 || kb == kt - 1) g32_commit(&d_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ================= scale + flush warps (2..5) =================
    const int q = warp & 3, t128 = (warp - 2) * 32 + lane;           // q: TMEM lane quadrant; t128: 0..127
    constexpr int CH = S32_RB / 16;                                   // 16-byte chunks (4 points) per row
    const int items = NC * CH;                                        // (row of the part, chunk); a multiple of 64
    constexpr int NI = 2 * CH;                                        // items per thread at most (NC <= 256)
    double bzacc[NI];
    float bzf[NI];                            // fp32 partial sums of one drain period, widened at the drain
#pragma unroll
    for (int i = 0; i < NI; ++i) { bzacc[i] = 0.0; bzf[i] = 0.f; }
    double* gp = a.Gpart + (int64_t)ks * a.ntile * Mp * 128;
    float (*wv)[32] = wsm[warp - 2];
    // the per-point weights of the next k-block are fetched one k-block ahead (one coalesced load per warp) and handed to
    // the items through a per-warp shared-memory copy: the item loop itself touches no global memory
    int nfl = 0;                              // drains done so far (parity of d_full)
    double wq_n = 0.0, wp_n = 0.0;
    if (kt > 0 && lane < S32_KP && p0 + lane < a.npts) { wq_n = a.wq[p0 + lane]; wp_n = a.wp[p0 + lane]; }
    for (int kb = 0; kb < kt; ++kb) {
      const int st = kb % S32_NS;
      uint8_t* sa = smem + st * stage_bytes;
      uint8_t* sb = sa + 2 * Mp * S32_RB;
      const int64_t n0 = p0 + S32_KP * (int64_t)kb;
      __syncwarp();
      wv[0][lane] = (float)wq_n; wv[1][lane] = (float)wp_n;
      __syncwarp();
      wq_n = 0.0; wp_n = 0.0;
      if (kb + 1 < kt && lane < S32_KP && n0 + S32_KP + lane < a.npts) { wq_n = a.wq[n0 + S32_KP + lane]; wp_n = a.wp[n0 + S32_KP + lane]; }
      mbar_wait(&full_a[st], (unsigned)((kb / S32_NS) & 1));
#pragma unroll
      for (int i = 0; i < NI; ++i) {          // fully unrolled so that bzacc stays in registers (items <= NI * 128)
        const int it = t128 + 128 * i;
        if (it >= items) break;               // whole warps drop out (items is a multiple of 64)
        const int rl = it / CH, cp = it % CH;
        const int j = h * NC + rl;
        float4 oh = make_float4(0.f, 0.f, 0.f, 0.f), ol = oh;
        if (j < Mp) {
          const float4 vh = *reinterpret_cast<const float4*>(sa + j * S32_RB + cp * 16);
          const float4 vl = *reinterpret_cast<const float4*>(sa + Mp * S32_RB + j * S32_RB + cp * 16);
          const int pc = 4 * (cp ^ (S32_KP == 32 ? (j & 7) : ((j >> 1) & 3)));   // the four points of this (swizzled) chunk
          // hi + lo is exact in fp32 (22 significant bits), the weighting is one fp32 multiply (relative 6e-8, below the
          // 2^-22 the re-split keeps) and the split is exact fp32 arithmetic: the scale pass runs on the FP32 pipe - the
          // fp64 version spent its time in F2F conversions (ncu r02i: 14 % of the samples)
          const float4 w = *reinterpret_cast<const float4*>(&wv[0][pc]), u = *reinterpret_cast<const float4*>(&wv[1][pc]);
          const float z0 = vh.x + vl.x, z1 = vh.y + vl.y, z2 = vh.z + vl.z, z3 = vh.w + vl.w;
          const float s0 = z0 * w.x, s1 = z1 * w.y, s2 = z2 * w.z, s3 = z3 * w.w;
          oh.x = tf32_round(s0); ol.x = tf32_round(s0 - oh.x);
          oh.y = tf32_round(s1); ol.y = tf32_round(s1 - oh.y);
          oh.z = tf32_round(s2); ol.z = tf32_round(s2 - oh.z);
          oh.w = tf32_round(s3); ol.w = tf32_round(s3 - oh.w);
          bzf[i] += (z0 * u.x + z1 * u.y) + (z2 * u.z + z3 * u.w);
        }
        *reinterpret_cast<float4*>(sb + rl * S32_RB + cp * 16) = oh;
        *reinterpret_cast<float4*>(sb + NC * S32_RB + rl * S32_RB + cp * 16) = ol;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of B visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_b[st]);

      if (((kb + foff) % KF) == KF - 1 || kb == kt - 1) {
        // ---- flush: TMEM -> += Gpart (fp64).  Tile t covers rows r0 .. r0 + 127; the last tile only contributes the rows
        // its predecessor does not cover ----
#pragma unroll
        for (int i = 0; i < NI; ++i) { bzacc[i] += (double)bzf[i]; bzf[i] = 0.f; }
        mbar_wait(&d_full, (unsigned)(nfl & 1));
        ++nfl;
        asm volatile("tcgen05.fence::after_thread_sync;");
        for (int t = 0; t < a.ntile; ++t) {
          double* gt = gp + ((int64_t)t * Mp + h * NC) * 128 + 32 * q + lane;   // [col][lane]: 256 B per warp and column
          for (int c0 = 0; c0 < NC; c0 += 32) {
            uint32_t v[32];
            g32_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(t * NC + c0), v);
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              if (c0 + jj < NC && h * NC + c0 + jj < Mp) gt[(int64_t)(c0 + jj) * 128] += (double)__uint_as_float(v[jj]);
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncwarp();
        if (lane == 0) mbar_arrive(&d_empty);
      }
    }
    // ---- bz of the part's rows: sum the 8 chunk lanes of every row (aligned groups of 8 lanes) ----
    double* bzp = a.bzpart + (int64_t)ks * Mp;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int it = t128 + 128 * i;
      if (it >= items) break;
      double s = bzacc[i];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (CH == 8) s += __shfl_xor_sync(0xffffffffu, s, 4);
      const int j = h * NC + it / CH;
      if ((it % CH) == 0 && j < Mp) bzp[j] += s;
    }
  }
  (void)nflush;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// Gz [Mp][Mp] = sum_ks Gpart[ks] (tile-native layout -> row-major; the last tile contributes only the rows its predecessors
// do not cover), bz [Mp] = sum_ks bzpart[ks]   (fixed order: deterministic)
__global__ void reduce_g32_kernel(const double* Gpart, const double* bzpart, double* Gz, double* bz, int Mp, int KS, int ntile) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t MM = (int64_t)Mp * Mp;
  if (e < MM) {
    const int col = (int)(e / Mp), row = (int)(e % Mp);          // consecutive threads walk the rows: coalesced reads
    int t = row / 128;
    if (t >= ntile - 1) t = ntile - 1;
    const int r0 = (t == ntile - 1) ? Mp - 128 : 128 * t;
    const double* src = Gpart + ((int64_t)t * Mp + col) * 128 + (row - r0);
    const int64_t slot = (int64_t)ntile * Mp * 128;
    double s = 0.0;
    for (int k = 0; k < KS; ++k) s += src[(int64_t)k * slot];
    Gz[(int64_t)row * Mp + col] = s;
  }
  if (e < Mp) {
    double s = 0.0;
    for (int k = 0; k < KS; ++k) s += bzpart[(int64_t)k * Mp + e];
    bz[e] = s;
  }
}
#endif  // __CUDACC__

}  // namespace gpfy
