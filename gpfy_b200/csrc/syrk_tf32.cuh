// K_S32  the weighted Gram matrix of the reverse pass on tcgen05 (f32 instantiation, GPFY_PREC_TF32X3):
//     G = sum_n wq_n z_n z_n^T  [Mp, Mp],     bz = sum_n wp_n z_n  [Mp]          (dSigma, dmu, dE come from these)
// Operands are the hi / lo TF32 planes Zh / Zl [Mp][ld] the forward kernel writes (points contiguous = K-major for this
// product, whose contraction index is the point).  A CTA (h, ks) owns the column part h (NC columns of G) and the point
// range ks of the chunk.  Per 32-point k-block:
//   TMA    : all Mp rows x 32 points of both planes -> the A stage (128 B swizzled rows, UTMALDG)
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

constexpr int S32_THREADS = 192, S32_KF = 32, S32_NS = 2;

struct Syrk32Maps { CUtensorMap zh, zl; };   // [Mp rows][ld cols] float, box 32 points x (Mp / nbox) rows, 128 B swizzle

struct Syrk32Args {
  const double* wq;    // [ld]
  const double* wp;    // [ld]
  int Mp, NC, H;       // columns per part (multiple of 16), parts
  int nbox;            // TMA boxes per plane and k-block (rows per box = Mp / nbox <= 256)
  int ntile;           // 128-row tiles of G: row offsets 0, 128, ..., the last one at Mp - 128
  int64_t npts;        // valid points of this launch
  int64_t range;       // points per split (multiple of 32)
  double* Gpart;       // [KS][Mp][Mp] fp64, accumulated
  double* bzpart;      // [KS][Mp]
};

__host__ __device__ static inline int s32_stage_bytes(int Mp, int NC) { return 2 * Mp * 128 + 2 * NC * 128; }
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
  __shared__ __align__(16) double wsm[4][2][32];   // per scale warp: wq / wp of the k-block's 32 points
  const int Mp = a.Mp, NC = a.NC;
  const int stage_bytes = s32_stage_bytes(Mp, NC);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = (int)blockIdx.x % a.H, ks = (int)blockIdx.x / a.H;
  const int64_t p0 = (int64_t)ks * a.range;
  int64_t p1 = p0 + a.range;
  if (p1 > a.npts) p1 = a.npts;
  const int kt = p1 > p0 ? (int)((p1 - p0 + 31) / 32) : 0;          // 32-point k-blocks of this CTA
  const int nflush = (kt + S32_KF - 1) / S32_KF;

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
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      const int rows_box = Mp / a.nbox;
      for (int kb = 0; kb < kt; ++kb) {
        const int st = kb % S32_NS;
        if (kb >= S32_NS) mbar_wait(&empty[st], (unsigned)(((kb / S32_NS) - 1) & 1));
        uint8_t* dst = smem + st * stage_bytes;
        mbar_expect_tx(&full_a[st], (unsigned)(2 * Mp * 128));
        const int c0 = (int)(p0 + 32 * (int64_t)kb);
        for (int b = 0; b < a.nbox; ++b) {
          g32_tma_2d(dst + b * rows_box * 128, &maps.zh, c0, b * rows_box, &full_a[st]);
          g32_tma_2d(dst + Mp * 128 + b * rows_box * 128, &maps.zl, c0, b * rows_box, &full_a[st]);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      // both operands K-major (the contraction runs over the points), D f32, M = 128, N = NC
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      for (int kb = 0; kb < kt; ++kb) {
        const int st = kb % S32_NS;
        const int inflush = kb % S32_KF;
        if (inflush == 0 && kb > 0) mbar_wait(&d_empty, (unsigned)(((kb / S32_KF) - 1) & 1));   // accumulators drained
        mbar_wait(&full_a[st], (unsigned)((kb / S32_NS) & 1));
        mbar_wait(&full_b[st], (unsigned)((kb / S32_NS) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t ah = g32_smem_u32(smem + st * stage_bytes), al = ah + Mp * 128;
        const uint32_t bh = ah + 2 * Mp * 128, bl = bh + NC * 128;
#pragma unroll
        for (int k8 = 0; k8 < 4; ++k8) {
          const uint64_t dbh = g32_desc(bh + k8 * 32, 16, 1024, 2), dbl = g32_desc(bl + k8 * 32, 16, 1024, 2);
          for (int t = 0; t < a.ntile; ++t) {
            const int r0 = (t == a.ntile - 1) ? Mp - 128 : 128 * t;
            const uint64_t dah = g32_desc(ah + r0 * 128 + k8 * 32, 16, 1024, 2), dal = g32_desc(al + r0 * 128 + k8 * 32, 16, 1024, 2);
            const uint32_t d = tmem + (uint32_t)(t * NC);
            g32_mma(d, dal, dbh, idesc, (inflush | k8) ? 1u : 0u);
            g32_mma(d, dah, dbl, idesc, 1u);
            g32_mma(d, dah, dbh, idesc, 1u);
          }
        }
        g32_commit(&empty[st]);
        if (inflush == S32_KF - 1 || kb == kt - 1) g32_commit(&d_full);
      }
    }
  } else {
    // ================= scale + flush warps (2..5) =================
    const int q = warp & 3, t128 = (warp - 2) * 32 + lane;           // q: TMEM lane quadrant; t128: 0..127
    const int items = NC * 8;                                         // (row of the part, 16-byte chunk)
    double bzacc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) bzacc[i] = 0.0;
    double* gp = a.Gpart + (int64_t)ks * Mp * Mp;
    double (*wv)[32] = wsm[warp - 2];
    // the per-point weights of the next k-block are fetched one k-block ahead (one coalesced load per warp) and handed to
    // the items through a per-warp shared-memory copy: the item loop itself touches no global memory
    double wq_n = 0.0, wp_n = 0.0;
    if (kt > 0 && p0 + lane < a.npts) { wq_n = a.wq[p0 + lane]; wp_n = a.wp[p0 + lane]; }
    for (int kb = 0; kb < kt; ++kb) {
      const int st = kb % S32_NS;
      uint8_t* sa = smem + st * stage_bytes;
      uint8_t* sb = sa + 2 * Mp * 128;
      const int64_t n0 = p0 + 32 * (int64_t)kb;
      __syncwarp();
      wv[0][lane] = wq_n; wv[1][lane] = wp_n;
      __syncwarp();
      wq_n = 0.0; wp_n = 0.0;
      if (kb + 1 < kt && n0 + 32 + lane < a.npts) { wq_n = a.wq[n0 + 32 + lane]; wp_n = a.wp[n0 + 32 + lane]; }
      mbar_wait(&full_a[st], (unsigned)((kb / S32_NS) & 1));
#pragma unroll
      for (int i = 0; i < 16; ++i) {          // fully unrolled so that bzacc stays in registers (items <= 16 * 128)
        const int it = t128 + 128 * i;
        if (it >= items) break;
        const int rl = it >> 3, cp = it & 7;
        const int j = h * NC + rl;
        float4 oh = make_float4(0.f, 0.f, 0.f, 0.f), ol = oh;
        if (j < Mp) {
          const float4 vh = *reinterpret_cast<const float4*>(sa + j * 128 + cp * 16);
          const float4 vl = *reinterpret_cast<const float4*>(sa + Mp * 128 + j * 128 + cp * 16);
          const int pc = 4 * (cp ^ (j & 7));                          // the four points of this (swizzled) chunk
          const double2 w01 = *reinterpret_cast<const double2*>(&wv[0][pc]), w23 = *reinterpret_cast<const double2*>(&wv[0][pc + 2]);
          const double2 u01 = *reinterpret_cast<const double2*>(&wv[1][pc]), u23 = *reinterpret_cast<const double2*>(&wv[1][pc + 2]);
          const double w[4] = {w01.x, w01.y, w23.x, w23.y}, u[4] = {u01.x, u01.y, u23.x, u23.y};
          const double z0 = (double)vh.x + (double)vl.x, z1 = (double)vh.y + (double)vl.y;
          const double z2 = (double)vh.z + (double)vl.z, z3 = (double)vh.w + (double)vl.w;
          tf32_split(z0 * w[0], oh.x, ol.x);
          tf32_split(z1 * w[1], oh.y, ol.y);
          tf32_split(z2 * w[2], oh.z, ol.z);
          tf32_split(z3 * w[3], oh.w, ol.w);
          bzacc[i] += (z0 * u[0] + z1 * u[1]) + (z2 * u[2] + z3 * u[3]);
        }
        *reinterpret_cast<float4*>(sb + rl * 128 + cp * 16) = oh;
        *reinterpret_cast<float4*>(sb + NC * 128 + rl * 128 + cp * 16) = ol;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of B visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_b[st]);

      if ((kb % S32_KF) == S32_KF - 1 || kb == kt - 1) {
        // ---- flush: TMEM -> += Gpart (fp64).  Tile t covers rows r0 .. r0 + 127; the last tile only contributes the rows
        // its predecessor does not cover ----
        const int fl = kb / S32_KF;
        mbar_wait(&d_full, (unsigned)(fl & 1));
        asm volatile("tcgen05.fence::after_thread_sync;");
        for (int t = 0; t < a.ntile; ++t) {
          const int r0 = (t == a.ntile - 1) ? Mp - 128 : 128 * t;
          const int row = r0 + 32 * q + lane;
          const bool mine = (t < a.ntile - 1) || a.ntile == 1 || row >= 128 * (a.ntile - 1);
          for (int c0 = 0; c0 < NC; c0 += 32) {
            uint32_t v[32];
            g32_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(t * NC + c0), v);
            if (mine) {
              double* dst = gp + (int64_t)row * Mp + h * NC + c0;
#pragma unroll
              for (int jj = 0; jj < 32; jj += 2) {
                if (c0 + jj < NC && h * NC + c0 + jj < Mp) {
                  double2 o = *reinterpret_cast<double2*>(dst + jj);
                  o.x += (double)__uint_as_float(v[jj]);
                  o.y += (double)__uint_as_float(v[jj + 1]);
                  *reinterpret_cast<double2*>(dst + jj) = o;
                }
              }
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
    for (int i = 0; i < 16; ++i) {
      const int it = t128 + 128 * i;
      if (it >= items) break;                 // items is a multiple of 128: uniform over the warp
      double s = bzacc[i];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      const int j = h * NC + (it >> 3);
      if ((it & 7) == 0 && j < Mp) bzp[j] += s;
    }
  }
  (void)nflush;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// Gz [Mp][Mp] = sum_ks Gpart[ks], bz [Mp] = sum_ks bzpart[ks]   (fixed order: deterministic)
__global__ void reduce_g32_kernel(const double* Gpart, const double* bzpart, double* Gz, double* bz, int Mp, int KS) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t MM = (int64_t)Mp * Mp;
  if (e < MM) {
    double s = 0.0;
    for (int k = 0; k < KS; ++k) s += Gpart[(int64_t)k * MM + e];
    Gz[e] = s;
  }
  if (e < Mp) {
    double s = 0.0;
    for (int k = 0; k < KS; ++k) s += bzpart[(int64_t)k * Mp + e];
    bz[e] = s;
  }
}
#endif  // __CUDACC__

}  // namespace gpfy
