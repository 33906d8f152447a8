// K_B4  backward through the zonal map for every P = 1 shape (any Mp, ambient dim up to 36, either likelihood).
//
// The headline kernel (zonal_bwd3) recomputes t = Vhat xhat and the Gegenbauer recurrence per tile, which ties it to
// small ambient dimensions (xhat in registers) and to Mp <= 288 (a whole [Mp x 32] tile pair in shared memory).  Here
// the forward kernel also stores the derivative panel ZP = d/dt C_l^alpha(t) (gegenbauer.py:73-81) next to Z, in the same
// tile-contiguous swizzled layout the GEMM uses for C, so the backward pass is
//   phase 1 : dT = (cq C + cp mu2) * ZP                 elementwise on a [RG x 32] slab of rows (lane = point)
//   phase 2a: dVhat[rows, dim] += dT[rows, 32] Xhat[32, dim]      FP64 tensor cores, accumulators live in registers over
//                                                                 all tiles of the CTA (VJP of spherical_harmonics.py:308)
//   phase 2b: dxhat[dim, 32]   += Vhat^T[dim, rows] dT[rows, 32]  every warp owns whole 8 x 8 output blocks (full K), so
//                                                                 there is no cross-warp reduction and no parking buffer
//   tail    : du = (dxhat - xhat (xhat . dxhat)) / r + dr xhat -> dw, db sums   (VJP of spherical.py:243-249)
// with no dependence of the per-thread state on dim or Mp: a tile is walked in G row groups of RG rows, each one stage
// of a TMA + mbarrier pipeline (C slab + ZP slab, contiguous in HBM), so Mp = 352 (13 degrees) or 576 (T = 64) fit.
// Roles: warp 0 issues the TMA copies and does the per-tile tail; warps 1..11 compute.  The kernel trades
// 16 Mp bytes/point of HBM reads for the FP64 work of the recomputation; HBM has the room (the step is FP64-bound).
#pragma once
#include "svgp_kernels.cuh"
#include "zonal_bwd3.cuh"

namespace gpfy {

constexpr int ZB4_THREADS = 384, ZB4_CW = 11, ZB4_MAX_NST = 4;

struct ZonalBwd4Args {
  const double* C;       // [tiles][Mp][32] blocked + swizzled (GemmArgs::c_blocked)
  const double* ZP;      // same layout: d/dt C_l^alpha(t) (forward kernel)
  const double* wp;      // [ldz]  p r
  const double* cq;      // [ldz]  2 q r^2
  const double* dr;      // [ldz]  direct d/dr terms (lik_point_kernel)
  const double* XH;      // [ldz][DIMP]
  const double* R;       // [npts]
  const double* vhat_p;  // [Mp][DIMP], mu2 [Mp] right behind it
  int Mp, M, d;
  int G, RG;             // row groups per tile, rows per group (Mp = G * RG, RG a multiple of 8)
  int NST, HB;           // pipeline stages (slabs in flight), header ring size (>= ceil(NST / G) + 1)
  int64_t npts;
  double* partials;      // [gridDim][npart]: PART_DB, PART_DW.. accumulated here
  int npart;
  double* dVpart;        // [gridDim][Mp][CT * 8]
};

static inline int zb4_header_ring(int G, int NST) { return (NST + G - 1) / G + 1; }
static inline int zb4_smem_bytes(int Mp, int dimp, int RG, int NST) {
  const int hdr = 128 + 32 * dimp;
  const int dbl = Mp * dimp + Mp + NST * 2 * RG * 32 + zb4_header_ring(Mp / RG, NST) * hdr + 2 * dimp * 32 + (1 + dimp) * 32;
  return dbl * 8 + 64;
}
// (rows per group, stages): three stages in flight if some split fits, else two; within that the fewest groups
static inline bool zb4_plan(int Mp, int dimp, int* RG_out, int* NST_out) {
  for (int NST = 3; NST >= 2; --NST)
    for (int G = 1; G <= Mp / 8; ++G) {
      if (Mp % G) continue;
      const int RG = Mp / G;
      if (RG % 8) continue;
      if (zb4_smem_bytes(Mp, dimp, RG, NST) <= 226 * 1024) { *RG_out = RG; *NST_out = NST; return true; }
    }
  return false;
}

template <int DIMP, int SLOTS>
__global__ void __launch_bounds__(ZB4_THREADS, 1) zonal_bwd4_kernel(const __grid_constant__ ZonalBwd4Args a) {
  constexpr int CT = (DIMP + 7) / 8;
  constexpr int HDR = 128 + 32 * DIMP;
  constexpr int NBLK = CT * 4;                               // 8 x 8 output blocks of dxhat [CT * 8, 32]
  constexpr int BSLOTS = (NBLK + ZB4_CW - 1) / ZB4_CW;
  extern __shared__ __align__(16) double smem[];
  __shared__ uint64_t bar_stage, bar_full[ZB4_MAX_NST], bar_empty[ZB4_MAX_NST], bar_tile[2], bar_tail[2];
  const int Mp = a.Mp, RG = a.RG, G = a.G, NST = a.NST, HB = a.HB;
  double* vs = smem;                               // [Mp][DIMP]
  double* mu2s = vs + Mp * DIMP;                   // [Mp]
  double* stg = mu2s + Mp;                         // [NST][ C slab RG x 32 | ZP slab RG x 32 ]
  double* hdr = stg + NST * 2 * RG * 32;           // [HB][wp(32) | cq(32) | r(32) | dr(32) | xhat(32 x DIMP)]
  double* dxs = hdr + HB * HDR;                    // [2][DIMP][32]
  double* lsum = dxs + 2 * DIMP * 32;              // [1 + DIMP][32] per-lane running sums of the tail

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g8 = lane >> 2, t4 = lane & 3;
  const int tiles = (int)((a.npts + 31) / 32);
  const int n_my = ((int)blockIdx.x < tiles) ? (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int n_stage = n_my * G;

  for (int i = tid; i < (1 + DIMP) * 32; i += ZB4_THREADS) lsum[i] = 0.0;
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], ZB4_CW); }
    mbar_init(&bar_tile[0], ZB4_CW); mbar_init(&bar_tile[1], ZB4_CW);
    mbar_init(&bar_tail[0], 1); mbar_init(&bar_tail[1], 1);
  }
  stage_params_tma(vs, a.vhat_p, (unsigned)((Mp * DIMP + Mp) * 8), &bar_stage);  // includes a __syncthreads

  if (warp == 0) {
    // ================= producer + tail =================
    const unsigned slab = (unsigned)RG * 256u;
    auto issue = [&](int s) {
      const int t = s / G, g = s - t * G, buf = s % NST;
      const int64_t p0 = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * 32;
      double* cb = stg + buf * 2 * RG * 32;
      fence_proxy_async();
      if (lane == 0) mbar_expect_tx(&bar_full[buf], 2u * slab + (g == 0 ? (unsigned)(HDR * 8) : 0u));
      __syncwarp();
      const int64_t off = ((p0 >> 5) * (int64_t)Mp + (int64_t)g * RG) * 32;
      const char* srcC = reinterpret_cast<const char*>(a.C + off);
      const char* srcZ = reinterpret_cast<const char*>(a.ZP + off);
      // pieces of <= 8 KB dealt to the lanes: C pieces on even lanes, ZP pieces on odd lanes
      const int half = lane >> 1;
      for (unsigned o = (unsigned)half * 8192u; o < slab; o += 16u * 8192u) {
        const unsigned n = min(8192u, slab - o);
        if (lane & 1) tma_bulk_g2s(reinterpret_cast<char*>(cb + RG * 32) + o, srcZ + o, n, &bar_full[buf]);
        else tma_bulk_g2s(reinterpret_cast<char*>(cb) + o, srcC + o, n, &bar_full[buf]);
      }
      if (g == 0) {
        double* hb = hdr + (t % HB) * HDR;
        if (lane == 27) tma_bulk_g2s(hb + 64, a.R + p0, 256, &bar_full[buf]);     // r and dr ride along for the tail (round 2: its own
        if (lane == 28) tma_bulk_g2s(hb + 96, a.dr + p0, 256, &bar_full[buf]);    // global loads kept the compute warps waiting on bar_tail)
        if (lane == 29) tma_bulk_g2s(hb, a.wp + p0, 256, &bar_full[buf]);
        if (lane == 30) tma_bulk_g2s(hb + 32, a.cq + p0, 256, &bar_full[buf]);
        if (lane == 31) tma_bulk_g2s(hb + 128, a.XH + p0 * DIMP, 32 * DIMP * 8, &bar_full[buf]);
      }
    };
    auto tail = [&](int t) {
      const int64_t n = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * 32 + lane;
      const bool valid = n < a.npts;
      double xh[DIMP], dxh[DIMP];
      mbar_wait(&bar_tile[t & 1], (unsigned)((t >> 1) & 1));
      // r, dr and the xhat rows of the tile sit in its header (TMA, with the first row group): no global load in the tail
      const double* hb = hdr + (t % HB) * HDR;
      const double r = valid ? hb[64 + lane] : 1.0, drv = valid ? hb[96 + lane] : 0.0;
#pragma unroll
      for (int c = 0; c < DIMP; c += 2) {
        const double2 v2 = *reinterpret_cast<const double2*>(hb + 128 + lane * DIMP + c);
        xh[c] = valid ? v2.x : 0.0; xh[c + 1] = valid ? v2.y : 0.0;
      }
      const double* dx = dxs + (t & 1) * DIMP * 32;
#pragma unroll
      for (int c = 0; c < DIMP; ++c) dxh[c] = dx[c * 32 + lane];
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tail[t & 1]);   // dxs[t & 1] may be overwritten (tile t + 2)
      if (valid) {
        const double rinv = 1.0 / r;
        double dot = 0.0;
#pragma unroll
        for (int c = 0; c < DIMP; ++c) dot = fma(xh[c], dxh[c], dot);
#pragma unroll
        for (int c = 0; c < DIMP; ++c) {
          const double du = (dxh[c] - xh[c] * dot) * rinv + drv * xh[c];
          const double u = xh[c] * r;
          if (c < a.d) lsum[(1 + c) * 32 + lane] += du * u;      // d w_j = sum du_j u_j / (2 w_j)
          else if (c == a.d) lsum[lane] += du;                  // d b   = sum du_dim / (2 sqrt b)
        }
      }
    };
    int next_issue = 0, next_tail = 0;
    while (next_tail < n_my) {
      if (next_issue < n_stage) {
        const int buf = next_issue % NST;
        const bool hdr_free = (next_issue % G) != 0 || next_issue / G < HB || next_tail > next_issue / G - HB;   // the tail reads the header too
        const bool free_now = hdr_free && (next_issue < NST || mbar_try_wait(&bar_empty[buf], (unsigned)(((next_issue / NST) - 1) & 1)));
        if (free_now) { issue(next_issue++); continue; }
        if (!hdr_free) { tail(next_tail++); continue; }
      }
      if (mbar_try_wait(&bar_tile[next_tail & 1], (unsigned)((next_tail >> 1) & 1))) { tail(next_tail++); continue; }
      if (next_issue >= n_stage) { tail(next_tail++); }   // nothing left to issue: block in the tail's wait
    }
  } else {
    // ================= compute warps =================
    const int cw = warp - 1;
    const int nrt = Mp / 8;                                  // row tiles; warp cw owns cw, cw + 11, ...
    double accv[SLOTS][CT][2];
#pragma unroll
    for (int r = 0; r < SLOTS; ++r)
#pragma unroll
      for (int c = 0; c < CT; ++c) accv[r][c][0] = accv[r][c][1] = 0.0;
    double acc2[BSLOTS][2];
#pragma unroll
    for (int b = 0; b < BSLOTS; ++b) acc2[b][0] = acc2[b][1] = 0.0;

    for (int s = 0; s < n_stage; ++s) {
      const int t = s / G, g = s - t * G, buf = s % NST;
      const int64_t p0 = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * 32;
      const bool valid = (p0 + lane) < a.npts;
      const int nvalid = (int)((a.npts - p0) < 32 ? (a.npts - p0) : 32);
      double* cb = stg + buf * 2 * RG * 32;
      const double* zp = cb + RG * 32;
      const double* hb = hdr + (t % HB) * HDR;
      const int row0 = g * RG;
      mbar_wait(&bar_full[buf], (unsigned)((s / NST) & 1));

      // ---- phase 1: dT in place (lane = point); points past the end of the chunk give exactly zero ----
      const double cpw = valid ? hb[lane] : 0.0, cqw = valid ? hb[32 + lane] : 0.0;
#pragma unroll 4
      for (int r = cw; r < RG; r += ZB4_CW) {
        const int o = zb_at(r, lane);
        const double cv = cb[o], zv = zp[o];
        const double dz = fma(cqw, valid ? cv : 0.0, cpw * mu2s[row0 + r]);
        cb[o] = dz * (valid ? zv : 0.0);
      }
      asm volatile("bar.sync 1, %0;\n" ::"n"(ZB4_CW * 32) : "memory");

      // ---- phase 2a: dVhat[own row tiles of this group] += dT Xhat ----
#pragma unroll
      for (int sl = 0; sl < SLOTS; ++sl) {
        const int rt = cw + ZB4_CW * sl;
        const int lr = rt * 8 - row0;                        // first local row of the tile
        if (rt < nrt && lr >= 0 && lr < RG) {
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const int pt = 4 * k4 + t4;
            const double af = cb[zb_at(lr + g8, pt)];
#pragma unroll
            for (int c = 0; c < CT; ++c) {
              const double bf = (8 * c + g8 < DIMP && pt < nvalid) ? hb[128 + pt * DIMP + 8 * c + g8] : 0.0;
              dmma884(accv[sl][c][0], accv[sl][c][1], af, bf);
            }
          }
        }
      }
      // ---- phase 2b: own 8 x 8 blocks of dxhat[dim, 32] += Vhat^T dT over the rows of this group ----
#pragma unroll
      for (int bs = 0; bs < BSLOTS; ++bs) {
        const int b = cw + ZB4_CW * bs;
        if (b < NBLK) {
          const int mt = b >> 2, nt = b & 3;
          const bool dok = 8 * mt + g8 < DIMP;
          const double* vrow = vs + (row0 + t4) * DIMP + 8 * mt + g8;
#pragma unroll 4
          for (int k4 = 0; k4 < RG / 4; ++k4) {
            const double av = dok ? vrow[4 * k4 * DIMP] : 0.0;
            const double bf = cb[zb_at(4 * k4 + t4, 8 * nt + g8)];
            dmma884(acc2[bs][0], acc2[bs][1], av, bf);
          }
          if (g == G - 1) {
            if (t >= 2) mbar_wait(&bar_tail[t & 1], (unsigned)(((t >> 1) - 1) & 1));   // the tail of tile t - 2 has read dxs[t & 1]
            if (dok) *reinterpret_cast<double2*>(dxs + (t & 1) * DIMP * 32 + (8 * mt + g8) * 32 + 8 * nt + 2 * t4) = make_double2(acc2[bs][0], acc2[bs][1]);
            acc2[bs][0] = acc2[bs][1] = 0.0;
          }
        }
      }
      fence_proxy_async();   // generic-proxy writes of this buffer are ordered before the TMA refill
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bar_empty[buf]);
        if (g == G - 1) mbar_arrive(&bar_tile[t & 1]);
      }
    }

    // ---- write-out of dVhat (accumulating over chunks) ----
#pragma unroll
    for (int sl = 0; sl < SLOTS; ++sl) {
      const int rt = cw + ZB4_CW * sl;
      if (rt >= nrt) continue;
      const int row = rt * 8 + g8;
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        double2* dst = reinterpret_cast<double2*>(a.dVpart + ((int64_t)blockIdx.x * Mp + row) * (CT * 8) + c * 8 + 2 * t4);
        double2 o = *dst;
        o.x += accv[sl][c][0];
        o.y += accv[sl][c][1];
        *dst = o;
      }
    }
  }

  __syncthreads();
  if (warp < 2) {  // fixed-order reduction of the per-lane sums
    for (int k = PART_DB + warp; k < a.npart; k += 2) {
      const int src = k == PART_DB ? 0 : 1 + (k - PART_DW);
      const double sres = warp_sum(lsum[src * 32 + lane]);
      if (lane == 0) a.partials[(int64_t)blockIdx.x * a.npart + k] += sres;
    }
  }
}

}  // namespace gpfy
