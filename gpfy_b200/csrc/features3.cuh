// K_F3  fused feature matrix without CTA barriers (spherical_harmonics.py:290-320, :344-362, :385-387):
//     out[off_l + i, n] = scale_l * (r_n) * sum_k Linv_l[i][k] C_l^alpha(Vhat_l xhat_n)[k]
// The whitening is block diagonal: the output rows of degree l need the zonal values of degree l only.  So the work unit is
// (32-point tile, group of whole degrees) and ONE WARP carries a unit from the points' coordinates to the stored rows - zonal forward
// on the FP64 tensor cores into the warp's private shared-memory tile, __syncwarp, lower-triangular DMMA whitening out of that tile,
// store.  No warp ever waits for another: `features2_kernel` (one CTA per tile, phase A | barrier | phase B | barrier) spent 19 % of
// its stall samples at the two barriers and ran its DMMAs at 30 % of the pipe (profiles/ncu_full_r02a_features2_summary.txt).
// Every warp re-projects its tile's points (8 d bytes per point out of L1/L2, ~5 % of a unit's instructions).  Degrees are grouped on the
// host (`ft3_groups`) so that tiny degrees (m_0 = 1, m_1 = d + 1) ride with a full one; units go round-robin over all warps of the grid,
// consecutive units (same tile, next group) to the warps of one CTA.
// One CTA of 16 warps per SM.  When it fits next to the 16 private tiles, the lower triangles of all Linv_l are staged once per CTA in
// shared memory, row tile i of a degree as [8][8 i + 12] (8 i + 8 columns + 4 of padding: conflict-free A-fragment reads) - through L1 the
// whitening's A fragments hit only every second time (ncu r02x: L1 hit rate 60 % with the coordinates included, long-scoreboard stalls).
// Shapes: every m_l <= 32 (the private tile is 32 rows), padded dimension <= 32; the rest stays on features2_kernel.
#pragma once
#include <algorithm>
#include "zonal_fwd2.cuh"

namespace gpfy {

struct Features3Args {
  FeaturesArgs f;
  int ngrp;
  int linv_doubles;   // ft3_linv_doubles: size of the staged triangles
  unsigned char grp_lo[GPFY_MAX_DEGREES + 2];   // group gi = degrees grp_lo[gi] .. grp_lo[gi + 1] - 1
};

// host: consecutive degrees are merged into ~96-row groups of about equal size (three full 30-row degrees): a unit then pays for its
// tile's coordinates (a load from HBM / L2 nothing hides, ncu r02x: 12 % of the stall samples) once per ~3 k FP64 instructions
static inline bool ft3_groups(Features3Args* fa, const DegreeTable& deg) {
  int total = 0;
  for (int l = 0; l < deg.F; ++l) {
    if (deg.m[l] > 32) return false;
    total += deg.m[l];
  }
  const int G = std::max(1, (total + 48) / 96);
  int n = 0, cum = 0;
  fa->grp_lo[0] = 0;
  for (int l = 0; l < deg.F; ++l) {
    cum += deg.m[l];
    // close the group at the degree whose end is nearest to the next multiple of total / G
    if (n + 1 < G && l + 1 < deg.F && 2 * cum * G >= (2 * (n + 1)) * total - deg.m[l + 1] * G) fa->grp_lo[++n] = (unsigned char)(l + 1);
  }
  fa->grp_lo[++n] = (unsigned char)deg.F;
  fa->ngrp = n;
  return true;
}

constexpr int FT3_WARPS = 16;
// packed lower triangle of one degree with nt row tiles: tile i starts at 32 i^2 + 64 i, rows of 8 i + 12 doubles
static inline int ft3_tri_doubles(int nt) { return 32 * nt * nt + 64 * nt; }
static inline int ft3_linv_doubles(const DegreeTable& deg) {
  int n = 0;
  for (int l = 0; l < deg.F; ++l) n += ft3_tri_doubles((deg.m[l] + 7) / 8);
  return n;
}
static inline int ft3_smem_bytes(int Mp, int dimp, int linv_doubles) {
  const int dbl = Mp * dimp + FT3_WARPS * (32 * 32 + 32) + linv_doubles + 3 * (GPFY_MAX_DEGREES + 2) + 4;
  return dbl * 8 + 2 * (GPFY_MAX_DEGREES + 1) * 4 + 16;
}

// NQ k-steps of one 8-row tile of the whitening: A fragments (rows of Linv_l through L1, columns past m_l as zeros) fetched ahead of
// the tensor-core steps, B fragments from the warp's swizzled tile.  lrow = &Linv_l[row][t4]; `lane_row` = row inside the degree.
template <int NQ, bool LSM>
__device__ __forceinline__ void ft3_whiten_steps(double (&acc)[8], const double* __restrict__ lrow, bool rok, int ml, const double* zt,
                                                 const int (&bidx)[4]) {
  const int t4 = threadIdx.x & 3;
  double av[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    if constexpr (LSM) av[q] = lrow[4 * q];   // staged copy: rows / columns past m_l are stored as zeros
    else av[q] = (rok && 4 * q + t4 < ml) ? __ldg(lrow + 4 * q) : 0.0;
  }
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const double* zk = zt + q * (4 * FT_LD);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) dmma884(acc[2 * nt], acc[2 * nt + 1], av[q], zk[bidx[nt]]);
  }
}

template <int DIMP, bool LSM>
__global__ void __launch_bounds__(32 * FT3_WARPS, 1) features3_kernel(const __grid_constant__ Features3Args args3) {
  static_assert(DIMP <= 32, "the coordinates are staged in the warp's 32 x 32 tile");
  constexpr int KS = DIMP / 4;
  const FeaturesArgs& args = args3.f;
  const ZonalArgs& a = args.z;
  extern __shared__ __align__(16) double smem[];
  __shared__ uint64_t bar;
  const int Mp = a.Mp;
  double* vs = smem;                           // [Mp][DIMP]
  double* ztw = vs + Mp * DIMP;                // [16 warps][32][32] swizzled zonal tile of one degree (first the tile's coordinates)
  double* rsw = ztw + FT3_WARPS * 32 * 32;     // [16][32] radii (or 1)
  double* lts = rsw + FT3_WARPS * 32;          // packed lower triangles of Linv_l (LSM)
  double* leads = lts + (LSM ? args3.linv_doubles : 0);   // [MAX_DEGREES + 2]
  double* scl = leads + GPFY_MAX_DEGREES + 2;  // [MAX_DEGREES + 2] per-degree output scale
  double* betas = scl + GPFY_MAX_DEGREES + 2;  // [MAX_DEGREES + 6]
  int* offs = reinterpret_cast<int*>(betas + GPFY_MAX_DEGREES + 6);  // [F + 1]
  int* lbase = offs + GPFY_MAX_DEGREES + 1;    // [F] start of degree l in lts

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int F = a.deg.F;
  constexpr int NTHR = 32 * FT3_WARPS;
  for (int i = tid; i < GPFY_MAX_DEGREES + 2; i += NTHR) {
    leads[i] = args.mt.lead[i];
    scl[i] = (i < F && args.scale) ? args.scale[i] : 1.0;
  }
  for (int i = tid; i < GPFY_MAX_DEGREES + 6; i += NTHR) betas[i] = i < GPFY_MAX_DEGREES + 2 ? args.mt.beta[i] : 0.0;
  if (tid == 0) {
    int o = 0, lb = 0;
    for (int l = 0; l < F; ++l) {
      offs[l] = o;
      lbase[l] = lb;
      o += a.deg.m[l];
      const int nt = (a.deg.m[l] + 7) >> 3;
      lb += 32 * nt * nt + 64 * nt;
    }
    offs[F] = o;
  }
  if constexpr (LSM) {
    for (int i = tid; i < args3.linv_doubles; i += NTHR) lts[i] = 0.0;
  }
  stage_params_tma(vs, a.vhat_p, (unsigned)(Mp * DIMP * 8), &bar);   // includes a CTA barrier before the copy is issued
  __syncthreads();
  if constexpr (LSM) {   // one warp per row: its columns 0 .. min(8 i + 8, m_l) - 1
    for (int R = warp; R < a.M; R += FT3_WARPS) {
      int l = 0;
      while (R >= offs[l + 1]) ++l;
      const int off = offs[l], ml = offs[l + 1] - off, r = R - off, i = r >> 3;
      const int ncol = min(8 * i + 8, ml);
      double* dst = lts + lbase[l] + 32 * i * i + 64 * i + (r & 7) * (8 * i + 12);
      for (int k = lane; k < ncol; k += 32) dst[k] = args.Linv[(int64_t)R * Mp + off + k];
    }
    __syncthreads();
  }
  // from here on no CTA barrier

  double* zt = ztw + warp * 32 * 32;
  double* rw = rsw + warp * 32;
  const int ngrp = args3.ngrp;
  const int64_t tiles = (a.npts + 31) / 32;
  // unit u = tile * ngrp + group.  Round k of the grid covers units [k * stride, (k + 1) * stride); in it warp w takes position
  // (w + k * skew) mod stride, the skew chosen so that a warp's group index walks through ALL groups (they differ in size) even
  // when ngrp divides the stride; neighbouring warps keep neighbouring units (same tile: its coordinates come out of L1 / L2).
  const int stride = (int)gridDim.x * FT3_WARPS;
  int skew = 0;
  {
    auto gcd = [](int x, int y) { while (y) { const int t = x % y; x = y; y = t; } return x; };
    while (gcd((stride + skew) % ngrp, ngrp) != 1 && skew < ngrp) ++skew;   // ngrp = 1: gcd(0, 1) = 1
  }
  const int64_t units = tiles * ngrp;
  const bool vec_ok = (args.ldo & 1) == 0;
  // lane-constant parts of the swizzled tile addresses (ft_at with a row index that is a multiple of 4 plus t4 / 8 j + g)
  int bidx[4];   // B fragment of k-step k0: zt[k0 * 32 + bidx[nt]] = tile[k0 + t4][8 nt + g]
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) bidx[nt] = ft_at(t4, 8 * nt + g);
  const int sidx = g * FT_LD;   // C fragment store: row 8 j + g, columns (8 nt + 2 t4) ^ ((g & 3) << 2)
  const int sswz = (g & 3) << 2;

  int pos = (int)blockIdx.x * FT3_WARPS + warp;
  for (int64_t ubase = 0; ubase < units; ubase += stride) {
    const int64_t u = ubase + pos;
    pos += skew;
    if (pos >= stride) pos -= stride;
    if (u >= units) continue;   // only in the last round
    const int64_t tile = u / ngrp;
    const int gi = (int)(u - tile * ngrp);
    const int64_t p0 = tile * 32;
    const bool full = vec_ok && p0 + 32 <= a.npts;   // whole tile in range and 16-byte aligned rows: unguarded vector stores
    // ---- the tile's points onto the sphere (lane = point), B fragments of the projection through the private tile ----
    {
      const int64_t n = p0 + lane;
      double xh[DIMP];
#pragma unroll
      for (int c = 0; c < DIMP; ++c) xh[c] = 0.0;
      double r = 1.0;
      if (n < a.npts) {
        r = load_point<DIMP>(a.X, a.xcols, a.apply_to_sphere, a.sw, a.sb, a.dim, n, xh, a.proj);
        if (args.r_out && gi == 0) args.r_out[n] = r;
      }
#pragma unroll
      for (int c = 0; c < DIMP; c += 2) *reinterpret_cast<double2*>(zt + lane * DIMP + c) = make_double2(xh[c], xh[c + 1]);
      rw[lane] = args.mul_r ? r : 1.0;
    }
    __syncwarp();
    double xb[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) xb[ks][nt] = zt[(8 * nt + g) * DIMP + 4 * ks + t4];
    __syncwarp();

    for (int l = args3.grp_lo[gi]; l < (int)args3.grp_lo[gi + 1]; ++l) {
      const int off = offs[l], ml = offs[l + 1] - off;
      const int ntl = (ml + 7) >> 3;
      // ---- zonal values of degree l into the private tile (rows past m_l as zeros) ----
      const double ld = leads[l];
      for (int j = 0; j < ntl; ++j) {
        const int rloc = 8 * j + g;
        const double* vrow = vs + min(off + rloc, Mp - 1) * DIMP + t4;
        double t[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) t[q] = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const double av = vrow[4 * ks];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) dmma884(t[2 * nt], t[2 * nt + 1], av, xb[ks][nt]);
        }
        double z[8];
        if (l >= 1) {
          double p[8], p1[8], p2[8];
          monic_uniform<8>(betas, l, t, p, p1, p2);
          const double s = rloc < ml ? ld : 0.0;
#pragma unroll
          for (int q = 0; q < 8; ++q) z[q] = s * p[q];
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) z[q] = rloc < ml ? 1.0 : 0.0;
        }
        double* zr = zt + j * (8 * FT_LD) + sidx;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) *reinterpret_cast<double2*>(zr + ((8 * nt + 2 * t4) ^ sswz)) = make_double2(z[2 * nt], z[2 * nt + 1]);
      }
      __syncwarp();
      // ---- whitening: row tile i of the degree needs columns 0 .. 8 i + 7 of Linv_l (lower triangular) ----
      double sr[8];   // scale_l * r of this lane's eight points
      {
        const double sc = scl[l];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const double2 rr = *reinterpret_cast<const double2*>(rw + 8 * nt + 2 * t4);
          sr[2 * nt] = sc * rr.x;
          sr[2 * nt + 1] = sc * rr.y;
        }
      }
      for (int i = 0; i < ntl; ++i) {
        const int rloc = 8 * i + g;
        const bool rok = rloc < ml;
        const int kend = min(ml, 8 * i + 8);
        const double* Lrow = LSM ? lts + lbase[l] + 32 * i * i + 64 * i + g * (8 * i + 12) + t4
                                 : args.Linv + (int64_t)(off + (rok ? rloc : 0)) * Mp + off + t4;
        double acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = 0.0;
        // exactly ceil(kend / 4) k-steps through a real branch: as predicated code the skipped DMMAs of the triangle were still issued
        // (ncu: 1 720 instead of 1 212 DMMAs per tile at the headline shape, on a kernel whose top stall is the FP64 pipe)
        switch ((kend + 3) >> 2) {
          case 1: ft3_whiten_steps<1, LSM>(acc, Lrow, rok, ml, zt, bidx); break;
          case 2: ft3_whiten_steps<2, LSM>(acc, Lrow, rok, ml, zt, bidx); break;
          case 3: ft3_whiten_steps<3, LSM>(acc, Lrow, rok, ml, zt, bidx); break;
          case 4: ft3_whiten_steps<4, LSM>(acc, Lrow, rok, ml, zt, bidx); break;
          case 5: ft3_whiten_steps<5, LSM>(acc, Lrow, rok, ml, zt, bidx); break;
          case 6: ft3_whiten_steps<6, LSM>(acc, Lrow, rok, ml, zt, bidx); break;
          case 7: ft3_whiten_steps<7, LSM>(acc, Lrow, rok, ml, zt, bidx); break;
          default: ft3_whiten_steps<8, LSM>(acc, Lrow, rok, ml, zt, bidx); break;
        }
        if (rok) {
          double* orow = args.out + (int64_t)(off + rloc) * args.ldo + p0;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const int c = 8 * nt + 2 * t4;
            const double v0 = sr[2 * nt] * acc[2 * nt], v1 = sr[2 * nt + 1] * acc[2 * nt + 1];
            if (full) *reinterpret_cast<double2*>(orow + c) = make_double2(v0, v1);
            else {
              if (p0 + c < a.npts) orow[c] = v0;
              if (p0 + c + 1 < a.npts) orow[c + 1] = v1;
            }
          }
        }
      }
      __syncwarp();   // the tile is consumed before the next degree (or the next unit's coordinates) overwrites it
    }
  }
}

}  // namespace gpfy
