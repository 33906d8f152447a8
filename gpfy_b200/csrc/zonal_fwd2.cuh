// K_A  zonal forward on the FP64 tensor cores: to_sphere (spherical.py:214-249), t = Vhat xhat
// (spherical_harmonics.py:308) as DMMA with the points' coordinates held in registers as B fragments, the
// Gegenbauer recurrence (gegenbauer.py:47-60, monic form) on the accumulator fragments, mu2 . z
// (variational.py:305 folded) and, for the Gaussian likelihood, the per-point weights (likelihoods.py:210-218).
// One CTA = 32 points x all rows; warp w takes the 8-row tiles w, w + 8, ...
#pragma once
#include "gemm_tf32.cuh"
#include "svgp_kernels.cuh"
#include "zonal_common.cuh"

namespace gpfy {

struct ZonalFwd2Args {
  ZonalArgs z;
  MonicTab mt;  // parameter alpha
  int dbg;      // development: 1 = skip the Z stores
};

static inline int zf2_smem_bytes(int Mp, int dimp, int P, bool with_mz, int pmax) {
  int dbl = Mp * dimp + (with_mz ? P * Mp : 0) + 2 * 32 * dimp + 2 * 8 * pmax * 32 + 2 * 32 + 3 * GPFY_MAX_DEGREES + 10;
  return dbl * 8 + Mp * 4 + 16;
}

// Persistent: every CTA stages Vhat / mu2 once (TMA bulk copy) and then walks 32-point tiles
// blockIdx, blockIdx + gridDim, ...  Per tile there is ONE CTA barrier: warp 0 (helper) loads and projects the NEXT
// tile's points while the seven compute warps work on the current one (coordinates and the mu2 . z partial sums are double-buffered in
// shared memory), and finishes the previous tile's per-point outputs right after the barrier.
// ZPOUT: also store the derivative panel (blocked like the GEMM's C) for zonal_bwd4; the headline path leaves it off.
template <int DIMP, int PMAX, bool ZPOUT>
__global__ void __launch_bounds__(256, 2) zonal_fwd2_kernel(const __grid_constant__ ZonalFwd2Args args) {
  constexpr int KS = DIMP / 4;
  const ZonalArgs& a = args.z;
  extern __shared__ __align__(16) double smem[];
  __shared__ uint64_t bar;
  const int Mp = a.Mp;
  double* vs = smem;                                   // [Mp][DIMP]
  double* mu2s = vs + Mp * DIMP;                       // [P][Mp] (only when a.mz)
  double* xs = mu2s + (a.mz ? a.P * Mp : 0);           // [2][32][DIMP]
  double* red = xs + 2 * 32 * DIMP;                    // [2][8][PMAX][32]
  double* rs = red + 2 * 8 * PMAX * 32;                // [2][32] radii of the tile's points
  double* leads = rs + 2 * 32;                         // [MAX_DEGREES + 2]
  double* betas = leads + GPFY_MAX_DEGREES + 2;        // [MAX_DEGREES + 6] recurrence coefficients (read ahead)
  double* dcos = betas + GPFY_MAX_DEGREES + 6;         // [MAX_DEGREES + 2] derivative coefficients (MonicTab::dco)
  int* rowdeg = reinterpret_cast<int*>(dcos + GPFY_MAX_DEGREES + 2);  // [Mp]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int64_t npts_e = (a.npts + 1) & ~(int64_t)1;
  const int tiles = (int)((npts_e + 31) / 32);
  const int n_my = ((int)blockIdx.x < tiles) ? (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  for (int i = tid; i < Mp; i += 256) {
    int l = -1;
    if (i < a.M) {
      int o = 0;
      l = 0;
      while (i >= o + a.deg.m[l]) { o += a.deg.m[l]; ++l; }
    }
    rowdeg[i] = l;
  }
  for (int i = tid; i < GPFY_MAX_DEGREES + 2; i += 256) { leads[i] = args.mt.lead[i]; dcos[i] = args.mt.dco[i]; }
  for (int i = tid; i < GPFY_MAX_DEGREES + 6; i += 256) betas[i] = i < GPFY_MAX_DEGREES + 2 ? args.mt.beta[i] : 0.0;

  // warp 0, lane = point: onto the sphere (spherical.py:214-249); coordinates to shared memory, r and xhat to HBM
  auto project_tile = [&](int it) {
    const int64_t n = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * 32 + lane;
    double xh[DIMP];
#pragma unroll
    for (int c = 0; c < DIMP; ++c) xh[c] = 0.0;
    double r = 1.0;
    if (n < a.npts) {
      r = load_point<DIMP>(a.X, a.xcols, a.apply_to_sphere, a.sw, a.sb, a.dim, n, xh, a.proj);
      if (a.R) a.R[n] = r;
      if (a.XH) {
#pragma unroll
        for (int c = 0; c < DIMP; c += 2) *reinterpret_cast<double2*>(a.XH + n * DIMP + c) = make_double2(xh[c], xh[c + 1]);
      }
    }
    double* x = xs + (it & 1) * 32 * DIMP;
#pragma unroll
    for (int c = 0; c < DIMP; c += 2) *reinterpret_cast<double2*>(x + lane * DIMP + c) = make_double2(xh[c], xh[c + 1]);
    rs[(it & 1) * 32 + lane] = r;
  };
  // warp 0, lane = point: mz_p[n] = mu2_p . z_n summed over the warps in a fixed order (variational.py:305 folded),
  // and the Gaussian per-point weights (likelihoods.py:210-218)
  auto finish_tile = [&](int it) {
    const int64_t n = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * 32 + lane;
    const bool valid = n < a.npts;
    const double* rd = red + (it & 1) * 8 * PMAX * 32;
    const double r = rs[(it & 1) * 32 + lane];
#pragma unroll
    for (int pp = 0; pp < PMAX; ++pp)
      if (pp < a.P) {
        double s = 0.0;
#pragma unroll
        for (int w = 1; w < 8; ++w) s += rd[(w * PMAX + pp) * 32 + lane];
        if (valid) {
          a.mz[(int64_t)pp * a.ldz + n] = s;
          if (a.Yg) {
            const double inv = 1.0 / a.s2g[0];
            const double pw = (a.Yg[n] - r * s) * inv, qw = -0.5 * inv;
            a.wp[n] = pw * r;
            a.wq[n] = qw * r * r;
            a.cq[n] = 2.0 * qw * r * r;
          }
        }
      }
    if (!valid && n < npts_e && a.Yg) { a.wp[n] = 0.0; a.wq[n] = 0.0; a.cq[n] = 0.0; }
  };

  if (warp == 0 && n_my > 0) project_tile(0);
  stage_params_tma(vs, a.vhat_p, (unsigned)((Mp * DIMP + (a.mz ? a.P * Mp : 0)) * 8), &bar);
  __syncthreads();

  const int ntile = (ZPOUT || a.Zh) ? Mp / 8 : (a.M + 7) / 8;   // the derivative panel / float planes cover the padding rows too (zeros)
  for (int it = 0; it < n_my; ++it) {
    if (warp == 0) {  // helper warp: the next tile's points while the seven compute warps work, then this tile's outputs
      if (it + 1 < n_my) project_tile(it + 1);
      __syncthreads();
      if (a.mz) finish_tile(it);
      continue;
    }
    const int64_t p0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * 32;
    const double* x = xs + (it & 1) * 32 * DIMP;
    double xb[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) xb[ks][nt] = x[(8 * nt + g) * DIMP + 4 * ks + t4];

    double mzacc[PMAX][8];
#pragma unroll
    for (int p = 0; p < PMAX; ++p)
#pragma unroll
      for (int u = 0; u < 8; ++u) mzacc[p][u] = 0.0;

    for (int rt = warp - 1; rt < ntile; rt += 7) {  // seven compute warps: 35 row tiles at M = 280 = 5 each
      const int row = rt * 8 + g;
      double t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = 0.0;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const double av = vs[row * DIMP + 4 * ks + t4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma884(t[2 * nt], t[2 * nt + 1], av, xb[ks][nt]);
      }
      const int l = rowdeg[row];
      const int lo = rowdeg[rt * 8], hi = rowdeg[rt * 8 + 7];
      double p[8], z[8];
      if constexpr (ZPOUT) {
        double zp[8];
        if (lo == hi && lo >= 1) {
          monic_uniform_d<8>(betas, dcos, lo, t, p, zp);
          const double ld = leads[lo];
#pragma unroll
          for (int u = 0; u < 8; ++u) z[u] = ld * p[u];
        } else {
          const int lmax = __reduce_max_sync(0xffffffffu, l);
          monic_mixed_d<8>(betas, dcos, l, lmax, t, p, zp);
          const double ld = l >= 1 ? leads[l] : 0.0;
#pragma unroll
          for (int u = 0; u < 8; ++u) z[u] = l >= 1 ? ld * p[u] : (l == 0 ? 1.0 : 0.0);
        }
        double* zrow = a.ZPb + ((p0 >> 5) * (int64_t)Mp + row) * 32;
        const int swz = (row & 3) << 2;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          *reinterpret_cast<double2*>(zrow + ((8 * nt + 2 * t4) ^ swz)) = make_double2(zp[2 * nt], zp[2 * nt + 1]);
      } else {
        double p1[8], p2[8];
        if (lo == hi && lo >= 1) {
          monic_uniform<8>(betas, lo, t, p, p1, p2);
          const double ld = leads[lo];
#pragma unroll
          for (int u = 0; u < 8; ++u) z[u] = ld * p[u];
        } else {
          const int lmax = __reduce_max_sync(0xffffffffu, l);
          monic_mixed<8>(betas, l, lmax, t, p, p1, p2);
          const double ld = l >= 1 ? leads[l] : 0.0;
#pragma unroll
          for (int u = 0; u < 8; ++u) z[u] = l >= 1 ? ld * p[u] : (l == 0 ? 1.0 : 0.0);
        }
      }
      if (a.Zh) {   // the same values as two TF32 planes (hi + lo), tile-contiguous [n / 32][Mp][32], for the tcgen05 contractions;
                    // padding rows (z = 0) and the columns past the end of the chunk are written as zeros
        float* zhrow = a.Zh + ((p0 >> 5) * (int64_t)Mp + row) * 32;
        float* zlrow = a.Zl + ((p0 >> 5) * (int64_t)Mp + row) * 32;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int64_t n0 = p0 + 8 * nt + 2 * t4;
          float h0, l0, h1, l1;
          tf32_split(row < a.M ? z[2 * nt] : 0.0, h0, l0);
          tf32_split(row < a.M ? z[2 * nt + 1] : 0.0, h1, l1);
          if (n0 + 1 >= a.npts) { h1 = 0.f; l1 = 0.f; }
          if (n0 >= a.npts) { h0 = 0.f; l0 = 0.f; }
          *reinterpret_cast<float2*>(zhrow + 8 * nt + 2 * t4) = make_float2(h0, h1);
          *reinterpret_cast<float2*>(zlrow + 8 * nt + 2 * t4) = make_float2(l0, l1);
        }
      }
      if (row < a.M) {
        if (!(args.dbg & 1) && a.Z)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int64_t n0 = p0 + 8 * nt + 2 * t4;
          double* dst = a.Z + (int64_t)row * a.ldz + n0;
          if (n0 + 1 < a.npts) *reinterpret_cast<double2*>(dst) = make_double2(z[2 * nt], z[2 * nt + 1]);
          else if (n0 < a.npts) {
            dst[0] = z[2 * nt];
            if (n0 + 1 < npts_e) dst[1] = 0.0;  // keep the even-padded tail column finite
          }
        }
        if (a.mz) {
#pragma unroll
          for (int pp = 0; pp < PMAX; ++pp)
            if (pp < a.P) {
              const double m2 = mu2s[pp * Mp + row];
#pragma unroll
              for (int u = 0; u < 8; ++u) mzacc[pp][u] = fma(m2, z[u], mzacc[pp][u]);
            }
        }
      }
    }
    if (a.mz) {  // over the 8 row lanes by shuffles; the warps' parts go to shared memory
      double* rd = red + (it & 1) * 8 * PMAX * 32;
#pragma unroll
      for (int pp = 0; pp < PMAX; ++pp)
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          double v = mzacc[pp][u];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (g == 0) rd[(warp * PMAX + pp) * 32 + 8 * (u >> 1) + 2 * t4 + (u & 1)] = v;
        }
    }
    __syncthreads();
  }
}

}  // namespace gpfy

namespace gpfy {

// ------------------------------------------------------------------------------------------------
// K_F  fused feature matrix (spherical_harmonics.py:290-320, :344-362, :385-387):
//     out[off_l + i, n] = scale_l * (r_n) * sum_k Linv_l[i][k] C_l^alpha(Vhat_l xhat_n)[k]
// One persistent CTA walks 32-point tiles.  Phase A is the tensor-core zonal forward above, but the zonal values go
// to a swizzled shared-memory tile instead of HBM; phase B is the per-degree whitening as a block-triangular DMMA
// product (row tile i of degree l only needs columns 0 .. 8 i + 7 of Linv_l), scaled and stored straight into the
// [M, N] output.  Z never touches HBM: the kernel writes 8 M bytes per point and reads 8 d.
// ------------------------------------------------------------------------------------------------
struct FeaturesArgs {
  ZonalArgs z;            // Z / ldz unused
  MonicTab mt;
  const double* Linv;     // dense block-diagonal [Mp][Mp] (prologue)
  const double* scale;    // [F] or null
  int mul_r;
  double* out;            // [M][ldo], first column of this launch
  int64_t ldo;
  double* r_out;          // [npts] or null
  unsigned char item_warp[160];   // phase B: warp of every (degree, 8-row tile) item, dealt on the host by falling cost (255 = round robin)
};

// host: longest-processing-time deal of the whitening items (cost = k-steps of the lower-triangular row tile) to the 8 warps
static inline void ft_deal_items(FeaturesArgs* fa, const DegreeTable& deg) {
  int cost[160], order[160], n = 0;
  for (int l = 0; l < deg.F; ++l)
    for (int i = 0; i * 8 < deg.m[l]; ++i) {
      if (n >= 160) { for (int k = 0; k < 160; ++k) fa->item_warp[k] = 255; return; }
      const int kend = deg.m[l] < 8 * i + 8 ? deg.m[l] : 8 * i + 8;
      cost[n] = (kend + 3) / 4 + 2;   // + the store
      order[n] = n;
      ++n;
    }
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j)
      if (cost[order[j]] > cost[order[i]]) { const int t = order[i]; order[i] = order[j]; order[j] = t; }
  int load[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = 0; k < 160; ++k) fa->item_warp[k] = 255;
  for (int i = 0; i < n; ++i) {
    int best = 0;
    for (int w = 1; w < 8; ++w) if (load[w] < load[best]) best = w;
    fa->item_warp[order[i]] = (unsigned char)best;
    load[best] += cost[order[i]];
  }
}

constexpr int FT_LD = 32;
__device__ __forceinline__ int ft_at(int row, int pt) { return row * FT_LD + (pt ^ ((row & 3) << 2)); }

static inline int ft_smem_bytes(int Mp, int dimp) {
  int dbl = Mp * dimp + Mp * FT_LD + 2 * 32 * dimp + 2 * 32 + 3 * (GPFY_MAX_DEGREES + 2) + 4;
  return dbl * 8 + Mp * 4 + (GPFY_MAX_DEGREES + 1) * 4 + 16;
}

template <int DIMP>
__global__ void __launch_bounds__(256, 2) features2_kernel(const __grid_constant__ FeaturesArgs args) {
  constexpr int KS = DIMP / 4;
  const ZonalArgs& a = args.z;
  extern __shared__ __align__(16) double smem[];
  __shared__ uint64_t bar;
  const int Mp = a.Mp;
  double* vs = smem;                          // [Mp][DIMP]
  double* zt = vs + Mp * DIMP;                // [Mp][32] swizzled zonal tile
  double* xs = zt + Mp * FT_LD;               // [2][32][DIMP]
  double* rs = xs + 2 * 32 * DIMP;            // [2][32]
  double* leads = rs + 2 * 32;                // [MAX_DEGREES + 2]
  double* scl = leads + GPFY_MAX_DEGREES + 2; // [MAX_DEGREES + 2] per-degree output scale
  double* betas = scl + GPFY_MAX_DEGREES + 2; // [MAX_DEGREES + 6]
  int* rowdeg = reinterpret_cast<int*>(betas + GPFY_MAX_DEGREES + 6);  // [Mp]
  int* offs = rowdeg + Mp;                    // [F + 1]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int tiles = (int)((a.npts + 31) / 32);
  const int n_my = ((int)blockIdx.x < tiles) ? (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int F = a.deg.F;

  for (int i = tid; i < Mp; i += 256) {
    int l = -1;
    if (i < a.M) {
      int o = 0;
      l = 0;
      while (i >= o + a.deg.m[l]) { o += a.deg.m[l]; ++l; }
    }
    rowdeg[i] = l;
  }
  for (int i = tid; i < Mp * FT_LD; i += 256) zt[i] = 0.0;  // padding rows stay zero
  for (int i = tid; i < GPFY_MAX_DEGREES + 2; i += 256) {
    leads[i] = args.mt.lead[i];
    scl[i] = (i < F && args.scale) ? args.scale[i] : 1.0;
  }
  for (int i = tid; i < GPFY_MAX_DEGREES + 6; i += 256) betas[i] = i < GPFY_MAX_DEGREES + 2 ? args.mt.beta[i] : 0.0;
  if (tid == 0) {
    int o = 0;
    for (int l = 0; l < F; ++l) { offs[l] = o; o += a.deg.m[l]; }
    offs[F] = o;
  }

  auto project_tile = [&](int it) {
    const int64_t n = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * 32 + lane;
    double xh[DIMP];
#pragma unroll
    for (int c = 0; c < DIMP; ++c) xh[c] = 0.0;
    double r = 1.0;
    if (n < a.npts) {
      r = load_point<DIMP>(a.X, a.xcols, a.apply_to_sphere, a.sw, a.sb, a.dim, n, xh, a.proj);
      if (args.r_out) args.r_out[n] = r;
    }
    double* x = xs + (it & 1) * 32 * DIMP;
#pragma unroll
    for (int c = 0; c < DIMP; c += 2) *reinterpret_cast<double2*>(x + lane * DIMP + c) = make_double2(xh[c], xh[c + 1]);
    rs[(it & 1) * 32 + lane] = args.mul_r ? r : 1.0;
  };

  if (warp == 0 && n_my > 0) project_tile(0);
  stage_params_tma(vs, a.vhat_p, (unsigned)(Mp * DIMP * 8), &bar);
  __syncthreads();

  const int ntile = (a.M + 7) / 8;
  const bool vec_ok = (args.ldo & 1) == 0;
  for (int it = 0; it < n_my; ++it) {
    const int64_t p0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * 32;
    const double* x = xs + (it & 1) * 32 * DIMP;
    // ---- phase A: zonal values of the tile into shared memory (seven compute warps; warp 0 projects the next
    // tile's points meanwhile) ----
    if (warp == 0) {
      if (it + 1 < n_my) project_tile(it + 1);
    } else {
      double xb[KS][4];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) xb[ks][nt] = x[(8 * nt + g) * DIMP + 4 * ks + t4];
      for (int rt = warp - 1; rt < ntile; rt += 7) {
        const int row = rt * 8 + g;
        double t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const double av = vs[row * DIMP + 4 * ks + t4];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) dmma884(t[2 * nt], t[2 * nt + 1], av, xb[ks][nt]);
        }
        const int l = rowdeg[row];
        const int lo = rowdeg[rt * 8], hi = rowdeg[rt * 8 + 7];
        double p[8], p1[8], p2[8], z[8];
        if (lo == hi && lo >= 1) {
          monic_uniform<8>(betas, lo, t, p, p1, p2);
          const double ld = leads[lo];
#pragma unroll
          for (int u = 0; u < 8; ++u) z[u] = ld * p[u];
        } else {
          const int lmax = __reduce_max_sync(0xffffffffu, l);
          monic_mixed<8>(betas, l, lmax, t, p, p1, p2);
          const double ld = l >= 1 ? leads[l] : 0.0;
#pragma unroll
          for (int u = 0; u < 8; ++u) z[u] = l >= 1 ? ld * p[u] : (l == 0 ? 1.0 : 0.0);
        }
        if (row < a.M) {
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) *reinterpret_cast<double2*>(zt + ft_at(row, 8 * nt + 2 * t4)) = make_double2(z[2 * nt], z[2 * nt + 1]);
        }
      }
    }
    __syncthreads();

    // ---- phase B: whitening, work item = (degree, 8-row tile of that degree) dealt round-robin to the warps ----
    {
      double rr[8];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        rr[2 * nt] = rs[(it & 1) * 32 + 8 * nt + 2 * t4];
        rr[2 * nt + 1] = rs[(it & 1) * 32 + 8 * nt + 2 * t4 + 1];
      }
      int item = 0;
      for (int l = 0; l < F; ++l) {
        const int off = offs[l], ml = offs[l + 1] - off;
        const double sc = scl[l];
        for (int i = 0; i * 8 < ml; ++i, ++item) {
          const int iw = item < 160 ? (int)args.item_warp[item] : 255;
          if ((iw == 255 ? (item & 7) : iw) != warp) continue;
          double acc[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) acc[u] = 0.0;
          const int rloc = 8 * i + g;                 // row inside the degree block
          const int kend = min(ml, 8 * i + 8);        // lower triangular: columns 0 .. 8 i + 7
          const bool rok = rloc < ml;
          const double* Lrow = args.Linv + (int64_t)(off + (rok ? rloc : 0)) * Mp + off;
          // the A fragments (rows of Linv, read through L1) of eight k-steps are fetched ahead of the tensor-core steps that use
          // them: one exposed load latency per 32 columns instead of one per step (ncu r02a: the DMMAs' long-scoreboard stalls)
          for (int kb0 = 0; kb0 < kend; kb0 += 32) {
            double av[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int kc = kb0 + 4 * q + t4;
              av[q] = (rok && kc < ml && kb0 + 4 * q < kend) ? __ldg(Lrow + kc) : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int k0 = kb0 + 4 * q;
              if (k0 < kend) {
                const int zrow = min(off + k0 + t4, Mp - 1);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) dmma884(acc[2 * nt], acc[2 * nt + 1], av[q], zt[ft_at(zrow, 8 * nt + g)]);
              }
            }
          }
          if (rok) {
            double* orow = args.out + (int64_t)(off + rloc) * args.ldo + p0;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              const int c = 8 * nt + 2 * t4;
              const double v0 = sc * rr[2 * nt] * acc[2 * nt], v1 = sc * rr[2 * nt + 1] * acc[2 * nt + 1];
              if (p0 + c + 1 < a.npts && vec_ok) *reinterpret_cast<double2*>(orow + c) = make_double2(v0, v1);
              else {
                if (p0 + c < a.npts) orow[c] = v0;
                if (p0 + c + 1 < a.npts) orow[c + 1] = v1;
              }
            }
          }
        }
      }
    }
    __syncthreads();  // the tile is consumed before the next phase A overwrites it
  }
}

}  // namespace gpfy
