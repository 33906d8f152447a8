// K_B  fused backward through the zonal map for P = 1 and shapes whose [Mp x 32] tile fits in shared memory
// (the headline shapes: Mp <= 288, ambient dim <= 16).
//
// One persistent CTA per SM (12 warps) walks 32-point tiles of the chunk.  Per tile:
//   load   : the [Mp x 32] slice of C = W2 Z plus the tile's xhat rows and per-point weights arrive in shared
//            memory by a handful of TMA bulk copies (cp.async.bulk + mbarrier complete_tx; the GEMM stores C tile by
//            tile, pre-swizzled: GemmArgs::c_blocked, so a tile is one contiguous slab), two tiles in flight
//   phase 1: lane = point; every warp owns three 8-row tiles (dealt on the host so that the recurrence lengths
//            balance): t = Vhat xhat, then the monic (alpha + 1) recurrence gives BOTH
//            d/dt C_l^alpha = 2 alpha C_{l-1}^{alpha+1} (gegenbauer.py:73-81) and, through
//            C_l^alpha = alpha / (l + alpha) (C_l^{alpha+1} - C_{l-2}^{alpha+1}), the zonal value itself, so
//            psi = z . C (variational.py:326-327 folded) costs no extra pass;  dt = dz * zp overwrites C in place
//   phase 2: FP64 tensor cores on the warp's own rows of the dT tile:
//            dVhat[rows, dim] += dT[rows, 32] Xhat[32, dim]      (VJP of spherical_harmonics.py:308; accumulators
//                                                                 persist in registers over all tiles of the CTA)
//            dxhat[dim, 32]   += Vhat^T[dim, rows] dT[rows, 32]   (per-warp partial, parked in the warp's dead dT rows)
//   tail   : ONE warp (rotating) waits for the other eleven on an mbarrier, adds the partials, does the
//            likelihood terms (Gaussian, fused) and du -> dw / db sums, then refills the buffer for tile + 2.
// There is no CTA-wide barrier in the loop: warps drift by up to two tiles, so the FP64 pipe stays fed while
// some warps run DFMA recurrences and others DMMA.  dT never goes to HBM; the separate dVhat and likelihood
// kernels and the GEMM's psi epilogue disappear.
#pragma once
#include "svgp_kernels.cuh"

namespace gpfy {

constexpr int ZB_THREADS = 384, ZB_WARPS = 12, ZB_LD = 32, ZB_SUMS = 40, ZB_SLOTS = 3;
// element (row, point) of a [Mp][32] tile as the GEMM stores it (GemmArgs::c_blocked): dense rows, the column
// XOR-swizzled by the row so that per-row accesses and 8 x 4 / 4 x 8 fragment reads are conflict-free
__device__ __forceinline__ int zb_at(int row, int pt) { return row * ZB_LD + (pt ^ ((row & 3) << 2)); }

struct ZonalBwd3Args {
  const double* C;       // [tiles][Mp][32] blocked + swizzled (GemmArgs::c_blocked)
  int64_t ldz;
  const double* wp;      // [ldz]  p r
  const double* cq;      // [ldz]  2 q r^2
  const double* dr;      // [ldz]  direct d/dr terms (non-fused likelihood only)
  const double* XH;      // [ldz][DIMP]
  const double* R;       // [npts]
  const double* mz;      // [ldz]  mu2 . z       (fused likelihood)
  const double* Y;       // [npts]               (fused likelihood)
  const double* vptr;    // [1] kernel variance  (fused likelihood)
  const double* s2ptr;   // [1] noise variance   (fused likelihood)
  int order;
  const double* vhat_p;  // [Mp][DIMP], mu2 [Mp] right behind it
  int Mp, M, d;
  int64_t npts;
  double* partials;      // [gridDim][npart]
  int npart;
  double* dVpart;        // [gridDim][Mp][CT * 8]
  int dbg;               // development: 1 = skip phases 1/2 (load pipeline only)
  int red_sep;           // 1: dxhat partials in their own buffer (small Mp), 0: parked in the warp's dead dT rows
  unsigned char rt[ZB_WARPS * ZB_SLOTS];  // row tile (8 rows) of each (warp, slot); 255 = none
  DegreeTable deg;
  // monic (alpha + 1) recurrence p_k = t p_{k-1} - beta[k] p_{k-2};  C_k^{alpha+1} = lead[k] p_k
  double beta[GPFY_MAX_DEGREES + 2];
  double ca[GPFY_MAX_DEGREES + 2];   // zp = ca[l] p_{l-1}            (2 alpha lead[l-1])
  double cb1[GPFY_MAX_DEGREES + 2];  // z  = cb1[l] p_l - cb2[l] p_{l-2}
  double cb2[GPFY_MAX_DEGREES + 2];
};

static inline int zb3_smem_bytes(int Mp, int dimp, int red_sep) {
  const int hdr = 64 + 32 * dimp;
  int dbl = Mp * dimp + Mp + 2 * Mp * ZB_LD + 2 * hdr + 2 * ZB_WARPS * 32 + ZB_SUMS * 32;
  if (red_sep) dbl += 2 * ZB_WARPS * dimp * 32;
  return dbl * 8 + Mp * 4 + 2 * ZB_WARPS * 4 + 64;
}

// host: recurrence constants and the balanced deal of row tiles to warps
static inline void zb3_setup(ZonalBwd3Args* a, double alpha, const DegreeTable& deg, int M, int Mp) {
  const double lam = alpha + 1.0;
  double lead[GPFY_MAX_DEGREES + 2];
  lead[0] = 1.0;
  lead[1] = 2.0 * lam;
  a->beta[0] = a->beta[1] = 0.0;
  for (int k = 2; k < GPFY_MAX_DEGREES + 2; ++k) {
    const double d1 = 2.0 * (k + lam - 1.0) / k, d2 = (k + 2.0 * lam - 2.0) / k;
    lead[k] = d1 * lead[k - 1];
    a->beta[k] = d2 * lead[k - 2] / lead[k];
  }
  for (int l = 0; l < GPFY_MAX_DEGREES + 2; ++l) {
    const double zr = (l == 0 && alpha == 0.0) ? 0.0 : alpha / (l + alpha);
    a->ca[l] = l >= 1 ? 2.0 * alpha * lead[l - 1] : 0.0;
    a->cb1[l] = zr * lead[l];
    a->cb2[l] = l >= 2 ? zr * lead[l - 2] : 0.0;
  }
  // cost of a row tile = sum over its rows of (fixed work + recurrence length); longest-processing-time deal
  const int nt = Mp / 8;
  int cost[64], order[64], load[ZB_WARPS], cnt[ZB_WARPS];
  for (int t = 0; t < nt; ++t) {
    int c = 0;
    for (int r = 8 * t; r < 8 * t + 8; ++r) {
      int l = -1;
      if (r < M) { int o = 0; l = 0; while (r >= o + deg.m[l]) { o += deg.m[l]; ++l; } }
      c += l < 0 ? 4 : 18 + 2 * l;
    }
    cost[t] = c;
    order[t] = t;
  }
  for (int i = 0; i < nt; ++i)
    for (int j = i + 1; j < nt; ++j)
      if (cost[order[j]] > cost[order[i]]) { int t = order[i]; order[i] = order[j]; order[j] = t; }
  for (int w = 0; w < ZB_WARPS; ++w) { load[w] = 0; cnt[w] = 0; }
  for (int i = 0; i < ZB_WARPS * ZB_SLOTS; ++i) a->rt[i] = 255;
  for (int i = 0; i < nt; ++i) {
    int best = -1;
    for (int w = 0; w < ZB_WARPS; ++w)
      if (cnt[w] < ZB_SLOTS && (best < 0 || cnt[w] < cnt[best] || (cnt[w] == cnt[best] && load[w] < load[best]))) best = w;
    a->rt[best * ZB_SLOTS + cnt[best]++] = (unsigned char)order[i];
    load[best] += cost[order[i]];
  }
}

#ifdef GPFY_ZB_CLOCKS
__device__ unsigned long long g_zb_clk[12 * 32 * 8];
#define ZB_CLK(i) do { if (blockIdx.x == 0 && lane == 0 && it < 32) g_zb_clk[(it * 12 + warp) * 8 + (i)] = clock64(); } while (0)
#else
#define ZB_CLK(i) do { } while (0)
#endif

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// four interleaved monic recurrences of degree L: p (= p_L), p1 (= p_{L-1}), p2 (= p_{L-2})
template <int L>
__device__ __forceinline__ void monic4(const double* beta, const double (&t)[4], double (&p)[4], double (&p1)[4], double (&p2)[4]) {
#pragma unroll
  for (int u = 0; u < 4; ++u) { p2[u] = 0.0; p1[u] = 1.0; p[u] = t[u]; }  // p_{-1}, p_0, p_1
#pragma unroll
  for (int k = 2; k <= L; ++k) {
    const double b = beta[k];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double pn = fma(t[u], p[u], -b * p1[u]);
      p2[u] = p1[u];
      p1[u] = p[u];
      p[u] = pn;
    }
  }
}

template <int DIMP, bool FUSE_LIK>
__global__ void __launch_bounds__(ZB_THREADS, 1) zonal_bwd3_kernel(const __grid_constant__ ZonalBwd3Args a) {
  constexpr int CT = (DIMP + 7) / 8;
  constexpr int HDR = 64 + 32 * DIMP;
  extern __shared__ __align__(16) double smem[];
  __shared__ uint64_t bar_stage, bar_full[2], bar_done[2];
  const int Mp = a.Mp;
  double* vs = smem;                          // [Mp][DIMP]
  double* mu2s = vs + Mp * DIMP;              // [Mp]
  double* cbuf = mu2s + Mp;                   // [2][Mp][32] swizzled (zb_at)
  double* hdr = cbuf + 2 * Mp * ZB_LD;        // [2][wp(32) | cq(32) | xhat(32 x DIMP)]
  double* psir = hdr + 2 * HDR;               // [2][12][32]
  double* lsum = psir + 2 * ZB_WARPS * 32;    // [40][32] per-lane running sums of the tail
  double* redsep = lsum + ZB_SUMS * 32;       // [2][12][DIMP][32] when a.red_sep
  int* rowdeg = reinterpret_cast<int*>(redsep + (a.red_sep ? 2 * ZB_WARPS * DIMP * 32 : 0));  // [Mp]
  int* prow = rowdeg + Mp;                    // [12][2] first row of each warp's two parking row tiles (-1: none)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int tiles = (int)((a.npts + 31) / 32);
  const int n_my = ((int)blockIdx.x < tiles) ? (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  // ---- one-time setup ----
  for (int i = tid; i < Mp; i += ZB_THREADS) {
    int l = -1;
    if (i < a.M) {
      int o = 0;
      l = 0;
      while (i >= o + a.deg.m[l]) { o += a.deg.m[l]; ++l; }
    }
    rowdeg[i] = l;
  }
  for (int i = tid; i < ZB_SUMS * 32; i += ZB_THREADS) lsum[i] = 0.0;
  if (tid < 2 * ZB_WARPS) {
    const unsigned char t = a.rt[(tid >> 1) * ZB_SLOTS + (tid & 1)];
    prow[tid] = t == 255 ? -1 : (int)t * 8;
  }
  if (tid == 0) {
    mbar_init(&bar_full[0], 1);
    mbar_init(&bar_full[1], 1);
    mbar_init(&bar_done[0], ZB_WARPS);
    mbar_init(&bar_done[1], ZB_WARPS);
  }
  stage_params_tma(vs, a.vhat_p, (unsigned)((Mp * DIMP + Mp) * 8), &bar_stage);  // includes a __syncthreads

  // producer (one warp): Mp row copies of 256 bytes + wp, cq, xhat rows, all completing on bar_full[buf]
  auto issue_tile = [&](int it) {
    const int buf = it & 1;
    const int64_t p0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * 32;
    double* cb = cbuf + buf * Mp * ZB_LD;
    double* hb = hdr + buf * HDR;
    // The tile's [Mp][32] slab is contiguous in C: a few TMA bulk copies (<= 32 KB each) + the header pieces, all
    // completing on bar_full[buf].  Issuing them costs the producing (tail) warp a handful of instructions, which
    // matters because the refill sits on the per-tile critical path (150 LDGSTS per lane cost ~5 k cycles).
    if (lane == 0) mbar_expect_tx(&bar_full[buf], (unsigned)(Mp * 256 + HDR * 8));
    __syncwarp();
    const unsigned slab = (unsigned)Mp * 256u;
    const char* src = reinterpret_cast<const char*>(a.C + (p0 >> 5) * (int64_t)Mp * 32);
    for (unsigned off = (unsigned)lane * 16384u; off < slab; off += 32u * 16384u)
      tma_bulk_g2s(reinterpret_cast<char*>(cb) + off, src + off, min(16384u, slab - off), &bar_full[buf]);
    if (lane == 29) tma_bulk_g2s(hb, a.wp + p0, 256, &bar_full[buf]);
    if (lane == 30) tma_bulk_g2s(hb + 32, a.cq + p0, 256, &bar_full[buf]);
    if (lane == 31) tma_bulk_g2s(hb + 64, a.XH + p0 * DIMP, 32 * DIMP * 8, &bar_full[buf]);
  };
  if (warp == 0) {
    if (n_my > 0) issue_tile(0);
    if (n_my > 1) issue_tile(1);
  }

  // this warp's row tiles
  int myrt[ZB_SLOTS];
#pragma unroll
  for (int r = 0; r < ZB_SLOTS; ++r) myrt[r] = a.rt[warp * ZB_SLOTS + r] == 255 ? -1 : (int)a.rt[warp * ZB_SLOTS + r];

  double accv[ZB_SLOTS][CT][2];
#pragma unroll
  for (int r = 0; r < ZB_SLOTS; ++r)
#pragma unroll
    for (int c = 0; c < CT; ++c) accv[r][c][0] = accv[r][c][1] = 0.0;

  const double s2 = FUSE_LIK ? a.s2ptr[0] : 1.0, s2inv = 1.0 / s2;
  const double e_const = -0.5 * 1.8378770664093454835606594728112 - 0.5 * log(s2);
  const double kv = FUSE_LIK ? a.vptr[0] : 1.0;

  for (int it = 0; it < n_my; ++it) {
    const int buf = it & 1;
    const unsigned par = (it >> 1) & 1;
    const int64_t p0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * 32;
    const int64_t n = p0 + lane;
    const bool valid = n < a.npts;
    const int nvalid = (int)((a.npts - p0) < 32 ? (a.npts - p0) : 32);
    const int tw = it % ZB_WARPS;  // tail warp of this tile
    double* cb = cbuf + buf * Mp * ZB_LD;
    const double* hb = hdr + buf * HDR;
    double t_r = 1.0, t_mz = 0.0, t_y = 0.0, t_dr = 0.0;  // the tail warp fetches its per-point scalars early
    if (warp == tw && valid) {
      t_r = a.R[n];
      if (FUSE_LIK) { t_mz = a.mz[n]; t_y = a.Y[n]; }
      else t_dr = a.dr[n];
    }
    ZB_CLK(0);
    mbar_wait(&bar_full[buf], par);
    ZB_CLK(1);

    // ---- phase 1: own rows, lane = point.  Points past the end of the chunk read whatever the buffers hold:
    // every per-point input is masked so that their dT is exactly zero ----
    double xh[DIMP];
#pragma unroll
    for (int c = 0; c < DIMP; c += 2) {
      const double2 v2 = *reinterpret_cast<const double2*>(hb + 64 + lane * DIMP + c);
      xh[c] = valid ? v2.x : 0.0;
      xh[c + 1] = valid ? v2.y : 0.0;
    }
    const double cpw = valid ? hb[lane] : 0.0, cqw = valid ? hb[32 + lane] : 0.0;
    double psi = 0.0;
#pragma unroll 1
    for (int s = 0; s < ZB_SLOTS; ++s) {
      if (myrt[s] < 0 || (a.dbg & 1)) continue;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int r0 = myrt[s] * 8 + 4 * h;
        double cval[4], tt[4], z[4], zp[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double cv = cb[zb_at(r0 + u, lane)];
          cval[u] = valid ? cv : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double* v = vs + (r0 + u) * DIMP;
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int c = 0; c < DIMP; c += 2) {
            const double2 vv = *reinterpret_cast<const double2*>(v + c);
            s0 = fma(vv.x, xh[c], s0);
            s1 = fma(vv.y, xh[c + 1], s1);
          }
          tt[u] = s0 + s1;
        }
        const int l0 = rowdeg[r0], l3 = rowdeg[r0 + 3];
        if (l0 == l3 && l0 >= 1) {  // four rows of one degree (the common case)
          double p[4], p1[4], p2[4];
          switch (l0) {
            case 1: monic4<1>(a.beta, tt, p, p1, p2); break;
            case 2: monic4<2>(a.beta, tt, p, p1, p2); break;
            case 3: monic4<3>(a.beta, tt, p, p1, p2); break;
            case 4: monic4<4>(a.beta, tt, p, p1, p2); break;
            case 5: monic4<5>(a.beta, tt, p, p1, p2); break;
            case 6: monic4<6>(a.beta, tt, p, p1, p2); break;
            case 7: monic4<7>(a.beta, tt, p, p1, p2); break;
            case 8: monic4<8>(a.beta, tt, p, p1, p2); break;
            case 9: monic4<9>(a.beta, tt, p, p1, p2); break;
            case 10: monic4<10>(a.beta, tt, p, p1, p2); break;
            case 11: monic4<11>(a.beta, tt, p, p1, p2); break;
            case 12: monic4<12>(a.beta, tt, p, p1, p2); break;
            default: {
#pragma unroll
              for (int u = 0; u < 4; ++u) { p2[u] = 0.0; p1[u] = 1.0; p[u] = tt[u]; }
              for (int k = 2; k <= l0; ++k) {
                const double b = a.beta[k];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const double pn = fma(tt[u], p[u], -b * p1[u]);
                  p2[u] = p1[u]; p1[u] = p[u]; p[u] = pn;
                }
              }
            }
          }
          const double ka = a.ca[l0], k1 = a.cb1[l0], k2 = a.cb2[l0];
#pragma unroll
          for (int u = 0; u < 4; ++u) { zp[u] = ka * p1[u]; z[u] = fma(k1, p[u], -k2 * p2[u]); }
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int l = rowdeg[r0 + u];
            if (l <= 0) { z[u] = l == 0 ? 1.0 : 0.0; zp[u] = 0.0; }
            else {
              double q2 = 0.0, q1 = 1.0, q = tt[u];
              for (int k = 2; k <= l; ++k) { const double qn = fma(tt[u], q, -a.beta[k] * q1); q2 = q1; q1 = q; q = qn; }
              zp[u] = a.ca[l] * q1;
              z[u] = fma(a.cb1[l], q, -a.cb2[l] * q2);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double dz = fma(cqw, cval[u], cpw * mu2s[r0 + u]);
          cb[zb_at(r0 + u, lane)] = dz * zp[u];
          psi = fma(z[u], cval[u], psi);
        }
      }
    }
    __syncwarp();
    ZB_CLK(2);

    // ---- phase 2a: dVhat[own rows] += dT Xhat ----
    if (!(a.dbg & 2))
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
      double af[ZB_SLOTS], bf[CT];
      const int pt = 4 * k4 + t4;
#pragma unroll
      for (int r = 0; r < ZB_SLOTS; ++r) af[r] = myrt[r] >= 0 ? cb[zb_at(myrt[r] * 8 + g, pt)] : 0.0;
#pragma unroll
      for (int c = 0; c < CT; ++c) bf[c] = (8 * c + g < DIMP && pt < nvalid) ? hb[64 + pt * DIMP + 8 * c + g] : 0.0;
#pragma unroll
      for (int r = 0; r < ZB_SLOTS; ++r)
#pragma unroll
        for (int c = 0; c < CT; ++c) dmma884(accv[r][c][0], accv[r][c][1], af[r], bf[c]);
    }
    // ---- phase 2b: dxhat partial over own rows: [dim, 32 points] = Vhat^T dT ----
    double acc2[CT][4][2];
#pragma unroll
    for (int mt = 0; mt < CT; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc2[mt][nt][0] = acc2[mt][nt][1] = 0.0;
#pragma unroll 1
    for (int s = 0; s < ZB_SLOTS; ++s) {
      if (myrt[s] < 0 || (a.dbg & 2)) continue;
#pragma unroll
      for (int k4 = 0; k4 < 2; ++k4) {
        const int row = myrt[s] * 8 + 4 * k4 + t4;
        double bf[4], av[CT];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) bf[nt] = cb[zb_at(row, 8 * nt + g)];
#pragma unroll
        for (int mt = 0; mt < CT; ++mt) av[mt] = (8 * mt + g < DIMP) ? vs[row * DIMP + 8 * mt + g] : 0.0;
#pragma unroll
        for (int mt = 0; mt < CT; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) dmma884(acc2[mt][nt][0], acc2[mt][nt][1], av[mt], bf[nt]);
      }
    }
    __syncwarp();  // all reads of the warp's dT rows are done: park the partial there (or in its own buffer)
    if (myrt[0] >= 0) {
#pragma unroll
      for (int mt = 0; mt < CT; ++mt) {
        const int dimi = 8 * mt + g;
        double* dst = a.red_sep ? redsep + ((buf * ZB_WARPS + warp) * DIMP + dimi) * 32
                                : cb;
        if (dimi < DIMP) {
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const int off = a.red_sep ? 8 * nt + 2 * t4 : zb_at(myrt[mt] * 8 + g, 8 * nt + 2 * t4);
            *reinterpret_cast<double2*>(dst + off) = make_double2(acc2[mt][nt][0], acc2[mt][nt][1]);
          }
        }
      }
    }
    psir[(buf * ZB_WARPS + warp) * 32 + lane] = psi;
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar_done[buf]);
    ZB_CLK(3);

    // ---- tail: one warp, lane = point.  Order matters for the pipeline: gather the parked partials, hand the
    // buffer back to the TMA for tile + 2 at once, and only then do the per-point arithmetic ----
    if (warp == tw) {
      mbar_wait(&bar_done[buf], par);
      ZB_CLK(4);
      double dxh[DIMP];
#pragma unroll
      for (int c = 0; c < DIMP; ++c) dxh[c] = 0.0;
      for (int w = 0; w < ZB_WARPS; ++w) {
        const int pr0 = prow[2 * w], pr1 = prow[2 * w + 1];
        if (pr0 < 0) continue;
#pragma unroll
        for (int c = 0; c < DIMP; ++c) {
          const double* src = a.red_sep ? redsep + ((buf * ZB_WARPS + w) * DIMP + c) * 32 + lane
                                        : cb + zb_at((c < 8 ? pr0 : pr1) + (c & 7), lane);
          dxh[c] += *src;
        }
      }
      double ps = 0.0;
      if (FUSE_LIK) {
#pragma unroll
        for (int w = 0; w < ZB_WARPS; ++w) ps += psir[(buf * ZB_WARPS + w) * 32 + lane];
      }
      __syncwarp();
      if (it + 2 < n_my) {
        fence_proxy_async();  // generic-proxy accesses to this buffer are ordered before the async-proxy refill
        issue_tile(it + 2);
      }
      ZB_CLK(5);
      if (valid) {
        double part[4 + DIMP];  // [0] DB, [1 .. 1 + d) DW, then E, DV, DS2
#pragma unroll
        for (int k = 0; k < 4 + DIMP; ++k) part[k] = 0.0;
        const double r = t_r, rinv = 1.0 / r;
        double drv = t_dr;
        if (FUSE_LIK) {
          double kd = kv, dkd = 0.0;
          {
            double rp = 1.0;
            for (int k = 0; k < 2 * a.order - 1; ++k) rp *= r;
            if (a.order > 0) { kd = kv * rp * r; dkd = 2.0 * a.order * kv * rp; }
          }
          // Gaussian terms (likelihoods.py:210-218) with the uniform pieces hoisted out of the loop
          const double fmu = r * t_mz, fvar = kd + r * r * ps;
          const double res = t_y - fmu, sq = res * res + fvar;
          const double pp = res * s2inv, qq = -0.5 * s2inv;
          part[1 + DIMP] = e_const - 0.5 * sq * s2inv;
          part[2 + DIMP] = qq * (kd / kv);
          part[3 + DIMP] = -0.5 * s2inv + 0.5 * sq * s2inv * s2inv;
          drv = pp * t_mz + qq * (dkd + 2.0 * r * ps);
        }
        double dot = 0.0;
#pragma unroll
        for (int c = 0; c < DIMP; ++c) dot = fma(xh[c], dxh[c], dot);
#pragma unroll
        for (int c = 0; c < DIMP; ++c) {
          const double du = (dxh[c] - xh[c] * dot) * rinv + drv * xh[c];
          const double u = xh[c] * r;
          if (c < a.d) part[1 + c] = du * u;
          else if (c == a.d) part[0] = du;
        }
        // per-lane running sums shared by the successive tail warps (they run strictly one after the other)
#pragma unroll
        for (int k = 0; k < 4 + DIMP; ++k) lsum[k * 32 + lane] += part[k];
      }
      ZB_CLK(6);
    }
  }

  // ---- write-out (accumulating over chunks) ----
  __syncthreads();
  if (warp < 2) {  // fixed-order reduction of the per-lane sums: slot k handled by (warp, pass)
    for (int k = warp; k < a.npart; k += 2) {
      int src;
      if (k == PART_E) src = 1 + DIMP;
      else if (k == PART_DV) src = 2 + DIMP;
      else if (k == PART_DS2) src = 3 + DIMP;
      else if (k == PART_DB) src = 0;
      else src = 1 + (k - PART_DW);
      const double sres = warp_sum(lsum[src * 32 + lane]);
      if (lane == 0 && (FUSE_LIK || (k != PART_E && k != PART_DV && k != PART_DS2))) a.partials[(int64_t)blockIdx.x * a.npart + k] += sres;
    }
  }
#pragma unroll
  for (int r = 0; r < ZB_SLOTS; ++r) {
    if (myrt[r] < 0) continue;
    const int row = myrt[r] * 8 + g;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
      double2* dst = reinterpret_cast<double2*>(a.dVpart + ((int64_t)blockIdx.x * Mp + row) * (CT * 8) + c * 8 + 2 * t4);
      double2 o = *dst;
      o.x += accv[r][c][0];
      o.y += accv[r][c][1];
      *dst = o;
    }
  }
}

}  // namespace gpfy
