// K_G8  the dense contraction  C = W2 Z  (variational.py:323-328 folded) in float64 accuracy on the 5th-generation
// tensor cores: an Ozaki-style error-free split of both float64 operands into signed radix-256 digits, the digit products as
// tcgen05 kind::i8 MMAs with exact int32 accumulators in TMEM, recombined in float64 in the epilogue.
//
//   z'[k, n]  = Z[k, n] / s_k          s_k  = power of two >= 2.04 max_t |C_l^alpha(t)|  (per feature row, i.e. per degree)
//   w'[o, k]  = W2[o, k] s_k / r_o     r_o  = power of two >= 2.04 max_k |W2[o, k] s_k|  (per output row)
//   x         = sum_{i = 1..S} x_i 256^-i + e,   x_i in [-128, 127],  |e| <= 2^-49         (S = 6 digits, |x| <= 0.4902)
//   C[o, n]   = r_o sum_{g = 2..S+1} 256^-g  sum_{i + j = g} sum_k a_i[n, k] b_j[o, k]     (21 digit products, 6 accumulators)
//
// Every digit product is exact (|sum| <= 6 * 864 * 2^14 < 2^27 in int32); the error is the two operands' truncation (2^-49 of the row /
// column maximum each) plus the dropped products i + j >= S + 2 (<= 5 * K * 2^14 * 256^-8 = 2^-39.5 K-worst-case, ~2^-46 typical) -
// normwise ~1e-14 of max|C| against ~1e-15 for the DMMA kernel and the 1e-10 contract (tests/test_gpu_i8.py).
//
// Shape: TMEM lanes = 128 points of a tile, TMEM columns = S accumulators x NB output rows of one OUT-BLOCK (NB <= 80, NB % 16 = 0:
// 6 x 80 = 480 of the 512 columns).  A CTA owns one out-block for the whole launch: the block's digit planes of W2 (S x NB x Mp
// bytes) stay resident in shared memory, the digit planes of Z stream through a ring of 24 KB stages (one 32-feature k-block of all
// S digits = ONE TMA bulk copy: the slicing kernel writes them in exactly the un-swizzled canonical K-major order the MMA reads,
// 8 x 16-byte core matrices, LBO = 2 KB between the two 16-feature halves, SBO = 128 B between 8-point groups).  The SMs are dealt
// to the out-blocks in proportion to a block's MMA time, max(NB / 2, 32 + NB / 4) cycles per product (measured, tools/i8_probe.cu:
// below N = 128 an MMA is bound by the 128 B/clk shared-memory read of its operands, 4 KB of A + 32 NB of B).
// Roles (704 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one lane; two issuing warps were measured SLOWER: 89 cycles per
// MMA against 52), warps 2..21 = epilogue, one per TMEM lane quarter and 16 columns of the block (tcgen05.ld, accumulator pairs folded
// in integer arithmetic, converted and chained in float64, row scale, C stored in the tile-contiguous swizzled layout zonal_bwd3 /
// zonal_bwd4 read; optional psi = z . C partial sums - the library's callers now form psi in lik_point_kernel / predict_kernel
// instead, because FP64 instructions of these warps are slow while the tensor pipe is saturated).  The accumulators of a digit-group
// pair are handed back to the MMA warp as soon as the epilogue has read them, so the next tile's first k-block overlaps the drain.
#pragma once
#include "common.cuh"
#include "gemm_tf32.cuh"

namespace gpfy {

constexpr int GI8_S = 6;                       // digits per operand
constexpr int GI8_BM = 128;                    // points per tile
constexpr int GI8_THREADS = 704;               // producer warp, MMA warp, 20 epilogue warps (one per TMEM lane quarter and 16 columns)
constexpr int GI8_A_STAGE = GI8_S * 4096;      // bytes of one k-block (32 features) of a tile, all digits
constexpr int GI8_MAXBLK = 32;
constexpr int GI8_MAXNB = 80;
constexpr int GI8_STATIC_SMEM = 12 * 1024;    // barriers, row scales, psi staging (2 x 5 x 4 x 32 doubles): see gemm_i8_kernel
#ifndef GI8_SLEEP_NS
#define GI8_SLEEP_NS 64
#endif
#ifndef GI8_PREFETCH
#define GI8_PREFETCH 4   // k-blocks of L2 prefetch ahead of the ring: with two stages (Mp = 352) the MMA warp's wait on the ring halves (2.9 -> 1.5 k cycles per tile)
#endif
constexpr double GI8_HEADROOM = 2.04;          // |x| <= 1 / 2.04 = 0.4902 < (127 + 127/256 + ...) / 256

struct GemmI8Blocks {
  int nblk;
  int row0[GI8_MAXBLK], rows[GI8_MAXBLK];
  int cta0[GI8_MAXBLK + 1];      // CTAs cta0[b] .. cta0[b + 1] - 1 work on block b
  int64_t woff[GI8_MAXBLK];      // byte offset of block b's digit planes [S][Mp / 16][rows / 8][8][16] inside the W2 image
};

struct GemmI8Args {
  const int8_t* Zs;     // digit planes of Z: [tiles of 128 points][Mp / 32 k-blocks][S][2 halves][16 point groups][8][16]
  const int8_t* Ws;     // digit planes of W2 (GemmI8Blocks::woff)
  const double* rs;     // [Mp] r_o * 2^-24
  const double* Z;      // float64 panel [Mp][ldz] for the psi epilogue (null when psi is null)
  int64_t ldz;
  double* C;            // [tiles32][Mp][32] blocked + swizzled (GemmArgs::c_blocked), or null
  double* psi;          // [nblk][ld_psi] partial sums of z_n . C_n over each out-block's rows, or null
  int64_t ld_psi;
  int Mp;
  int64_t npts;
  int nsa;              // stages of the Z ring
#ifdef GI8_PROFILE
  long long* prof;      // [grid][8] cycle counters (development builds, tools/i8_gemm_bench.cu)
#endif
  GemmI8Blocks blk;
};

// out-blocks: the widest blocks (<= 80 rows, whole 16-row units, as equal as possible) whose resident W2 planes leave room for at
// least two stages of the Z ring - 80 rows up to Mp = 352, 64 at 384 .. 448, 48 at 576 (T = 64), 32 beyond; SMs in proportion to the
// MMA time per tile
static inline bool gi8_plan(int Mp, int sms, GemmI8Blocks* b, int* nsa) {
  if (Mp % 32 || Mp < 32) return false;
  const int units = Mp / 16;
  for (int nbmax = GI8_MAXNB; nbmax >= 32; nbmax -= 16) {
    const int nblk = (Mp + nbmax - 1) / nbmax;
    if (nblk > GI8_MAXBLK || sms < nblk) continue;
    b->nblk = nblk;
    int r0 = 0, maxrows = 0;
    int64_t off = 0;
    double wsum = 0.0, w[GI8_MAXBLK];
    for (int i = 0; i < nblk; ++i) {
      const int u = units / nblk + (i < units % nblk ? 1 : 0);
      b->row0[i] = r0;
      b->rows[i] = 16 * u;
      b->woff[i] = off;
      r0 += 16 * u;
      off += (int64_t)GI8_S * 16 * u * Mp;
      maxrows = 16 * u > maxrows ? 16 * u : maxrows;
      const double wi = 16 * u / 2.0 > 32.0 + 16 * u / 4.0 ? 16 * u / 2.0 : 32.0 + 16 * u / 4.0;
      w[i] = wi > 45.0 ? wi : 45.0;   // no kind::i8 MMA issues faster than ~45 cycles (tools/i8_probe.cu)
      wsum += w[i];
    }
    const int avail = 227 * 1024 - 1024 - GI8_STATIC_SMEM - GI8_S * maxrows * Mp;   // alignment slack, static shared memory, resident W2 planes
    if (avail < 2 * GI8_A_STAGE) continue;
    *nsa = avail / GI8_A_STAGE > 4 ? 4 : avail / GI8_A_STAGE;
    int given = 0;
    for (int i = 0; i < nblk; ++i) {
      b->cta0[i] = given;
      int c = (int)(sms * w[i] / wsum);
      if (c < 1) c = 1;
      given += c;
    }
    b->cta0[nblk] = given;
    // leftover SMs to the blocks in order (heaviest first: blocks are sorted by size)
    int left = sms - given;
    for (int i = 0; left > 0; i = (i + 1) % nblk, --left)
      for (int j = i + 1; j <= nblk; ++j) b->cta0[j] += 1;
    if (b->cta0[nblk] > sms) continue;
    return true;
  }
  return false;
}
static inline int gi8_smem_bytes(int Mp, const GemmI8Blocks& b, int nsa) {
  int maxrows = 0;
  for (int i = 0; i < b.nblk; ++i) maxrows = b.rows[i] > maxrows ? b.rows[i] : maxrows;
  return 1024 + GI8_S * maxrows * Mp + nsa * GI8_A_STAGE;
}
static inline int64_t gi8_w_image_bytes(int Mp) { return (int64_t)GI8_S * Mp * Mp; }
static inline int64_t gi8_z_image_bytes(int Mp, int64_t nc) { return (nc + GI8_BM - 1) / GI8_BM * (int64_t)(Mp / 32) * GI8_A_STAGE; }

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t gi8_desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14); }   // version 1, no swizzle
__device__ __forceinline__ void gi8_mma(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\nsetp.ne.b32 p, %6, 0;\n"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %5, p;\n}\n"
               ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void gi8_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ bool gi8_elect_one() { return g32_elect_one(); }

// waits of the warps that are NOT on the critical path back off between polls: nine warps spinning on mbarriers
// were measured to slow the MMA warp's issue rate (tools/i8_gemm_bench.cu)
__device__ __forceinline__ void gi8_wait_relaxed(uint64_t* bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(GI8_SLEEP_NS);
}

__device__ __forceinline__ void gi8_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
}

// 16 values -> their S signed radix-256 digits, digit s of all 16 packed into one 16-byte vector (= one row of a core matrix).
// q = round(x 2^48) by the magic-number add; T = q + 0x80..80 makes every digit unsigned (digit + 128), XOR 0x80 turns the bytes
// back into two's-complement digits; byte 5 of T is the most significant digit (s = 1).
struct Digits16 { uint32_t w[GI8_S][4]; };
__device__ __forceinline__ uint64_t gi8_digits(double x, double scale) {
  // one FMA: x * scale (a power of two: exact) + 1.5 * 2^52 rounds to the nearest integer in the low mantissa bits.  |x scale| <=
  // 2^48 / 2.04 by construction of the scales (measured on B200: DMUL issues at a fraction of the DFMA rate, so no separate multiply)
  const double magic = 6755399441055744.0;        // 1.5 * 2^52
  const long long q = __double_as_longlong(fma(x, scale, magic)) - __double_as_longlong(magic);
  return (uint64_t)(q + 0x0000808080808080ll) ^ 0x0000808080808080ull;
}
// x * y as an FMA with -0.0 (bit-identical to the product; the compiler would turn fma(x, y, 0) back into the slow DMUL)
__device__ __forceinline__ double gi8_mul(double x, double y, double neg_zero) { return fma(x, y, neg_zero); }
// -0.0 the compiler cannot constant-fold (flag is zero at run time)
__device__ __forceinline__ double gi8_neg_zero(int flag_zero) { return __longlong_as_double((long long)(0x8000000000000000ull | (unsigned long long)(flag_zero != 0))); }
__device__ __forceinline__ void gi8_pack4(const uint64_t (&t)[4], Digits16& d, int word) {
  const uint32_t l0 = (uint32_t)t[0], l1 = (uint32_t)t[1], l2 = (uint32_t)t[2], l3 = (uint32_t)t[3];
  const uint32_t h0 = (uint32_t)(t[0] >> 32), h1 = (uint32_t)(t[1] >> 32), h2 = (uint32_t)(t[2] >> 32), h3 = (uint32_t)(t[3] >> 32);
  // low words hold digits 6 (byte 0) .. 3 (byte 3), high words digits 2 (byte 0) and 1 (byte 1)
#pragma unroll
  for (int by = 0; by < 4; ++by) {
    const uint32_t sel = (uint32_t)by | ((uint32_t)(4 + by) << 4);
    const uint32_t p01 = __byte_perm(l0, l1, sel), p23 = __byte_perm(l2, l3, sel);
    d.w[5 - by][word] = __byte_perm(p01, p23, 0x5410);
  }
#pragma unroll
  for (int by = 0; by < 2; ++by) {
    const uint32_t sel = (uint32_t)by | ((uint32_t)(4 + by) << 4);
    const uint32_t p01 = __byte_perm(h0, h1, sel), p23 = __byte_perm(h2, h3, sel);
    d.w[1 - by][word] = __byte_perm(p01, p23, 0x5410);
  }
}

// Z [Mp][ldz] float64 -> digit planes.  grid (tiles of 128 points, Mp / 16), 128 threads: thread = one point x 16 features; the 16
// row reads are coalesced over the warp's points and each digit vector is one 16-byte store, 512 contiguous bytes per warp.
__global__ void __launch_bounds__(128) z_slice_i8_kernel(const double* __restrict__ Z, int64_t ldz, int Mp, int64_t npts,
                                                         const double* __restrict__ kscale, int8_t* out) {
  const int kc = blockIdx.y, p = threadIdx.x;
  const int64_t n = (int64_t)blockIdx.x * GI8_BM + p;
  const bool ok = n < npts;
  Digits16 d;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    uint64_t t[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kc * 16 + 4 * w + j;
      const double z = ok ? Z[(int64_t)k * ldz + n] : 0.0;
      t[j] = gi8_digits(z, kscale[k]);
    }
    gi8_pack4(t, d, w);
  }
  int8_t* dst = out + ((int64_t)blockIdx.x * (Mp / 32) + (kc >> 1)) * GI8_A_STAGE + (kc & 1) * 2048 + p * 16;
#pragma unroll
  for (int s = 0; s < GI8_S; ++s) *reinterpret_cast<uint4*>(dst + s * 4096) = make_uint4(d.w[s][0], d.w[s][1], d.w[s][2], d.w[s][3]);
}

// W2 [Mp][Mp] float64 (row = output) -> per-row scale + digit planes in the out-block images; also the per-feature scales of Z.
// One warp per output row; bound[k] = max |z_k| over the sphere (C_l^alpha(1) of the row's degree, 1 for padding rows).
__global__ void __launch_bounds__(128) w2_slice_i8_kernel(const double* __restrict__ W2, int Mp, const double* __restrict__ bound,
                                                          GemmI8Blocks blk, int8_t* Ws, double* rs, double* kscale) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blockIdx.x * 4 + warp;
  if (o >= Mp) return;
  const double* row = W2 + (int64_t)o * Mp;
  double mx = 0.0;
  for (int k = lane; k < Mp; k += 32) {
    int e;
    frexp(GI8_HEADROOM * bound[k], &e);
    const double sk = ldexp(1.0, e);           // power of two > 2.04 bound
    if (o == 0) kscale[k] = 281474976710656.0 / sk;   // 2^48 / s_k
    mx = fmax(mx, fabs(row[k]) * sk);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, s));
  int e = 0;
  if (mx > 0.0) frexp(GI8_HEADROOM * mx, &e);
  const double r = ldexp(1.0, e);
  if (lane == 0) rs[o] = r * (1.0 / 16777216.0);   // 256^-2 of the leading digit pair, 256^-1 of the pairwise fold
  int b = 0;
  while (b + 1 < blk.nblk && o >= blk.row0[b + 1]) ++b;
  const int NB = blk.rows[b], lr = o - blk.row0[b];
  const double sc = 281474976710656.0 / r;
  for (int kc = lane; kc < Mp / 16; kc += 32) {
    Digits16 d;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      uint64_t t[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = kc * 16 + 4 * w + j;
        int ek;
        frexp(GI8_HEADROOM * bound[k], &ek);
        t[j] = gi8_digits(row[k], sc * ldexp(1.0, ek));
      }
      gi8_pack4(t, d, w);
    }
    int8_t* dst = Ws + blk.woff[b] + (int64_t)kc * NB * 16 + (lr >> 3) * 128 + (lr & 7) * 16;
#pragma unroll
    for (int s = 0; s < GI8_S; ++s) *reinterpret_cast<uint4*>(dst + (int64_t)s * NB * Mp) = make_uint4(d.w[s][0], d.w[s][1], d.w[s][2], d.w[s][3]);
  }
}

// epilogue of one tile, one warp: thread = one point (TMEM lane), c[] = its outputs for the 16 columns col0 .. col0 + 15 of the
// out-block.  The six int32 accumulators are folded pairwise in INTEGER arithmetic (acc_2p * 256 + acc_2p+1, one IMAD.WIDE),
// converted with the 1.5 * 2^52 magic-number add and chained by one FMA per pair: 6 FP64-pipe instructions per output instead of
// 13 (I2F.F64 + DFMA per accumulator: measured, the conversions made the epilogue - not the MMAs - the bottleneck).  Both
// accumulators of a pair are handed back to the MMA warp as soon as they have been read.  rs_s = the block's row scales.
__device__ __forceinline__ double gi8_epilogue16(const GemmI8Args& a, uint32_t tmem_lane, int NB, int row0, int col0, int64_t p0, int lane,
                                                 uint64_t* d_empty, const double* rs_s) {
#if defined(GI8_KO) && (GI8_KO & 8)   // development knock-out: the epilogue only hands the accumulators back
  __syncwarp();
  if (lane == 0)
    for (int g = 0; g < GI8_S; ++g) mbar_arrive(&d_empty[g]);
  return 0.0;
#endif
  double c[16];
#pragma unroll
  for (int pr = GI8_S / 2 - 1; pr >= 0; --pr) {   // accumulators 2 pr (weight 256) and 2 pr + 1; lowest weights first
    uint32_t vh[16], vl[16];
#if defined(GI8_KO) && (GI8_KO & 2)   // development knock-out: no TMEM loads
#pragma unroll
    for (int j = 0; j < 16; ++j) { vh[j] = (uint32_t)(pr + j + lane); vl[j] = vh[j] * 3u; }
#else
    gi8_ld16_nowait(tmem_lane + (uint32_t)((2 * pr + 1) * NB + col0), vl);
    gi8_ld16_nowait(tmem_lane + (uint32_t)((2 * pr) * NB + col0), vh);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#endif
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncwarp();
    if (lane == 0) { mbar_arrive(&d_empty[2 * pr + 1]); mbar_arrive(&d_empty[2 * pr]); }   // the MMA warp may overwrite them
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const long long pair = (long long)(int)vh[j] * 256 + (long long)(int)vl[j];
#if defined(GI8_KO) && (GI8_KO & 16)   // development knock-out: no FP64-pipe instruction in the epilogue (wrong values)
      c[j] = __longlong_as_double((pr == GI8_S / 2 - 1) ? pair : (__double_as_longlong(c[j]) >> 16) + pair);
#else
      const double pd = __longlong_as_double(pair + 0x4338000000000000ll) - 6755399441055744.0;   // exact: |pair| < 2^51
      c[j] = (pr == GI8_S / 2 - 1) ? pd : fma(c[j], 1.0 / 65536.0, pd);
#endif
    }
  }
  const int Mp = a.Mp;
  const int64_t n = p0 + lane;
  const bool store = a.C != nullptr && p0 < ((a.npts + 31) & ~(int64_t)31);
  double* ctile = a.C + (p0 >> 5) * (int64_t)Mp * 32;
  double psi = 0.0;
  const double nz = gi8_neg_zero(Mp == 0);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int row = row0 + col0 + j;
#if defined(GI8_KO) && (GI8_KO & 16)
    const double cv = __longlong_as_double(__double_as_longlong(c[j]) + __double_as_longlong(rs_s[col0 + j]));
#else
    const double cv = gi8_mul(c[j], rs_s[col0 + j], nz);
#endif
    if (store) ctile[row * 32 + (lane ^ ((row & 3) << 2))] = cv;
    if (a.psi) psi = fma(n < a.npts ? a.Z[(int64_t)row * a.ldz + n] : 0.0, cv, psi);
  }
  return psi;
}

__global__ void __launch_bounds__(GI8_THREADS, 1) gemm_i8_kernel(const __grid_constant__ GemmI8Args a) {
  extern __shared__ __align__(1024) uint8_t gi8_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)gi8_smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t a_full[4], a_empty[4], b_full, d_full, d_empty[GI8_S];
  __shared__ uint32_t tmem_base_s;
  __shared__ double rs_s[GI8_MAXNB];
  __shared__ double psi_s[2][GI8_MAXNB / 16][4][32];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler: the role branches are divergence-free
  int b = 0;
  while (b + 1 < a.blk.nblk && (int)blockIdx.x >= a.blk.cta0[b + 1]) ++b;
  const int NB = a.blk.rows[b], row0 = a.blk.row0[b];
  const int gsize = a.blk.cta0[b + 1] - a.blk.cta0[b], gidx = (int)blockIdx.x - a.blk.cta0[b];
  const int Mp = a.Mp, KB = Mp / 32, NSA = a.nsa;
  const int tiles = (int)((a.npts + GI8_BM - 1) / GI8_BM);
  const int n_my = gidx < tiles ? (tiles - gidx + gsize - 1) / gsize : 0;
  const int b_slice = NB * Mp;                                  // bytes of one digit plane of the block
  uint8_t* sB = smem;                                           // [S][Mp / 16][NB / 8][8][16]
  uint8_t* sA = smem + ((GI8_S * b_slice + 1023) & ~1023);      // [NSA][S][2][16][8][16]

  if (tid == 0) {
    for (int s = 0; s < NSA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    mbar_init(&b_full, 1);
    mbar_init(&d_full, 1);
    for (int g = 0; g < GI8_S; ++g) mbar_init(&d_empty[g], 4 * (NB / 16));
  }
  if (tid < NB) rs_s[tid] = a.rs[row0 + tid];
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(g32_smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_base_s, 0);

  if (warp == 0) {
    // ================= TMA producer =================
    if (n_my > 0) {
      if (gi8_elect_one()) {
        mbar_expect_tx(&b_full, (unsigned)(GI8_S * b_slice));
        for (int s = 0; s < GI8_S; ++s) tma_bulk_g2s(sB + s * b_slice, a.Ws + a.blk.woff[b] + (int64_t)s * b_slice, (unsigned)b_slice, &b_full);
      }
      __syncwarp();
      int ia = 0;
      for (int it = 0; it < n_my; ++it) {
        const int8_t* src = a.Zs + (int64_t)(gidx + it * gsize) * KB * GI8_A_STAGE;
        for (int kb = 0; kb < KB; ++kb, ++ia) {
          const int st = ia % NSA;
          if (ia >= NSA) gi8_wait_relaxed(&a_empty[st], (unsigned)(((ia / NSA) - 1) & 1));
          if (gi8_elect_one()) {
#if defined(GI8_KO) && (GI8_KO & 4)   // development knock-out: no Z traffic (the MMAs read whatever is in the ring)
            mbar_expect_tx(&a_full[st], 0);
            (void)src;
#else
            mbar_expect_tx(&a_full[st], GI8_A_STAGE);
            tma_bulk_g2s(sA + st * GI8_A_STAGE, src + (int64_t)kb * GI8_A_STAGE, GI8_A_STAGE, &a_full[st]);
#if GI8_PREFETCH > 0
            {   // pull the stage GI8_PREFETCH k-blocks ahead into L2 (the ring holds only nsa stages of latency cover)
              const int f = kb + GI8_PREFETCH;
              const int8_t* nx = f < KB ? src + (int64_t)f * GI8_A_STAGE
                                        : (it + 1 < n_my ? a.Zs + ((int64_t)(gidx + (it + 1) * gsize) * KB + (f - KB)) * GI8_A_STAGE : nullptr);
              if (nx) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nx), "r"(GI8_A_STAGE) : "memory");
            }
#endif
#endif
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The whole warp walks the loop and waits on the barriers; ONE elected lane issues.  With the issue under `if (lane == 0)`
    // the operands live in per-thread registers and the compiler wraps every tcgen05.mma in an ELECT / R2UR / BRA.U.ANY
    // "uniformisation" loop: ~75 cycles per MMA measured against the 52 of the tensor pipe itself.
    if (n_my > 0) {
      // instruction descriptor: D s32 (2 << 4), A / B signed 8 bit (1 << 7, 1 << 10), both K-major, N = NB, M = 128
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(GI8_BM >> 4) << 24);
      const uint32_t hi = gi8_desc_hi(128);
      const uint32_t bstep = (uint32_t)b_slice >> 4;            // digit-plane stride of B in descriptor units
      const uint32_t b_lo0 = g32_desc_lo(g32_smem_u32(sB), (uint32_t)NB * 16);
      mbar_wait(&b_full, 0);
      int ia = 0;
#ifdef GI8_PROFILE
      long long t_a = 0, t_d = 0, t_all = clock64();
#define GI8_T0 const long long t0__ = clock64()
#define GI8_T1(acc) acc += clock64() - t0__
#else
#define GI8_T0
#define GI8_T1(acc)
#endif
      for (int it = 0; it < n_my; ++it) {
        for (int kb = 0; kb < KB; ++kb, ++ia) {
          const int st = ia % NSA;
          { GI8_T0; mbar_wait(&a_full[st], (unsigned)((ia / NSA) & 1)); GI8_T1(t_a); }
          asm volatile("tcgen05.fence::after_thread_sync;");
          const uint32_t a_lo = g32_desc_lo(g32_smem_u32(sA + st * GI8_A_STAGE), 2048);
          const uint32_t b_lo = b_lo0 + (uint32_t)(kb * 2 * NB);   // k-block kb = 16-feature chunks 2 kb, 2 kb + 1 (NB * 16 bytes each)
          if (kb == 0 && it > 0) {
            // first k-block of a tile: the accumulators come back from the epilogue in pairs (5, 4), (3, 2), (1, 0) - lowest weights
            // first; the products of a pair are issued alternately, because consecutive MMAs on ONE accumulator run at ~80 cycles
            // each instead of 52 (tools/i8_probe.cu, tools/i8_gemm_bench.cu)
#pragma unroll
            for (int pr = GI8_S / 2 - 1; pr >= 0; --pr) {
              const int g1 = 2 * pr + 1, g0 = 2 * pr;
              { GI8_T0; mbar_wait(&d_empty[g1], (unsigned)((it - 1) & 1)); mbar_wait(&d_empty[g0], (unsigned)((it - 1) & 1)); GI8_T1(t_d); }
              asm volatile("tcgen05.fence::after_thread_sync;");
              if (gi8_elect_one()) {
#pragma unroll
                for (int i = 0; i <= g1; ++i) {                 // digits i + 1 of Z and g - i + 1 of W2
                  gi8_mma(tmem + (uint32_t)(g1 * NB), a_lo + 256u * i, hi, b_lo + bstep * (uint32_t)(g1 - i), hi, idesc, i ? 1u : 0u);
                  if (i <= g0) gi8_mma(tmem + (uint32_t)(g0 * NB), a_lo + 256u * i, hi, b_lo + bstep * (uint32_t)(g0 - i), hi, idesc, i ? 1u : 0u);
                }
              }
              __syncwarp();
            }
          } else {
            // every other k-block: consecutive MMAs go to DIFFERENT accumulators (digit i + 1 of Z against digits 1 .. S - i of W2),
            // the order tools/i8_probe.cu measured at the shared-memory floor of 32 + NB / 4 cycles per MMA
            if (gi8_elect_one()) {
#pragma unroll
              for (int i = 0; i < GI8_S; ++i)
#pragma unroll
                for (int j = 0; j < GI8_S - i; ++j)
                  gi8_mma(tmem + (uint32_t)((i + j) * NB), a_lo + 256u * i, hi, b_lo + bstep * (uint32_t)j, hi, idesc, (kb | i) ? 1u : 0u);
            }
            __syncwarp();
          }
          if (gi8_elect_one()) g32_commit(&a_empty[st]);
          __syncwarp();
        }
        if (gi8_elect_one()) g32_commit(&d_full);
        __syncwarp();
      }
#ifdef GI8_PROFILE
      if (lane == 0) {
      a.prof[blockIdx.x * 8 + 0] = clock64() - t_all;
      a.prof[blockIdx.x * 8 + 1] = t_a;
      a.prof[blockIdx.x * 8 + 2] = t_d;
      a.prof[blockIdx.x * 8 + 3] = n_my;
      }
#endif
    }
  } else {
    // ================= epilogue: warps 2..21; warp w reads TMEM lanes 32 (w % 4) .. + 31, columns 16 u .. 16 u + 15, u = (w - 2) / 4 ====
    const int q = warp & 3, u = (warp - 2) >> 2, nunits = NB / 16;
    if (u < nunits)
    for (int it = 0; it < n_my; ++it) {
      const int64_t p0 = (int64_t)(gidx + it * gsize) * GI8_BM + 32 * q;
#ifdef GI8_PROFILE
      const long long e0 = clock64();
#endif
      gi8_wait_relaxed(&d_full, (unsigned)(it & 1));
#ifdef GI8_PROFILE
      const long long e1 = clock64();
#endif
      asm volatile("tcgen05.fence::after_thread_sync;");
      const double psi = gi8_epilogue16(a, tmem + ((uint32_t)(32 * q) << 16), NB, row0, 16 * u, p0, lane, d_empty, rs_s);
#ifdef GI8_PROFILE
      if (warp == 2 && lane == 0) { a.prof[blockIdx.x * 8 + 4] += e1 - e0; a.prof[blockIdx.x * 8 + 5] += clock64() - e1; }
#endif
      if (a.psi) {   // the column units of a lane quarter park their sums, unit 0 adds them in a fixed order and stores;
                     // psi_s is double-buffered over the tiles, so one named barrier (quarter q, 32 nunits threads) per tile is enough
        psi_s[it & 1][u][q][lane] = psi;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(32 * nunits) : "memory");
        if (u == 0 && p0 + lane < a.npts) {
          double t = 0.0;
          for (int k = 0; k < nunits; ++k) t += psi_s[it & 1][k][q][lane];
          a.psi[(int64_t)b * a.ld_psi + p0 + lane] = t;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
#endif  // __CUDACC__

}  // namespace gpfy
