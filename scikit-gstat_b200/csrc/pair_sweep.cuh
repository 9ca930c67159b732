// Sweep kernels shared by the translation units that instantiate them (pair_kernels.cu: scalar
// values; pair_select.cu: select passes; pair_mv.cu: multi-column values).  Device code only.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace skg {

// ------------------------------------------------------------ tile culling ---
//
// Points are Z-order sorted, so the 256-point blocks are spatially compact.  From the
// bounding boxes of two blocks the squared distances of all their pairs lie in
// [dmin2, dmax2]; a tile whose interval misses every window of interest (beyond maxlag;
// outside the classes that hold a wanted rank; cannot contain the maximum) is dropped
// before the sweep.  Exact: dropped pairs contribute to nothing (1e-12 relative slack
// covers the rounding of the box arithmetic).  The surviving tiles are written in
// ascending order (single-CTA ordered compaction), so the sweep and its fixed-order
// reductions stay deterministic.

struct CullWindows {
  int count;       // windows in use
  int want_max;    // also keep tiles that can hold the maximum squared distance
  double lo[64];   // inclusive
  double hi[64];   // exclusive
  // directional mask (0 = none): work items whose box of difference vectors cannot meet the
  // azimuth cone / bandwidth strip are dropped as well (conservative, hence exact)
  int dir_model, dir_tol_all;
  double cos_az, sin_az, tan_ht, half_bw;
};

inline void cull_set_dir(CullWindows* w, const DirParams& dp) {
  w->dir_model = dp.model;
  w->dir_tol_all = dp.tol_all;
  w->cos_az = dp.cos_az; w->sin_az = dp.sin_az; w->tan_ht = dp.tan_ht; w->half_bw = dp.half_bw;
}

template <class W> struct DirView;  // uniform access to the direction fields of CullWindows / DirParams
template <> struct DirView<CullWindows> {
  static __device__ __forceinline__ int model(const CullWindows& w) { return w.dir_model; }
  static __device__ __forceinline__ int tol_all(const CullWindows& w) { return w.dir_tol_all; }
};
template <> struct DirView<DirParams> {
  static __device__ __forceinline__ int model(const DirParams& w) { return w.model; }
  static __device__ __forceinline__ int tol_all(const DirParams& w) { return w.tol_all; }
};

// Can any difference vector P_i - P_j between the two boxes pass the directional mask?
// dot = dx cos - dy sin and crs = dx sin + dy cos (dir_fast) are linear in (dx, dy), so their
// ranges over the rectangle of difference vectors come from its corners; the mask needs
// |crs| <= tan(tol/2) |dot| (and |crs| <= bandwidth/2 for the triangle).
template <class W>
__device__ __forceinline__ bool dir_box_may_pass(const double* ibox, const double* jbox, const W& w) {
  const double dx0 = ibox[0] - jbox[3], dx1 = ibox[3] - jbox[0];
  const double dy0 = ibox[1] - jbox[4], dy1 = ibox[4] - jbox[1];
  double dmin = 1e300, dmax = -1e300, cmin = 1e300, cmax = -1e300;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double dx = (k & 1) ? dx1 : dx0, dy = (k & 2) ? dy1 : dy0;
    const double dot = dx * w.cos_az - dy * w.sin_az;
    const double crs = dx * w.sin_az + dy * w.cos_az;
    dmin = fmin(dmin, dot); dmax = fmax(dmax, dot);
    cmin = fmin(cmin, crs); cmax = fmax(cmax, crs);
  }
  if (!(dmin == dmin && cmin == cmin)) return true;  // NaN boxes: keep
  const double dot_hi = fmax(fabs(dmin), fabs(dmax));
  const double crs_lo = (cmin <= 0.0 && cmax >= 0.0) ? 0.0 : fmin(fabs(cmin), fabs(cmax));
  const double slack = 1e-9 * (fabs(dx0) + fabs(dx1) + fabs(dy0) + fabs(dy1));
  if (!DirView<W>::tol_all(w) && crs_lo > w.tan_ht * dot_hi + slack) return false;
  if (DirView<W>::model(w) == SKG_DIR_TRIANGLE && crs_lo > w.half_bw + slack) return false;
  return true;
}

struct SweepScratch {  // views into the caller's scratch buffer
  unsigned long long* counters;  // [0] = number of listed tiles
  double* bbox;                  // [nt][6]: lo x,y,w ; hi x,y,w
  double* bbox32;                // [nt*8][6]: boxes of the 32-point sub-blocks
  double* bbox8;                 // [nt*32][4]: lo x, lo y, hi x, hi y of the 8-point groups (2-D)
  unsigned long long* scan;      // [0] block ticket, [1 + b] look-back status of the ordered compaction
  int64_t scan_bytes;
  unsigned int* list;            // surviving work items, ascending
  unsigned char* tail;           // kernel-specific remainder (hist partials)
  int64_t tail_bytes;
};

template <int DIM>
__device__ __forceinline__ void tile_range(const double* __restrict__ bbox, int I, int J,
                                           double& dmin2, double& dmax2) {
  dmin2 = 0.0; dmax2 = 0.0;
#pragma unroll
  for (int k = 0; k < DIM; ++k) {
    const double loI = bbox[(int64_t)I * 6 + k], hiI = bbox[(int64_t)I * 6 + 3 + k];
    const double loJ = bbox[(int64_t)J * 6 + k], hiJ = bbox[(int64_t)J * 6 + 3 + k];
    const double gap = fmax(0.0, fmax(loJ - hiI, loI - hiJ));
    const double far = fmax(fabs(hiJ - loI), fabs(hiI - loJ));
    dmin2 += gap * gap;
    dmax2 += far * far;
  }
}

// squared-distance range between the box of a 32-point sub-block and a box held in
// shared memory (lo[3], hi[3]); used for warp-level culling inside the sweeps
template <int DIM>
__device__ __forceinline__ void subblock_range(const double* __restrict__ bbox32, int64_t sub,
                                               const double* jbox, double& dmin2, double& dmax2) {
  dmin2 = 0.0; dmax2 = 0.0;
#pragma unroll
  for (int k = 0; k < DIM; ++k) {
    const double loI = bbox32[sub * 6 + k], hiI = bbox32[sub * 6 + 3 + k];
    const double loJ = jbox[k], hiJ = jbox[3 + k];
    const double gap = fmax(0.0, fmax(loJ - hiI, loI - hiJ));
    const double far = fmax(fabs(hiJ - loI), fabs(hiI - loJ));
    dmin2 += gap * gap;
    dmax2 += far * far;
  }
}

// box of j-chunk h (TILE/js points) of block J: union of the 32-point sub-block boxes
template <int DIM>
__device__ __forceinline__ void chunk_box(const double* __restrict__ bbox32, int J, int h, int js,
                                          double* box /*[6]*/) {
  const int jchunk = TILE / js;
  const int s0 = (h * jchunk) >> 5, s1 = (h * jchunk + jchunk - 1) >> 5;
  const double inf = __longlong_as_double((long long)INF_BITS);
#pragma unroll
  for (int k = 0; k < DIM; ++k) { box[k] = inf; box[3 + k] = -inf; }
  for (int sb = s0; sb <= s1; ++sb) {
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      box[k] = fmin(box[k], bbox32[((int64_t)J * 8 + sb) * 6 + k]);
      box[3 + k] = fmax(box[3 + k], bbox32[((int64_t)J * 8 + sb) * 6 + 3 + k]);
    }
  }
}

// -------------------------------------------------------------- K2: hist ---
//
// The hot kernel.  CTA = HB threads, each thread keeps HR "i" points in
// registers (HB*HR = TILE) and walks the "j" tile from shared memory.  Per-lag
// accumulators are PRIVATE to the thread (acc[lag][tid], bank-conflict free, no
// atomics): plain shared-memory read-modify-write.  The j tile and the lag tables
// are STATIC shared arrays and the accumulators the only DYNAMIC one, so the
// compiler knows accumulator stores cannot alias table/tile loads and is free to
// hoist the next pairs' loads and fp64 math above them.

constexpr int HB_THREADS = 64;  // threads per CTA of the fused / select sweeps (= HB)
constexpr int SEL_GROUP = 16;   // j points per group of the select passes' group-level culling
constexpr int SEL_MAXWIN = 32;  // distance windows the group-level test scans directly

struct DzTable {  // per-lag |dz| windows of the below-count mode (STATS == 4)
  double lo[THR_MAX];     // pairs with |dz| < lo are counted "below"
  double hi[THR_MAX];     // with a window histogram: lo <= |dz| < hi lands in a bucket
  double scale[THR_MAX];  // bucket = min(B-1, trunc(fl(fl(|dz| - lo) * scale))), as the select passes
};

struct HistPartial {  // one per (CTA, lag class)
  unsigned long long count;
  double sum_sq;
  double sum_sqrt;
};

// Per-thread accumulator record in shared memory: {sum, count} (16 B) or
// {sum_sq, count, sum_sqrt, pad} (32 B).  One predicated 128-bit load/store pair per
// in-range pair of points; a warp's records for one lag class are contiguous, i.e.
// bank-conflict free.  Written in PTX so that ptxas emits predicated LDS/STS instead
// of a branch per pair.  No "memory" clobber on purpose: records are only touched
// through these helpers between the two __syncthreads() that fence the accumulation.
struct HistRec16 { double sum; unsigned long long cnt; };
struct HistRec32 { double sq; unsigned long long cnt; double rt; unsigned long long pad; };

// Private per-thread accumulators, split layout (PRIV): sums[lag][thread] for all lag classes,
// then u32 counts[lag][thread], so a warp's read-modify-write of one lag moves 8+4 bytes per
// lane instead of a padded 16-byte record (shared-memory wavefronts are what bounds the sweep).  sum part per thread: STATS 1/2 one
// double, 3 two doubles {sum_sq, sum_sqrt}, 4 one u32 (below-count), 0 nothing.
template <int STATS>
struct PrivLayout {
  static constexpr int SUMB = STATS == 3 ? 16 : (STATS == 4 ? 4 : (STATS == 0 ? 0 : 8));
  static constexpr int LAGB = HB_THREADS * (SUMB + 4);  // bytes per lag class in total
};

// sqrt|dz| for the Cressie-Hawkins sum (estimators.py:123).  The correctly rounded library sqrt
// costs 9 FP64 instructions plus a range branch per pair; the sum only has to hold 1e-12, so:
// y = rsqrt.approx (MUFU.RSQ64H, ~2^-20 relative on the high word), then ONE third-order
// (Halley) step on the full operand: g = a y, r = 1 - g y, sqrt(a) ~ g + g r (1/2 + 3/8 r),
// error ~ 2.5 r^3 < 2^-58, i.e. the result is within an ulp.  6 FP64 instructions, no branch.
// The 1e-300 keeps a = 0 (tied observations) away from 0 * inf; it changes no a > 1e-284 and
// contributes 1e-150 otherwise.
__device__ __forceinline__ double sqrt_abs(double dz) {
  const double a = fabs(dz) + 1e-300;
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double g = __dmul_rn(a, y);
  const double r = fma(-g, y, 1.0);
  const double p = fma(0.375, r, 0.5);
  return fma(__dmul_rn(g, r), p, g);
}

template <int STATS>
__device__ __forceinline__ void rmw_split(unsigned saddr, unsigned caddr, double dz, int pred) {
  if (STATS == 3) {
    const double rt = sqrt_abs(dz);
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 tb, rb;\n\t.reg .f64 t, u;\n\t.reg .b32 c;\n\t"
        "setp.ne.s32 p, %4, 0;\n\t"
        "@p ld.shared.v2.b64 {tb, rb}, [%0];\n\t"
        "@p ld.shared.u32 c, [%1];\n\t"
        "mov.b64 t, tb;\n\tmov.b64 u, rb;\n\t"
        "fma.rn.f64 t, %2, %2, t;\n\t"
        "add.rn.f64 u, u, %3;\n\t"
        "add.u32 c, c, 1;\n\t"
        "mov.b64 tb, t;\n\tmov.b64 rb, u;\n\t"
        "@p st.shared.v2.b64 [%0], {tb, rb};\n\t"
        "@p st.shared.u32 [%1], c;\n\t}"
        :: "r"(saddr), "r"(caddr), "d"(dz), "d"(rt), "r"(pred));
  } else if (STATS == 2) {
    const double rt = sqrt_abs(dz);
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 tb;\n\t.reg .f64 t;\n\t.reg .b32 c;\n\t"
        "setp.ne.s32 p, %3, 0;\n\t"
        "@p ld.shared.b64 tb, [%0];\n\t"
        "@p ld.shared.u32 c, [%1];\n\t"
        "mov.b64 t, tb;\n\t"
        "add.rn.f64 t, t, %2;\n\t"
        "add.u32 c, c, 1;\n\t"
        "mov.b64 tb, t;\n\t"
        "@p st.shared.b64 [%0], tb;\n\t"
        "@p st.shared.u32 [%1], c;\n\t}"
        :: "r"(saddr), "r"(caddr), "d"(rt), "r"(pred));
  } else if (STATS == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 tb;\n\t.reg .f64 t;\n\t.reg .b32 c;\n\t"
        "setp.ne.s32 p, %3, 0;\n\t"
        "@p ld.shared.b64 tb, [%0];\n\t"
        "@p ld.shared.u32 c, [%1];\n\t"
        "mov.b64 t, tb;\n\t"
        "fma.rn.f64 t, %2, %2, t;\n\t"
        "add.u32 c, c, 1;\n\t"
        "mov.b64 tb, t;\n\t"
        "@p st.shared.b64 [%0], tb;\n\t"
        "@p st.shared.u32 [%1], c;\n\t}"
        :: "r"(saddr), "r"(caddr), "d"(dz), "r"(pred));
  } else if (STATS == 4) {
    // dz carries the 0/1 "below the per-lag |dz| threshold" flag
    const int below = dz != 0.0 ? 1 : 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 b, c;\n\t"
        "setp.ne.s32 p, %3, 0;\n\t"
        "@p ld.shared.u32 b, [%0];\n\t"
        "@p ld.shared.u32 c, [%1];\n\t"
        "add.u32 b, b, %2;\n\tadd.u32 c, c, 1;\n\t"
        "@p st.shared.u32 [%0], b;\n\t"
        "@p st.shared.u32 [%1], c;\n\t}"
        :: "r"(saddr), "r"(caddr), "r"(below), "r"(pred));
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 c;\n\t"
        "setp.ne.s32 p, %1, 0;\n\t"
        "@p ld.shared.u32 c, [%0];\n\t"
        "add.u32 c, c, 1;\n\t"
        "@p st.shared.u32 [%0], c;\n\t}"
        :: "r"(caddr), "r"(pred));
  }
}

// Multi-column values (VM = true): `values` is [nv][n]; the per-pair difference is
//   SKG_VALUE_CROSS     |dz_0| * |dz_1|             (Variogram.py:1979-1984, cross-variogram)
//   SKG_VALUE_AGGREGATE sqrt(sum_k dz_k^2)          (Variogram.py:1974-1977, multivariate='aggregate')
// formed with the reference's operand order (no FMA), i.e. the same bits as its diffs array.
// Both tiles' value columns live in shared memory (mvI[k][TILE], mvJ[k][TILE + 2]).
struct MvSpec {
  int nv;
  int mode;    // SKG_VALUE_* without flags
  int square;  // SKG_VALUE_SQUARE_FLAG: the difference is squared before the statistics
};

__device__ __forceinline__ double mv_diff(const double* __restrict__ mvI, const double* __restrict__ mvJ,
                                          int i, int jj, const MvSpec mv) {
  if (mv.mode == SKG_VALUE_SCALAR) return fabs(mvI[i] - mvJ[jj]);  // only reached with the SQUARE flag
  if (mv.mode == SKG_VALUE_CROSS) {
    const double a = fabs(mvI[i] - mvJ[jj]);
    const double b = fabs(mvI[TILE + i] - mvJ[(TILE + 2) + jj]);
    return __dmul_rn(a, b);
  }
  double acc = 0.0;
  for (int k = 0; k < mv.nv; ++k) {
    const double d = mvI[k * TILE + i] - mvJ[k * (TILE + 2) + jj];
    acc = __dadd_rn(acc, __dmul_rn(d, d));
  }
  return sqrt(acc);
}

constexpr int HB = HB_THREADS;   // threads per CTA
constexpr int HR = TILE / HB;   // i-points per thread

template <int EMAX>
__device__ __forceinline__ int lag_of(uint64_t sb, const uint64_t* sThr, const uint8_t* sLutBiased,
                                      int shift, int klo, int khi, int emax) {
  // cell = clamp(hi32 >> shift, klo, khi); sLutBiased = sLut - klo
  // Non-negative doubles order like their bit patterns, so the threshold test is done
  // as ONE fp64 compare (the FP64 pipe has slack, the integer pipe does not); NaN (padded
  // or excluded pairs) compares false and lands in class nb through the top LUT cell.
  const int k = (int)((uint32_t)(sb >> 32) >> shift);
  int g = sLutBiased[min(max(k, klo), khi)];
  const double s = __longlong_as_double((long long)sb);
  const double* dThr = reinterpret_cast<const double*>(sThr);
  if (EMAX > 0) {
#pragma unroll
    for (int e = 0; e < EMAX; ++e) g += (s >= dThr[g]) ? 1 : 0;
  } else {
    // many thresholds may share one cell (clustered edges): stop at the first one above s
    for (int e = 0; e < emax; ++e) {
      if (!(s >= dThr[g])) break;
      ++g;
    }
  }
  return g;
}

// Lag class times 8 (= byte offset of the lag's threshold), straight from a u16 LUT that
// stores lut[cell] * 8: the threshold load needs no index scaling, and both accumulator
// addresses are one multiply-add each from the result (thread base folded in).
template <int EMAX>
__device__ __forceinline__ unsigned lag_off8(uint64_t sb, const uint64_t* sThr,
                                             const unsigned char* lutBytesBiased, int shift,
                                             int klo, int khi, int emax) {
  // lutBytesBiased = (bytes of the u16 LUT) - 2 * klo
  const int k = (int)((uint32_t)(sb >> 32) >> shift);
  unsigned g8 = *reinterpret_cast<const uint16_t*>(lutBytesBiased + 2 * min(max(k, klo), khi));
  const double s = __longlong_as_double((long long)sb);
  const char* thrBytes = reinterpret_cast<const char*>(sThr);
  if (EMAX > 0) {
#pragma unroll
    for (int e = 0; e < EMAX; ++e)
      g8 += (s >= *reinterpret_cast<const double*>(thrBytes + g8)) ? 8u : 0u;
  } else {
    // many thresholds may share one cell (clustered edges): stop at the first one above s
    for (int e = 0; e < emax; ++e) {
      if (!(s >= *reinterpret_cast<const double*>(thrBytes + g8))) break;
      g8 += 8u;
    }
  }
  return g8;
}

// PRIV: private per-thread records + software-pipelined loop (front end of pair set
// jj+1 is issued before the read-modify-writes of pair set jj).
// !PRIV (n_lags too large for private records): one shared record per lag + atomics.
template <int DIM, int DIR, int STATS, int EMAX, bool PRIV, bool VM = false>
__global__ void __launch_bounds__(HB)
pair_hist_kernel(const double* __restrict__ coords, const double* __restrict__ values, int64_t n,
                 const __grid_constant__ BinTable tab, const __grid_constant__ DirParams dp,
                 HistPartial* __restrict__ partial, const unsigned int* __restrict__ tile_list,
                 const unsigned long long* __restrict__ tile_count,
                 const double* __restrict__ bbox, const double* __restrict__ bbox32, double wlo,
                 double whi, const __grid_constant__ DzTable dzt, int js, const MvSpec mv,
                 unsigned long long* __restrict__ whist, long long wbuckets) {
  __shared__ double2 sXY[TILE + 2];
  __shared__ double sW[DIM == 3 ? TILE + 2 : 1];
  __shared__ double sZ[TILE + 2];
  __shared__ uint64_t sThr[THR_MAX];
  __shared__ uint16_t sLut[LUT_MAX];  // lut[cell] * 8
  __shared__ double sDzLo[STATS == 4 ? THR_MAX : 1];  // per-lag |dz| thresholds (below-count mode)
  __shared__ double sDzHi[STATS == 4 ? THR_MAX : 1];
  __shared__ double sDzSc[STATS == 4 ? THR_MAX : 1];
  extern __shared__ __align__(16) unsigned char acc_raw[];

  constexpr int REC = (STATS == 3) ? 32 : 16;              // shared records of the !PRIV path
  using PL = PrivLayout<STATS>;
  constexpr int LAGB = PRIV ? PL::LAGB : REC;               // bytes per lag class
  const int nb = tab.nb;
  const int tid = threadIdx.x;
  const double qnan = __longlong_as_double(NAN_BITS);

  for (int k = tid; k < THR_MAX; k += HB) sThr[k] = tab.thr[k];
  for (int k = tid; k < tab.ncell; k += HB) sLut[k] = (uint16_t)(tab.lut[k] * 8);
  if (STATS == 4) {
    for (int k = tid; k < THR_MAX; k += HB) {
      sDzLo[k] = k < nb ? dzt.lo[k] : 0.0;
      sDzHi[k] = k < nb ? dzt.hi[k] : 0.0;
      sDzSc[k] = k < nb ? dzt.scale[k] : 0.0;
    }
  }
  {
    unsigned long long* z = reinterpret_cast<unsigned long long*>(acc_raw);
    for (int k = tid; k < nb * (LAGB / 8); k += HB) z[k] = 0ull;
  }
  __syncthreads();

  const int shift = tab.shift, klo = tab.key0, khi = tab.key0 + tab.ncell - 1, emax = tab.emax;
  const unsigned char* lutb = reinterpret_cast<const unsigned char*>(sLut) - 2 * klo;
  constexpr int GS = 8;  // lag ids are carried times 8 (threshold byte offsets)
  const unsigned nbS = (unsigned)nb * GS;
  const unsigned aBase = (unsigned)__cvta_generic_to_shared(acc_raw);
  const unsigned aSum = aBase + tid * PL::SUMB;
  const unsigned aCnt = aBase + (unsigned)nb * (HB * PL::SUMB) + tid * 4;  // counts behind all sums
  // multi-column values: both tiles' columns behind the records
  double* mvI = reinterpret_cast<double*>(acc_raw + (size_t)nb * LAGB);
  double* mvJ = mvI + (VM ? mv.nv * TILE : 0);
  const int64_t nt = (n + TILE - 1) / TILE;
  const double* __restrict__ X = coords;
  const double* __restrict__ Y = coords + n;
  const double* __restrict__ W = coords + 2 * n;

  // Work item = (tile, j-chunk): each tile's j range is cut into `js` chunks so that the
  // static round-robin over the grid stays balanced when there are few tiles per CTA.
  const int jchunk = TILE / js;
  const int64_t nwork = (int64_t)(*tile_count);  // surviving (tile, j-chunk) work items
  for (int64_t w = blockIdx.x; w < nwork; w += gridDim.x) {
    const unsigned int item = tile_list[w];
    const int64_t t = item / (unsigned)js;
    const int jbeg = (int)(item % (unsigned)js) * jchunk;
    int I, J;
    tile_decode(t, nt, I, J);
    // warp-level culling: this warp's four 32-point sub-blocks against the chunk box
    bool wact = false;
    {
      double jb[6];
      chunk_box<DIM>(bbox32, J, jbeg / jchunk, js, jb);
#pragma unroll
      for (int r = 0; r < HR; ++r) {
        double a, b;
        const int64_t sub = (int64_t)I * 8 + r * (HB / 32) + (tid >> 5);
        subblock_range<DIM>(bbox32, sub, jb, a, b);
        bool on = (a * (1.0 - 1e-12) < whi && b * (1.0 + 1e-12) >= wlo);
        if (DIR && DIM == 2 && on) on = dir_box_may_pass(bbox32 + sub * 6, jb, dp);
        wact |= on;
      }
    }
    const int64_t jn = n - ((int64_t)J * TILE + jbeg);
    const int jcount = jn < jchunk ? (jn > 0 ? (int)jn : 0) : jchunk;
    double xi[HR], yi[HR], wi[HR], zi[HR];
#pragma unroll
    for (int r = 0; r < HR; ++r) {
      const int64_t i = (int64_t)I * TILE + r * HB + tid;
      const bool v = i < n;
      xi[r] = v ? X[i] : qnan;
      yi[r] = v ? Y[i] : qnan;
      wi[r] = (DIM == 3) ? (v ? W[i] : qnan) : 0.0;
      zi[r] = (v && !VM) ? values[i] : 0.0;
    }
    __syncthreads();  // every read of the previous j chunk is done
    for (int k = tid; k < jcount + 2; k += HB) {  // two NaN entries pad the pipeline tail
      const int64_t j = (int64_t)J * TILE + jbeg + k;
      const bool v = k < jcount;
      sXY[k] = make_double2(v ? X[j] : qnan, v ? Y[j] : qnan);
      if (DIM == 3) sW[k] = v ? W[j] : qnan;
      sZ[k] = (v && !VM) ? values[j] : 0.0;
    }
    if (VM) {
      for (int c = 0; c < mv.nv; ++c) {
        for (int k = tid; k < TILE; k += HB) {
          const int64_t i = (int64_t)I * TILE + k;
          mvI[c * TILE + k] = i < n ? values[(int64_t)c * n + i] : 0.0;
        }
        for (int k = tid; k < jcount + 2; k += HB) {
          const int64_t j = (int64_t)J * TILE + jbeg + k;
          mvJ[c * (TILE + 2) + k] = k < jcount ? values[(int64_t)c * n + j] : 0.0;
        }
      }
    }
    __syncthreads();
    if (jcount == 0 || !wact) continue;  // warp-uniform; both barriers are already passed

    auto body = [&](auto diag_tag) {
      constexpr bool DIAG = decltype(diag_tag)::value;
      // front end of pair set jj: lag classes g[] (nb = not accumulated) and dz[]
      auto front = [&](int jj, unsigned (&g)[HR], double (&dz)[HR]) {
        const double2 pj = sXY[jj];
        const double zj = sZ[jj];
        const double wj = (DIM == 3) ? sW[jj] : 0.0;
#pragma unroll
        for (int r = 0; r < HR; ++r) {
          const double dx = xi[r] - pj.x;
          const double dy = yi[r] - pj.y;
          double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
          if (DIM == 3) {
            const double dw = wi[r] - wj;
            s = __dadd_rn(s, __dmul_rn(dw, dw));
          }
          const uint64_t sb = (uint64_t)__double_as_longlong(s);
          unsigned gg = lag_off8<EMAX>(sb, sThr, lutb, shift, klo, khi, emax);
          if (DIAG) gg = (jbeg + jj <= r * HB + tid) ? nbS : gg;  // keep i < j only
          if (DIR) {
            if (gg < nbS && !dir_mask_t<DIR == 2>(dx, dy, s, dp)) gg = nbS;
          }
          g[r] = gg;
          if (VM) {
            // SKG_VALUE_DIST_*: the "difference" is a power of the distance itself, so the
            // estimator sums become moments of the distance distribution (histogram-rule binners)
            if (mv.mode == SKG_VALUE_DIST_S) dz[r] = s;
            else if (mv.mode == SKG_VALUE_DIST_D) dz[r] = sqrt(s);
            else if (mv.mode == SKG_VALUE_DIST_D15) { const double d = sqrt(s); dz[r] = d * sqrt(d); }
            else dz[r] = mv_diff(mvI, mvJ, r * HB + tid, jj, mv);
            if (mv.square) dz[r] = __dmul_rn(dz[r], dz[r]);  // sum_sqrt then is the sum of |difference|
          } else {
            dz[r] = zi[r] - zj;
          }
          if (STATS == 4) {
            const double ad = fabs(dz[r]);
            const unsigned lg = min(gg, nbS - GS) / GS;
            const double lo = sDzLo[lg];
            // the few in-window keys also go to the fine histogram of the median select (one
            // sweep instead of a count sweep plus a histogram sweep)
            if (whist && gg < nbS && ad >= lo && ad < sDzHi[lg]) {
              long long bk = __double2ll_rz(__dmul_rn(__dsub_rn(ad, lo), sDzSc[lg]));
              bk = bk < wbuckets - 1 ? bk : wbuckets - 1;
              atomicAdd(&whist[(long long)lg * wbuckets + bk], 1ull);
            }
            dz[r] = (ad < lo) ? 1.0 : 0.0;
          }
        }
      };
      auto accumulate = [&](const unsigned (&g)[HR], const double (&dz)[HR]) {
#pragma unroll
        for (int r = 0; r < HR; ++r) {
          const int in = g[r] < nbS ? 1 : 0;
          if (PRIV) {
            rmw_split<STATS>(aSum + g[r] * (unsigned)(PL::SUMB * HB / 8), aCnt + g[r] * (unsigned)(HB * 4 / 8),
                             dz[r], in);
          } else if (in) {
            unsigned char* rec = acc_raw + (g[r] / GS) * REC;
            atomicAdd(reinterpret_cast<unsigned long long*>(rec + 8), 1ull);
            if (STATS & SKG_STAT_SUM_SQ) atomicAdd(reinterpret_cast<double*>(rec), dz[r] * dz[r]);
            if (STATS & SKG_STAT_SUM_SQRT)
              atomicAdd(reinterpret_cast<double*>(rec + (STATS == 3 ? 16 : 0)), sqrt(fabs(dz[r])));
          }
        }
      };
      unsigned gA[HR], gB[HR];
      double dA[HR], dB[HR];
      front(0, gA, dA);
      int jj = 1;
      for (; jj + 1 <= jcount; jj += 2) {  // entries >= jcount are NaN-padded -> class nb
        front(jj, gB, dB);
        accumulate(gA, dA);
        front(jj + 1, gA, dA);
        accumulate(gB, dB);
      }
      if (jj <= jcount) {
        front(jj, gB, dB);
        accumulate(gA, dA);
      }
    };
    if (I == J) body(std::true_type{}); else body(std::false_type{});
  }
  __syncthreads();

  // fixed-order block reduction of the per-thread records
  HistPartial* out = partial + (int64_t)blockIdx.x * nb;
  if (PRIV) {
    const int warp = tid >> 5, lane = tid & 31;
    for (int gI = warp; gI < nb; gI += HB / 32) {
      unsigned long long c = 0;
      double a = 0.0, b = 0.0;
#pragma unroll
      for (int k = 0; k < HB / 32; ++k) {
        const int th = lane + 32 * k;
        const unsigned char* sp = acc_raw + (size_t)(gI * HB + th) * PL::SUMB;
        c += *reinterpret_cast<const unsigned int*>(acc_raw + (size_t)nb * (HB * PL::SUMB) +
                                                    (size_t)(gI * HB + th) * 4);  // counts < 2^32
        if (STATS == 4) a += (double)*reinterpret_cast<const unsigned int*>(sp);
        else if (STATS & SKG_STAT_SUM_SQ) a += *reinterpret_cast<const double*>(sp);
        if (STATS == 2) b += *reinterpret_cast<const double*>(sp);
        if (STATS == 3) b += *reinterpret_cast<const double*>(sp + 8);
      }
      c = warp_sum(c);
      if ((STATS & SKG_STAT_SUM_SQ) || STATS == 4) a = warp_sum(a);
      if ((STATS & SKG_STAT_SUM_SQRT) && STATS != 4) b = warp_sum(b);
      if (lane == 0) {
        out[gI].count = c;
        out[gI].sum_sq = a;
        out[gI].sum_sqrt = b;
      }
    }
  } else {
    for (int gI = tid; gI < nb; gI += HB) {
      const unsigned char* rec = acc_raw + (size_t)gI * REC;
      out[gI].count = *reinterpret_cast<const unsigned long long*>(rec + 8);
      out[gI].sum_sq = (STATS & SKG_STAT_SUM_SQ) ? *reinterpret_cast<const double*>(rec) : 0.0;
      out[gI].sum_sqrt = (STATS & SKG_STAT_SUM_SQRT)
                             ? *reinterpret_cast<const double*>(rec + (STATS == 3 ? 16 : 0)) : 0.0;
    }
  }
}

// ------------------------------------------------- K1/K3: select passes ---

// Same sweep structure as the fused kernel (64 threads x 4 "i" points, static tables,
// j-chunk work items, software-pipelined front end).  Histogram pass: one u64 atomicAdd
// (L2) per key INSIDE the segment's window - the host keeps the windows narrow (the class
// that holds a wanted rank), so most pairs only pay the front end.  Gather pass: bitmap
// test and append.  The per-thread running maximum of the keys feeds atomicMax once.
template <int DIM, int DIR, bool SEG, int KEY, bool GATHER, int EMAX, bool VM = false>
__global__ void __launch_bounds__(HB)
pair_select_kernel(const double* __restrict__ coords, const double* __restrict__ values, int64_t n,
                   const __grid_constant__ BinTable tab, const __grid_constant__ DirParams dp,
                   const __grid_constant__ SegTable seg_tab, long long nbuckets,
                   unsigned long long* __restrict__ hist, unsigned long long* __restrict__ max_bits,
                   const uint32_t* __restrict__ bitmap, double* __restrict__ cand_key,
                   uint32_t* __restrict__ cand_slot, unsigned long long* __restrict__ cand_count,
                   long long capacity, const unsigned int* __restrict__ tile_list,
                   const unsigned long long* __restrict__ tile_count,
                   const double* __restrict__ bbox, const double* __restrict__ bbox32, double wlo,
                   double whi, int js, const MvSpec mv, const double* __restrict__ bbox8,
                   int fast_seg, double fast_lo, double fast_hi) {
  // fast_seg >= 0 (distance keys, exactly ONE segment with a window, e.g. the gather of a median):
  // a pair belongs to segment fast_seg iff fast_lo <= s < fast_hi, everything else is skipped without
  // the lag-class lookup - the sweep then costs the distance and two compares per pair
  __shared__ double2 sXY[TILE + 2];
  __shared__ double sW[DIM == 3 ? TILE + 2 : 1];
  __shared__ double sZ[TILE + 2];
  __shared__ uint64_t sThr[THR_MAX];
  __shared__ uint8_t sLut[LUT_MAX];
  __shared__ SegTable sSeg;
  __shared__ double sGB[DIM == 2 ? TILE / SEL_GROUP : 1][4];  // boxes of the chunk's 16-point j groups
  // distance keys with lag classes: the (few) windows as high words of their bounds, for the
  // group-level test (integer compares on truncated bounds: a superset, never a miss)
  __shared__ uint32_t sWinLo[SEL_MAXWIN], sWinHi[SEL_MAXWIN];
  __shared__ int sNWin;
  extern __shared__ __align__(16) unsigned char mv_raw[];  // VM: value columns of both tiles
  double* mvI = reinterpret_cast<double*>(mv_raw);
  double* mvJ = mvI + (VM ? mv.nv * TILE : 0);
  const int tid = threadIdx.x;
  const int nb = SEG ? tab.nb : 1;
  const double qnan = __longlong_as_double(NAN_BITS);
  for (int k = tid; k < THR_MAX; k += HB) sThr[k] = tab.thr[k];
  for (int k = tid; k < tab.ncell; k += HB) sLut[k] = tab.lut[k];
  for (int k = tid; k < nb; k += HB) {
    sSeg.lo_bits[k] = seg_tab.lo_bits[k];
    sSeg.hi_bits[k] = seg_tab.hi_bits[k];
    sSeg.scale[k] = seg_tab.scale[k];
  }
  if (tid == 0) {
    int nw = 0;
    if (SEG && KEY == SKG_KEY_SQDIST) {
      for (int k = 0; k < nb; ++k) {
        const uint64_t wl = seg_tab.lo_bits[k], wh = seg_tab.hi_bits[k];
        if (wh <= wl) continue;
        if (nw == SEL_MAXWIN) { nw = -1; break; }   // too many windows: class look-ups instead
        sWinLo[nw] = (uint32_t)(wl >> 32);
        sWinHi[nw] = (uint32_t)((wh - 1) >> 32);
        ++nw;
      }
    } else {
      nw = -1;
    }
    sNWin = nw;
  }
  __syncthreads();
  const int nwin = sNWin;
  const int shift = tab.shift, klo = tab.key0, khi = tab.key0 + tab.ncell - 1, emax = tab.emax;
  const uint8_t* lutb = sLut - klo;
  const int64_t nt = (n + TILE - 1) / TILE;
  const double* __restrict__ X = coords;
  const double* __restrict__ Y = coords + n;
  const double* __restrict__ W = coords + 2 * n;
  unsigned long long local_max = 0ull;
  const int jchunk = TILE / js;
  const int64_t nwork = (int64_t)(*tile_count);

  for (int64_t w = blockIdx.x; w < nwork; w += gridDim.x) {
    const unsigned int item = tile_list[w];
    const int64_t t = item / (unsigned)js;
    const int jbeg = (int)(item % (unsigned)js) * jchunk;
    int I, J;
    tile_decode(t, nt, I, J);
    bool wact = false;
    {
      double jb[6];
      chunk_box<DIM>(bbox32, J, jbeg / jchunk, js, jb);
#pragma unroll
      for (int r = 0; r < HR; ++r) {
        double a, b;
        const int64_t sub = (int64_t)I * 8 + r * (HB / 32) + (tid >> 5);
        subblock_range<DIM>(bbox32, sub, jb, a, b);
        bool on = (a * (1.0 - 1e-12) < whi && b * (1.0 + 1e-12) >= wlo);
        if (DIR && DIM == 2 && on) on = dir_box_may_pass(bbox32 + sub * 6, jb, dp);
        wact |= on;
      }
    }
    const int64_t jn = n - ((int64_t)J * TILE + jbeg);
    const int jcount = jn < jchunk ? (jn > 0 ? (int)jn : 0) : jchunk;
    double xi[HR], yi[HR], wi[HR], zi[HR];
#pragma unroll
    for (int r = 0; r < HR; ++r) {
      const int64_t i = (int64_t)I * TILE + r * HB + tid;
      const bool v = i < n;
      xi[r] = v ? X[i] : qnan;
      yi[r] = v ? Y[i] : qnan;
      wi[r] = (DIM == 3) ? (v ? W[i] : qnan) : 0.0;
      zi[r] = (v && values && !VM) ? values[i] : 0.0;
    }
    __syncthreads();
    for (int k = tid; k < jcount + 2; k += HB) {
      const int64_t j = (int64_t)J * TILE + jbeg + k;
      const bool v = k < jcount;
      sXY[k] = make_double2(v ? X[j] : qnan, v ? Y[j] : qnan);
      if (DIM == 3) sW[k] = v ? W[j] : qnan;
      sZ[k] = (v && values && !VM) ? values[j] : 0.0;
    }
    if (VM) {
      for (int c = 0; c < mv.nv; ++c) {
        for (int k = tid; k < TILE; k += HB) {
          const int64_t i = (int64_t)I * TILE + k;
          mvI[c * TILE + k] = i < n ? values[(int64_t)c * n + i] : 0.0;
        }
        for (int k = tid; k < jcount + 2; k += HB) {
          const int64_t j = (int64_t)J * TILE + jbeg + k;
          mvJ[c * (TILE + 2) + k] = k < jcount ? values[(int64_t)c * n + j] : 0.0;
        }
      }
    }
    const bool grouped = DIM == 2 && bbox8 != nullptr && I != J;  // group-level culling (below)
    if (grouped) {
      for (int g = tid; g < jchunk / SEL_GROUP; g += HB) {
        const int64_t b8 = ((int64_t)J * TILE + jbeg) / 8 + (int64_t)g * (SEL_GROUP / 8);
        double lx = bbox8[b8 * 4], ly = bbox8[b8 * 4 + 1], hx = bbox8[b8 * 4 + 2], hy = bbox8[b8 * 4 + 3];
        bool bad = !(lx == lx);
#pragma unroll
        for (int u = 1; u < SEL_GROUP / 8; ++u) {
          const double l2 = bbox8[(b8 + u) * 4];
          bad |= !(l2 == l2);
          lx = fmin(lx, l2);
          ly = fmin(ly, bbox8[(b8 + u) * 4 + 1]);
          hx = fmax(hx, bbox8[(b8 + u) * 4 + 2]);
          hy = fmax(hy, bbox8[(b8 + u) * 4 + 3]);
        }
        if (bad) lx = ly = hx = hy = qnan;  // NaN / missing point: the group is always swept
        sGB[DIM == 2 ? g : 0][0] = lx; sGB[DIM == 2 ? g : 0][1] = ly;
        sGB[DIM == 2 ? g : 0][2] = hx; sGB[DIM == 2 ? g : 0][3] = hy;
      }
    }
    __syncthreads();
    if (jcount == 0 || !wact) continue;

    // Group-level culling: can any pair of (this warp's points) x (16 consecutive j points) fall
    // into a window?  [smin, smax] = point-to-box range of the squared distance, formed with the
    // pair arithmetic (rigorous by monotone rounding).  Distance keys: some window must meet the
    // range; |dz| keys: some lag class met by the range must have a window at all.
    auto group_needed = [&](int g0) -> bool {
      if (!grouped) return true;
      if (!SEG && !GATHER && max_bits) return true;  // the running maximum looks at every pair
      const int gi = g0 / SEL_GROUP;
      const double lox = sGB[DIM == 2 ? gi : 0][0], loy = sGB[DIM == 2 ? gi : 0][1];
      const double hix = sGB[DIM == 2 ? gi : 0][2], hiy = sGB[DIM == 2 ? gi : 0][3];
      bool need = false;
#pragma unroll
      for (int r = 0; r < HR; ++r) {
        const double ax = xi[r] - lox, bx = xi[r] - hix, ay = yi[r] - loy, by = yi[r] - hiy;
        const double ax2 = __dmul_rn(ax, ax), bx2 = __dmul_rn(bx, bx);
        const double ay2 = __dmul_rn(ay, ay), by2 = __dmul_rn(by, by);
        const bool gx = ax2 > bx2, gy = ay2 > by2;
        const double fx = gx ? ax2 : bx2, fy = gy ? ay2 : by2;
        double nx = gx ? bx2 : ax2, ny = gy ? by2 : ay2;
        if ((__double2hiint(ax) ^ __double2hiint(bx)) < 0) nx = 0.0;
        if ((__double2hiint(ay) ^ __double2hiint(by)) < 0) ny = 0.0;
        const double smin = __dadd_rn(nx, ny), smax = __dadd_rn(fx, fy);
        if (!(smax == smax) || !(smin == smin)) { need = true; continue; }  // flagged box / padded point
        const uint64_t bmin = (uint64_t)__double_as_longlong(smin), bmax = (uint64_t)__double_as_longlong(smax);
        if (!SEG) {
          if (KEY == SKG_KEY_SQDIST) need |= (bmin < sSeg.hi_bits[0]) & (bmax >= sSeg.lo_bits[0]);
          else need |= sSeg.hi_bits[0] > sSeg.lo_bits[0];
        } else if (KEY == SKG_KEY_SQDIST && nwin >= 0) {
          const uint32_t a = (uint32_t)(bmin >> 32), b = (uint32_t)(bmax >> 32);
          bool hit = false;
          for (int w = 0; w < nwin; ++w) hit |= (a <= sWinHi[w]) & (b >= sWinLo[w]);
          need |= hit;
        } else {
          const int c0 = lag_of<EMAX>(bmin, sThr, lutb, shift, klo, khi, emax);
          int c1 = lag_of<EMAX>(bmax, sThr, lutb, shift, klo, khi, emax);
          c1 = c1 < nb - 1 ? c1 : nb - 1;
          if (c1 - c0 > 3) { need = true; continue; }
          for (int c = c0; c <= c1; ++c) {
            const uint64_t wl = sSeg.lo_bits[c], wh = sSeg.hi_bits[c];
            if (KEY == SKG_KEY_SQDIST) need |= (wh > wl) & (bmin < wh) & (bmax >= wl);
            else need |= wh > wl;
          }
        }
      }
      return __any_sync(0xffffffffu, need);
    };

    auto body = [&](auto diag_tag, int j0, int j1) {
      constexpr bool DIAG = decltype(diag_tag)::value;
      auto front = [&](int jj, int (&g)[HR], double (&key)[HR]) {
        const double2 pj = sXY[jj];
        const double zj = sZ[jj];
        const double wj = (DIM == 3) ? sW[jj] : 0.0;
#pragma unroll
        for (int r = 0; r < HR; ++r) {
          const double dx = xi[r] - pj.x;
          const double dy = yi[r] - pj.y;
          double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
          if (DIM == 3) {
            const double dw = wi[r] - wj;
            s = __dadd_rn(s, __dmul_rn(dw, dw));
          }
          const uint64_t sb = (uint64_t)__double_as_longlong(s);
          int gg;
          if (KEY == SKG_KEY_SQDIST && fast_seg >= 0) gg = (s >= fast_lo && s < fast_hi) ? fast_seg : nb;
          else if (SEG) gg = lag_of<EMAX>(sb, sThr, lutb, shift, klo, khi, emax);
          else gg = sb <= INF_BITS ? 0 : nb;  // NaN: padded point or NaN coordinate
          if (DIAG) gg = (jbeg + jj <= r * HB + tid) ? nb : gg;  // keep i < j only
          if (DIR) {
            // the mask costs more than the rest of the pair: evaluate it only where the pair can
            // still count (its segment has a window and, for distance keys, s lies inside it);
            // the running maximum of the one-segment histogram pass looks at every masked pair
            bool cand = gg < nb;
            if ((SEG || GATHER) && cand) {
              const uint64_t wl = sSeg.lo_bits[gg], wh = sSeg.hi_bits[gg];
              cand = (KEY == SKG_KEY_SQDIST) ? (sb >= wl && sb < wh) : (wh > wl);
            }
            if (!cand || !dir_mask_t<DIR == 2>(dx, dy, s, dp)) gg = nb;
          }
          g[r] = gg;
          key[r] = (KEY == SKG_KEY_SQDIST) ? s
                   : (VM ? mv_diff(mvI, mvJ, r * HB + tid, jj, mv) : fabs(zi[r] - zj));
        }
      };
      auto accumulate = [&](const int (&g)[HR], const double (&key)[HR]) {
        if (KEY == SKG_KEY_SQDIST && fast_seg >= 0) {  // hits are rare: skip the whole step when the warp has none
          bool any = false;
#pragma unroll
          for (int r = 0; r < HR; ++r) any |= g[r] < nb;
          if (!__any_sync(0xffffffffu, any)) return;
        }
        // window test for all HR pairs first, branch-free (most in-range keys lie outside the
        // windows: only the few that hit take the bucket / atomic path below)
        bool inw[HR];
        uint64_t wlo_b[HR];
#pragma unroll
        for (int r = 0; r < HR; ++r) {
          const int gs = g[r] < nb ? g[r] : 0;
          const uint64_t kb = (uint64_t)__double_as_longlong(key[r]);
          wlo_b[r] = sSeg.lo_bits[gs];
          const uint64_t whi_b = sSeg.hi_bits[gs];
          const bool ok = (g[r] < nb) & (kb <= INF_BITS);   // NaN value: never counted
          if (!SEG && !GATHER) local_max = (ok && kb > local_max) ? kb : local_max;
          inw[r] = ok & (kb >= wlo_b[r]) & (kb < whi_b);
        }
#pragma unroll
        for (int r = 0; r < HR; ++r) {
          if (!inw[r]) continue;
          const double lo = __longlong_as_double((long long)wlo_b[r]);
          long long bk = __double2ll_rz(__dmul_rn(__dsub_rn(key[r], lo), sSeg.scale[g[r]]));
          bk = bk < nbuckets - 1 ? bk : nbuckets - 1;
          const long long slot = (long long)g[r] * nbuckets + bk;
          if (!GATHER) {
            atomicAdd(&hist[slot], 1ull);
          } else if ((bitmap[slot >> 5] >> (slot & 31)) & 1u) {
            const unsigned long long pos = atomicAdd(cand_count, 1ull);
            if ((long long)pos < capacity) {
              cand_key[pos] = key[r];
              cand_slot[pos] = (uint32_t)slot;
            }
          }
        }
      };
      int gA[HR], gB[HR];
      double kA[HR], kB[HR];
      front(j0, gA, kA);
      int jj = j0 + 1;
      for (; jj + 1 <= j1; jj += 2) {  // the entry at j1 is a later point or NaN padding: never used
        front(jj, gB, kB);
        accumulate(gA, kA);
        front(jj + 1, gA, kA);
        accumulate(gB, kB);
      }
      if (jj <= j1) {
        front(jj, gB, kB);
        accumulate(gA, kA);
      }
    };
    if (I == J) {
      body(std::true_type{}, 0, jcount);
    } else {
      // consecutive needed groups are swept as one range (keeps the software pipeline long)
      int run0 = -1;
      for (int g0 = 0;; g0 += SEL_GROUP) {
        const bool need = g0 < jcount && group_needed(g0);
        if (need && run0 < 0) run0 = g0;
        if (!need && run0 >= 0) { body(std::false_type{}, run0, min(g0, jcount)); run0 = -1; }
        if (g0 >= jcount) break;
      }
    }
  }
  if (!SEG && !GATHER && max_bits) {
    unsigned long long m = local_max;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      unsigned long long other = __shfl_down_sync(0xffffffffu, m, o);
      m = other > m ? other : m;
    }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(max_bits, m);
  }
}

// ------------------------------------------------------ launch plumbing ---

constexpr int MAX_GRID_PER_SM = 16;

inline size_t mv_smem_bytes(const MvSpec& mv) {
  return mv.nv >= 1 ? (size_t)mv.nv * (2 * TILE + 2) * sizeof(double) : 0;
}

struct HistLaunch {
  const double* coords; const double* values; int64_t n;
  const BinTable* tab; const DirParams* dp; HistPartial* partial;
  const unsigned int* list; const unsigned long long* list_count;
  const double* bbox; const double* bbox32; double wlo, whi; const DzTable* dzt;
  int max_blocks; int64_t ntiles; int js; size_t smem; cudaStream_t st; int* nblocks_out;
  MvSpec mv;
  const double* bbox8;
  unsigned long long* whist;  // STATS == 4: optional in-window |dz| histogram [n_lags][wbuckets]
  long long wbuckets;
  const DzTable* dzt_dev;     // the |dz| window table in device memory (window kernel, STATS == 4)
};

template <int DIM, int DIR, int STATS, int EMAX, bool PRIV, bool VM>
static int launch_hist(const HistLaunch& L) {
  auto kern = pair_hist_kernel<DIM, DIR, STATS, EMAX, PRIV, VM>;
  static thread_local size_t cached_smem = (size_t)-1;
  static thread_local int cached_occ = 0;
  if (cached_smem != L.smem) {
    SKG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
    SKG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cached_occ, kern, HB, L.smem));
    cached_smem = L.smem;
  }
  int occ = cached_occ;
  if (occ < 1) { set_error("pair_hist kernel does not fit (smem=%zu)", L.smem); return SKG_E_UNSUPPORTED; }
  occ = occ < MAX_GRID_PER_SM ? occ : MAX_GRID_PER_SM;
  const int js = L.js;
  int64_t grid = (int64_t)sm_count() * occ;
  if (grid > L.ntiles * js) grid = L.ntiles * js;
  if (grid > L.max_blocks) grid = L.max_blocks;
  *L.nblocks_out = (int)grid;
  kern<<<(unsigned)grid, HB, L.smem, L.st>>>(L.coords, L.values, L.n, *L.tab, *L.dp, L.partial,
                                             L.list, L.list_count, L.bbox, L.bbox32, L.wlo, L.whi,
                                             *L.dzt, js, L.mv, L.whist, L.wbuckets);
  SKG_CUDA(cudaGetLastError());
  return 0;
}

struct SelectLaunch {
  const double* coords; const double* values; int64_t n;
  const BinTable* tab; const DirParams* dp; const SegTable* seg; int64_t nbuckets;
  uint64_t* hist; uint64_t* max_bits; const uint32_t* bitmap;
  double* cand_key; uint32_t* cand_slot; uint64_t* cand_count; int64_t capacity;
  int64_t ntiles; const unsigned int* list; const unsigned long long* list_count;
  const double* bbox; const double* bbox32; double wlo, whi; int js; cudaStream_t st;
  MvSpec mv;
  const double* bbox8;
  int fast_seg; double fast_lo, fast_hi;  // single-window distance select (see pair_select_kernel)
};

template <int DIM, int DIR, bool SEG, int KEY, bool GATHER, int EMAX, bool VM>
static int launch_select(const SelectLaunch& L) {
  auto kern = pair_select_kernel<DIM, DIR, SEG, KEY, GATHER, EMAX, VM>;
  const size_t smem = VM ? mv_smem_bytes(L.mv) : 0;
  int occ = 0;
  SKG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, HB, smem));
  occ = occ < 1 ? 1 : (occ < MAX_GRID_PER_SM ? occ : MAX_GRID_PER_SM);
  int64_t grid = (int64_t)sm_count() * occ;
  if (grid > L.ntiles * L.js) grid = L.ntiles * L.js;
  kern<<<(unsigned)grid, HB, smem, L.st>>>(
      L.coords, L.values, L.n, *L.tab, *L.dp, *L.seg, (long long)L.nbuckets,
      reinterpret_cast<unsigned long long*>(L.hist), reinterpret_cast<unsigned long long*>(L.max_bits),
      L.bitmap, L.cand_key, L.cand_slot, reinterpret_cast<unsigned long long*>(L.cand_count),
      (long long)L.capacity, L.list, L.list_count, L.bbox, L.bbox32, L.wlo, L.whi, L.js, L.mv, L.bbox8,
      L.fast_seg, L.fast_lo, L.fast_hi);
  SKG_CUDA(cudaGetLastError());
  return 0;
}

// host helpers shared by the translation units (defined in pair_kernels.cu)
int check_pair_args(const double* coords, int64_t n, int dim, int64_t t0, int64_t t1, int64_t stride);
int64_t sweep_head_bytes(int64_t n);
int carve_scratch(void* scratch, int64_t scratch_bytes, int64_t n, int64_t need_tail, SweepScratch* out);
int choose_js(int64_t tiles);
int launch_cull(const double* coords, int64_t n, int dim, const CullWindows& win, int64_t t0,
                int64_t t1, int64_t stride, int js, const SweepScratch& S, cudaStream_t st,
                const BinTable* bulk_tab = nullptr, unsigned long long* bulk_count = nullptr);

// dir: 0 none, 1 mask with guard-band pairs excluded, 2 literal mask on the device
int dispatch_hist(bool priv, int emax, int dim, int dir, int stats, const HistLaunch& L);   // pair_kernels.cu
int dispatch_hist_mv(int dim, int dir, int stats, const HistLaunch& L);                      // pair_mv.cu
int dispatch_hist_win(int emax, int stats, const HistLaunch& L);                              // pair_win.cu
size_t hist_win_smem_bytes(int nb, int stats);
int dispatch_select(bool gather, int dim, int dir, bool seg, int key, const SelectLaunch& L);  // pair_select.cu
int dispatch_select_mv(bool gather, int dim, int dir, const SelectLaunch& L);                // pair_mv.cu
bool dz_gather_applicable(int dim, int use_dir, bool seg, int key, bool use_mv, const SegTable& sg, int nseg);  // pair_dzgather.cu
int launch_dz_gather(const SelectLaunch& L);

}  // namespace skg
