// Pair-sweep kernels (K1-K4 of SURVEY.md section 2.2) for sm_100a.
//
// All kernels walk the upper triangle of the N x N pair matrix in
// TILE x TILE point tiles.  A CTA of BLOCK threads keeps RI = TILE/BLOCK
// "i" points per thread in registers, stages the "j" tile in shared memory with
// coalesced loads and reads it back as warp-wide broadcasts.  The N(N-1)/2
// distance vector the reference materialises (MetricSpace.py:151-153,
// Variogram.py:1202-1208) never exists.
//
// Per pair:  s = fl(fl(dx*dx) + fl(dy*dy)) [+ fl(dw*dw)]  - scipy's euclidean
// operand order, no FMA (explicit __dmul_rn/__dadd_rn), so that every
// comparison against the squared-space thresholds reproduces np.digitize on
// d = sqrt(s) exactly (see include/skg_b200.h).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "common.cuh"

namespace skg {

// ------------------------------------------------------------------ sweep ---

template <int DIM>
struct TileSmem {
  double x[TILE];
  double y[TILE];
  double w[DIM == 3 ? TILE : 1];
  double z[TILE];
};

// Generic sweep over tiles [t0, t1) with static striding over the grid.
// Op must provide: void pair(double s, double dx, double dy, double zi, double zj, int64_t gi, int64_t gj)
template <int DIM, class Op>
__device__ __forceinline__ void sweep_tiles(const double* __restrict__ coords,
                                            const double* __restrict__ values, int64_t n,
                                            int64_t t0, int64_t t1, int64_t stride,
                                            TileSmem<DIM>& sm, Op& op) {
  const int tid = threadIdx.x;
  const int64_t nt = (n + TILE - 1) / TILE;
  const double qnan = __longlong_as_double(NAN_BITS);
  const double* __restrict__ X = coords;
  const double* __restrict__ Y = coords + n;
  const double* __restrict__ W = coords + 2 * n;

  for (int64_t t = t0 + (int64_t)blockIdx.x * stride; t < t1; t += (int64_t)gridDim.x * stride) {
    int I, J;
    tile_decode(t, nt, I, J);
    double xi[RI], yi[RI], wi[RI], zi[RI];
#pragma unroll
    for (int r = 0; r < RI; ++r) {
      int64_t i = (int64_t)I * TILE + r * BLOCK + tid;
      bool v = i < n;
      xi[r] = v ? X[i] : qnan;
      yi[r] = v ? Y[i] : qnan;
      wi[r] = (DIM == 3 && v) ? W[i] : (DIM == 3 ? qnan : 0.0);
      zi[r] = (v && values) ? values[i] : 0.0;
    }
    __syncthreads();  // all reads of the previous j tile are done
    for (int k = tid; k < TILE; k += BLOCK) {
      int64_t j = (int64_t)J * TILE + k;
      bool v = j < n;
      sm.x[k] = v ? X[j] : qnan;
      sm.y[k] = v ? Y[j] : qnan;
      if (DIM == 3) sm.w[k] = v ? W[j] : qnan;
      sm.z[k] = (v && values) ? values[j] : 0.0;
    }
    __syncthreads();
    const bool diag = (I == J);
    int64_t jn = n - (int64_t)J * TILE;
    const int jcount = jn < TILE ? (int)jn : TILE;
#pragma unroll 2
    for (int jj = 0; jj < jcount; ++jj) {
      const double xj = sm.x[jj], yj = sm.y[jj], zj = sm.z[jj];
      const double wj = (DIM == 3) ? sm.w[jj] : 0.0;
#pragma unroll
      for (int r = 0; r < RI; ++r) {
        double dx = xi[r] - xj;
        double dy = yi[r] - yj;
        double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        if (DIM == 3) {
          double dw = wi[r] - wj;
          s = __dadd_rn(s, __dmul_rn(dw, dw));
        }
        if (diag && jj <= r * BLOCK + tid) s = qnan;  // keep i < j only
        op.pair(s, dx, dy, zi[r], zj, (int64_t)I * TILE + r * BLOCK + tid, (int64_t)J * TILE + jj);
      }
    }
  }
}

__device__ __forceinline__ void load_tables(const BinTable& tab, uint64_t* sThr, uint8_t* sLut) {
  for (int k = threadIdx.x; k < THR_MAX; k += blockDim.x) sThr[k] = tab.thr[k];
  for (int k = threadIdx.x; k < tab.ncell; k += blockDim.x) sLut[k] = tab.lut[k];
}

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
};

struct SweepScratch {  // views into the caller's scratch buffer
  unsigned long long* counters;  // [0] = number of listed tiles
  double* bbox;                  // [nt][6]: lo x,y,w ; hi x,y,w
  double* bbox32;                // [nt*8][6]: boxes of the 32-point sub-blocks
  unsigned int* list;            // surviving work items, ascending
  unsigned char* flags;          // one byte per candidate work item
  unsigned char* tail;           // kernel-specific remainder (hist partials)
  int64_t tail_bytes;
};

template <int DIM>
__global__ void __launch_bounds__(256)
block_bbox_kernel(const double* __restrict__ coords, int64_t n, double* __restrict__ bbox,
                  double* __restrict__ bbox32) {
  __shared__ double sLo[DIM][8], sHi[DIM][8];
  const int64_t i = (int64_t)blockIdx.x * TILE + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double inf = __longlong_as_double((long long)INF_BITS);
#pragma unroll
  for (int k = 0; k < DIM; ++k) {
    const double v = i < n ? coords[(int64_t)k * n + i] : __longlong_as_double(NAN_BITS);
    double lo = (v == v) ? v : inf, hi = (v == v) ? v : -inf;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) {
      sLo[k][warp] = lo; sHi[k][warp] = hi;
      bbox32[((int64_t)blockIdx.x * 8 + warp) * 6 + k] = lo;      // 32-point sub-block boxes
      bbox32[((int64_t)blockIdx.x * 8 + warp) * 6 + 3 + k] = hi;
    }
  }
  __syncthreads();
  if (threadIdx.x < DIM) {
    double lo = inf, hi = -inf;
    for (int w = 0; w < 8; ++w) { lo = fmin(lo, sLo[threadIdx.x][w]); hi = fmax(hi, sHi[threadIdx.x][w]); }
    bbox[(int64_t)blockIdx.x * 6 + threadIdx.x] = lo;
    bbox[(int64_t)blockIdx.x * 6 + 3 + threadIdx.x] = hi;
  }
}

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

// Work item = tile * js + j-chunk.  Three small kernels keep, in ascending order, the items
// whose range of squared distances (i block box x j chunk box) meets a window (or can hold
// the maximum): (optional) max over tiles of dmin2, per-item flags (fully parallel),
// ordered compaction of the flags (one CTA, 4096 flags per step).
template <int DIM>
__global__ void tile_need_kernel(const double* __restrict__ bbox, int64_t nt, int64_t t_begin,
                                 int64_t ntile, int64_t t_stride,
                                 unsigned long long* __restrict__ counters) {
  double m = 0.0;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < ntile;
       k += (int64_t)gridDim.x * blockDim.x) {
    int I, J;
    tile_decode(t_begin + k * t_stride, nt, I, J);
    double a, b;
    tile_range<DIM>(bbox, I, J, a, b);
    if (a == a && b == b && a <= b) m = fmax(m, a);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(&counters[1], (unsigned long long)__double_as_longlong(m));
}

template <int DIM>
__global__ void item_flags_kernel(const double* __restrict__ bbox, const double* __restrict__ bbox32,
                                  int64_t n, int64_t nt, int64_t t_begin, int64_t ntile,
                                  int64_t t_stride, int js, const __grid_constant__ CullWindows win,
                                  const unsigned long long* __restrict__ counters,
                                  unsigned char* __restrict__ flags) {
  const int64_t nitem = ntile * js;
  const int jchunk = TILE / js;
  const double need = win.want_max
      ? __longlong_as_double((long long)counters[1]) * (1.0 - 1e-12) : 0.0;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nitem;
       k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = t_begin + (k / js) * t_stride;
    const int h = (int)(k % js);
    int I, J;
    tile_decode(t, nt, I, J);
    int keep = 0;
    if ((int64_t)J * TILE + (int64_t)h * jchunk < n) {  // the chunk holds points
      double jb[6], a, b;
      chunk_box<DIM>(bbox32, J, h, js, jb);
      subblock_range<DIM>(bbox, I, jb, a, b);
      if (a == a && b == b && a <= b) {
        const double lo = a * (1.0 - 1e-12), hi = b * (1.0 + 1e-12);
        for (int w = 0; w < win.count; ++w) keep |= (lo < win.hi[w] && hi >= win.lo[w]) ? 1 : 0;
        if (win.want_max && hi >= need) keep = 1;
      }
    }
    flags[k] = (unsigned char)keep;
  }
}

__global__ void __launch_bounds__(1024)
item_compact_kernel(const unsigned char* __restrict__ flags, int64_t t_begin, int64_t ntile,
                    int64_t t_stride, int js, unsigned int* __restrict__ list,
                    unsigned long long* __restrict__ counters) {
  __shared__ int warp_tot[32];
  __shared__ long long carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t nitem = ntile * js;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < nitem; base += 4096) {
    const int64_t k0 = base + (int64_t)tid * 4;
    int f[4], cnt = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      f[q] = (k0 + q < nitem) ? flags[k0 + q] : 0;
      cnt += f[q];
    }
    int x = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int v = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += y;
      }
      warp_tot[lane] = v;
    }
    __syncthreads();
    long long pos = carry + (warp ? warp_tot[warp - 1] : 0) + (x - cnt);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (f[q]) {
        const int64_t k = k0 + q;
        const int64_t t = t_begin + (k / js) * t_stride;
        list[pos++] = (unsigned int)(t * js + (k % js));
      }
    }
    __syncthreads();
    if (tid == blockDim.x - 1) carry = pos;
    __syncthreads();
  }
  if (tid == 0) {
    counters[0] = (unsigned long long)carry;  // surviving work items
    counters[2] = (unsigned long long)js;     // j-chunks per tile (for reporting)
    counters[3] = (unsigned long long)nitem;  // candidate work items before culling
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

struct DzTable {  // per-lag |dz| thresholds of the below-count mode (STATS == 4)
  double lo[THR_MAX];
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

template <int STATS>
__device__ __forceinline__ void rmw_record(unsigned addr, double dz, int pred) {
  if (STATS == 3) {
    const double rt = sqrt(fabs(dz));
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 tb, cb, rb, pb;\n\t.reg .f64 t, u;\n\t"
        "setp.ne.s32 p, %3, 0;\n\t"
        "@p ld.shared.v2.b64 {tb, cb}, [%0];\n\t"
        "@p ld.shared.v2.b64 {rb, pb}, [%0+16];\n\t"
        "mov.b64 t, tb;\n\tmov.b64 u, rb;\n\t"
        "fma.rn.f64 t, %1, %1, t;\n\t"
        "add.rn.f64 u, u, %2;\n\t"
        "{ .reg .b32 lo, hi; mov.b64 {lo, hi}, cb; add.u32 lo, lo, 1; mov.b64 cb, {lo, hi}; }\n\t"
        "mov.b64 tb, t;\n\tmov.b64 rb, u;\n\t"
        "@p st.shared.v2.b64 [%0], {tb, cb};\n\t"
        "@p st.shared.b64 [%0+16], rb;\n\t}"
        :: "r"(addr), "d"(dz), "d"(rt), "r"(pred));
  } else if (STATS == 2) {
    const double rt = sqrt(fabs(dz));
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 tb, cb;\n\t.reg .f64 t;\n\t"
        "setp.ne.s32 p, %2, 0;\n\t"
        "@p ld.shared.v2.b64 {tb, cb}, [%0];\n\t"
        "mov.b64 t, tb;\n\t"
        "add.rn.f64 t, t, %1;\n\t"
        "{ .reg .b32 lo, hi; mov.b64 {lo, hi}, cb; add.u32 lo, lo, 1; mov.b64 cb, {lo, hi}; }\n\t"
        "mov.b64 tb, t;\n\t"
        "@p st.shared.v2.b64 [%0], {tb, cb};\n\t}"
        :: "r"(addr), "d"(rt), "r"(pred));
  } else if (STATS == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 tb, cb;\n\t.reg .f64 t;\n\t"
        "setp.ne.s32 p, %2, 0;\n\t"
        "@p ld.shared.v2.b64 {tb, cb}, [%0];\n\t"
        "mov.b64 t, tb;\n\t"
        "fma.rn.f64 t, %1, %1, t;\n\t"
        "{ .reg .b32 lo, hi; mov.b64 {lo, hi}, cb; add.u32 lo, lo, 1; mov.b64 cb, {lo, hi}; }\n\t"
        "mov.b64 tb, t;\n\t"
        "@p st.shared.v2.b64 [%0], {tb, cb};\n\t}"
        :: "r"(addr), "d"(dz), "r"(pred));
  } else if (STATS == 4) {
    // dz carries the 0/1 "below the per-lag |dz| threshold" flag: low word of the first
    // 8 bytes counts those pairs, low word of the second 8 bytes counts all pairs
    const int below = dz != 0.0 ? 1 : 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 bb, cb;\n\t.reg .b32 lo, hi;\n\t"
        "setp.ne.s32 p, %2, 0;\n\t"
        "@p ld.shared.v2.b64 {bb, cb}, [%0];\n\t"
        "mov.b64 {lo, hi}, bb;\n\tadd.u32 lo, lo, %1;\n\tmov.b64 bb, {lo, hi};\n\t"
        "mov.b64 {lo, hi}, cb;\n\tadd.u32 lo, lo, 1;\n\tmov.b64 cb, {lo, hi};\n\t"
        "@p st.shared.v2.b64 [%0], {bb, cb};\n\t}"
        :: "r"(addr), "r"(below), "r"(pred));
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 cb;\n\t"
        "setp.ne.s32 p, %1, 0;\n\t"
        "@p ld.shared.b64 cb, [%0+8];\n\t"
        "{ .reg .b32 lo, hi; mov.b64 {lo, hi}, cb; add.u32 lo, lo, 1; mov.b64 cb, {lo, hi}; }\n\t"
        "@p st.shared.b64 [%0+8], cb;\n\t}"
        :: "r"(addr), "r"(pred));
  }
}

constexpr int HB = 64;          // threads per CTA
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
    for (int e = 0; e < emax; ++e) g += (s >= dThr[g]) ? 1 : 0;
  }
  return g;
}

// PRIV: private per-thread records + software-pipelined loop (front end of pair set
// jj+1 is issued before the read-modify-writes of pair set jj).
// !PRIV (n_lags too large for private records): one shared record per lag + atomics.
template <int DIM, int DIR, int STATS, int EMAX, bool PRIV>
__global__ void __launch_bounds__(HB)
pair_hist_kernel(const double* __restrict__ coords, const double* __restrict__ values, int64_t n,
                 const __grid_constant__ BinTable tab, const __grid_constant__ DirParams dp,
                 HistPartial* __restrict__ partial, const unsigned int* __restrict__ tile_list,
                 const unsigned long long* __restrict__ tile_count,
                 const double* __restrict__ bbox, const double* __restrict__ bbox32, double wlo,
                 double whi, const __grid_constant__ DzTable dzt, int js) {
  __shared__ double2 sXY[TILE + 2];
  __shared__ double sW[DIM == 3 ? TILE + 2 : 1];
  __shared__ double sZ[TILE + 2];
  __shared__ uint64_t sThr[THR_MAX];
  __shared__ uint8_t sLut[LUT_MAX];
  __shared__ double sDzLo[STATS == 4 ? THR_MAX : 1];  // per-lag |dz| thresholds (below-count mode)
  extern __shared__ __align__(16) unsigned char acc_raw[];

  constexpr int REC = (STATS == 3) ? 32 : 16;
  const int nb = tab.nb;
  const int S = PRIV ? HB : 1;
  const int tid = threadIdx.x;
  const double qnan = __longlong_as_double(NAN_BITS);

  for (int k = tid; k < THR_MAX; k += HB) sThr[k] = tab.thr[k];
  for (int k = tid; k < tab.ncell; k += HB) sLut[k] = tab.lut[k];
  if (STATS == 4) {
    for (int k = tid; k < THR_MAX; k += HB) sDzLo[k] = k < nb ? dzt.lo[k] : 0.0;
  }
  {
    unsigned long long* z = reinterpret_cast<unsigned long long*>(acc_raw);
    for (int k = tid; k < nb * S * (REC / 8); k += HB) z[k] = 0ull;
  }
  __syncthreads();

  const int shift = tab.shift, klo = tab.key0, khi = tab.key0 + tab.ncell - 1, emax = tab.emax;
  const uint8_t* lutb = sLut - klo;
  const unsigned aRec = (unsigned)__cvta_generic_to_shared(acc_raw) + (PRIV ? tid * REC : 0);
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
        subblock_range<DIM>(bbox32, (int64_t)I * 8 + r * (HB / 32) + (tid >> 5), jb, a, b);
        wact |= (a * (1.0 - 1e-12) < whi && b * (1.0 + 1e-12) >= wlo);
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
      zi[r] = v ? values[i] : 0.0;
    }
    __syncthreads();  // every read of the previous j chunk is done
    for (int k = tid; k < jcount + 2; k += HB) {  // two NaN entries pad the pipeline tail
      const int64_t j = (int64_t)J * TILE + jbeg + k;
      const bool v = k < jcount;
      sXY[k] = make_double2(v ? X[j] : qnan, v ? Y[j] : qnan);
      if (DIM == 3) sW[k] = v ? W[j] : qnan;
      sZ[k] = v ? values[j] : 0.0;
    }
    __syncthreads();
    if (jcount == 0 || !wact) continue;  // warp-uniform; both barriers are already passed

    auto body = [&](auto diag_tag) {
      constexpr bool DIAG = decltype(diag_tag)::value;
      // front end of pair set jj: lag classes g[] (nb = not accumulated) and dz[]
      auto front = [&](int jj, int (&g)[HR], double (&dz)[HR]) {
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
          int gg = lag_of<EMAX>(sb, sThr, lutb, shift, klo, khi, emax);
          if (DIAG) gg = (jbeg + jj <= r * HB + tid) ? nb : gg;  // keep i < j only
          if (DIR) {
            if (gg < nb && !dir_mask_t<DIR == 2>(dx, dy, s, dp)) gg = nb;
          }
          g[r] = gg;
          dz[r] = zi[r] - zj;
          if (STATS == 4) dz[r] = (fabs(dz[r]) < sDzLo[min(gg, nb - 1)]) ? 1.0 : 0.0;
        }
      };
      auto accumulate = [&](const int (&g)[HR], const double (&dz)[HR]) {
#pragma unroll
        for (int r = 0; r < HR; ++r) {
          const int in = g[r] < nb ? 1 : 0;
          if (PRIV) {
            rmw_record<STATS>(aRec + (unsigned)g[r] * (unsigned)(HB * REC), dz[r], in);
          } else if (in) {
            unsigned char* rec = acc_raw + g[r] * REC;
            atomicAdd(reinterpret_cast<unsigned long long*>(rec + 8), 1ull);
            if (STATS & SKG_STAT_SUM_SQ) atomicAdd(reinterpret_cast<double*>(rec), dz[r] * dz[r]);
            if (STATS & SKG_STAT_SUM_SQRT)
              atomicAdd(reinterpret_cast<double*>(rec + (STATS == 3 ? 16 : 0)), sqrt(fabs(dz[r])));
          }
        }
      };
      int gA[HR], gB[HR];
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
        const unsigned char* rec = acc_raw + (size_t)(gI * HB + lane + 32 * k) * REC;
        c += *reinterpret_cast<const unsigned int*>(rec + 8);  // per-thread counts stay < 2^32
        if (STATS == 4) a += (double)*reinterpret_cast<const unsigned int*>(rec);
        else if (STATS & SKG_STAT_SUM_SQ) a += *reinterpret_cast<const double*>(rec);
        if (STATS & SKG_STAT_SUM_SQRT) b += *reinterpret_cast<const double*>(rec + (STATS == 3 ? 16 : 0));
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

// Folds the per-CTA partials in a FIXED order (one warp per lag class; lane l takes CTAs
// l, l+32, ... in sequence, then a shuffle tree): bitwise reproducible results.
__global__ void __launch_bounds__(32)
pair_hist_reduce_kernel(const HistPartial* __restrict__ partial, int nblocks, int nb,
                        unsigned long long* __restrict__ count, double* __restrict__ sum_sq,
                        double* __restrict__ sum_sqrt) {
  const int g = blockIdx.x, lane = threadIdx.x;
  unsigned long long c = 0;
  double a = 0.0, b = 0.0;
#pragma unroll 4
  for (int k = lane; k < nblocks; k += 32) {
    const HistPartial p = partial[(int64_t)k * nb + g];
    c += p.count;
    a += p.sum_sq;
    b += p.sum_sqrt;
  }
  c = warp_sum(c);
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) {
    count[g] += c;
    if (sum_sq) sum_sq[g] += a;
    if (sum_sqrt) sum_sqrt[g] += b;
  }
}

// ------------------------------------------------- K1/K3: select passes ---

// Same sweep structure as the fused kernel (64 threads x 4 "i" points, static tables,
// j-chunk work items, software-pipelined front end).  Histogram pass: one u64 atomicAdd
// (L2) per key INSIDE the segment's window - the host keeps the windows narrow (the class
// that holds a wanted rank), so most pairs only pay the front end.  Gather pass: bitmap
// test and append.  The per-thread running maximum of the keys feeds atomicMax once.
template <int DIM, int DIR, bool SEG, int KEY, bool GATHER, int EMAX>
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
                   double whi, int js) {
  __shared__ double2 sXY[TILE + 2];
  __shared__ double sW[DIM == 3 ? TILE + 2 : 1];
  __shared__ double sZ[TILE + 2];
  __shared__ uint64_t sThr[THR_MAX];
  __shared__ uint8_t sLut[LUT_MAX];
  __shared__ SegTable sSeg;
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
  __syncthreads();
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
        subblock_range<DIM>(bbox32, (int64_t)I * 8 + r * (HB / 32) + (tid >> 5), jb, a, b);
        wact |= (a * (1.0 - 1e-12) < whi && b * (1.0 + 1e-12) >= wlo);
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
      zi[r] = (v && values) ? values[i] : 0.0;
    }
    __syncthreads();
    for (int k = tid; k < jcount + 2; k += HB) {
      const int64_t j = (int64_t)J * TILE + jbeg + k;
      const bool v = k < jcount;
      sXY[k] = make_double2(v ? X[j] : qnan, v ? Y[j] : qnan);
      if (DIM == 3) sW[k] = v ? W[j] : qnan;
      sZ[k] = (v && values) ? values[j] : 0.0;
    }
    __syncthreads();
    if (jcount == 0 || !wact) continue;

    auto body = [&](auto diag_tag) {
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
          if (SEG) gg = lag_of<EMAX>(sb, sThr, lutb, shift, klo, khi, emax);
          else gg = sb <= INF_BITS ? 0 : nb;  // NaN: padded point or NaN coordinate
          if (DIAG) gg = (jbeg + jj <= r * HB + tid) ? nb : gg;  // keep i < j only
          if (DIR) {
            if (gg < nb && !dir_mask_t<DIR == 2>(dx, dy, s, dp)) gg = nb;
          }
          g[r] = gg;
          key[r] = (KEY == SKG_KEY_SQDIST) ? s : fabs(zi[r] - zj);
        }
      };
      auto accumulate = [&](const int (&g)[HR], const double (&key)[HR]) {
#pragma unroll
        for (int r = 0; r < HR; ++r) {
          if (g[r] >= nb) continue;
          const uint64_t kb = (uint64_t)__double_as_longlong(key[r]);
          if (kb > INF_BITS) continue;  // NaN value
          if (!SEG && !GATHER) local_max = kb > local_max ? kb : local_max;
          if (kb < sSeg.lo_bits[g[r]] || kb >= sSeg.hi_bits[g[r]]) continue;
          const double lo = __longlong_as_double((long long)sSeg.lo_bits[g[r]]);
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
      front(0, gA, kA);
      int jj = 1;
      for (; jj + 1 <= jcount; jj += 2) {
        front(jj, gB, kB);
        accumulate(gA, kA);
        front(jj + 1, gA, kA);
        accumulate(gB, kB);
      }
      if (jj <= jcount) {
        front(jj, gB, kB);
        accumulate(gA, kA);
      }
    };
    if (I == J) body(std::true_type{}); else body(std::false_type{});
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

// ------------------------------------------- guard-band pair collection ---

struct AmbOp {
  const DirParams* dp;
  int* out_i;
  int* out_j;
  unsigned long long* count;
  long long capacity;
  __device__ __forceinline__ void pair(double s, double dx, double dy, double, double, int64_t gi, int64_t gj) {
    if (!(s > 0.0)) return;
    bool amb;
    dir_fast(dx, dy, *dp, &amb);
    if (amb) {
      const unsigned long long pos = atomicAdd(count, 1ull);
      if ((long long)pos < capacity) {
        out_i[pos] = (int)gi;
        out_j[pos] = (int)gj;
      }
    }
  }
};

__global__ void __launch_bounds__(BLOCK)
pair_dir_ambiguous_kernel(const double* __restrict__ coords, int64_t n,
                          const __grid_constant__ DirParams dp, int* __restrict__ out_i,
                          int* __restrict__ out_j, unsigned long long* __restrict__ count,
                          long long capacity, int64_t t0, int64_t t1, int64_t stride) {
  __shared__ TileSmem<2> sm;
  AmbOp op;
  op.dp = &dp; op.out_i = out_i; op.out_j = out_j; op.count = count; op.capacity = capacity;
  sweep_tiles<2>(coords, nullptr, n, t0, t1, stride, sm, op);
}

// ------------------------------------------------------- materialisation ---

template <int DIM>
__global__ void pair_materialize_kernel(const double* __restrict__ coords,
                                        const double* __restrict__ values, int64_t n,
                                        const __grid_constant__ BinTable tab,
                                        const __grid_constant__ DirParams dp, int use_dir,
                                        double* __restrict__ dist, double* __restrict__ diffs,
                                        long long* __restrict__ groups, uint8_t* __restrict__ mask,
                                        int64_t k0, int64_t k1) {
  const double* X = coords;
  const double* Y = coords + n;
  const double* W = coords + 2 * n;
  for (int64_t k = k0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < k1;
       k += (int64_t)gridDim.x * blockDim.x) {
    // condensed index -> (i, j): rowstart(i) = i*n - i(i+1)/2
    double nn = (double)n;
    int64_t i = (int64_t)(nn - 0.5 - sqrt((nn - 0.5) * (nn - 0.5) - 2.0 * (double)k));
    if (i < 0) i = 0;
    if (i > n - 2) i = n - 2;
    while (i > 0 && i * n - i * (i + 1) / 2 > k) --i;
    while ((i + 1) * n - (i + 1) * (i + 2) / 2 <= k) ++i;
    const int64_t j = k - (i * n - i * (i + 1) / 2) + i + 1;
    const double dx = X[i] - X[j];
    const double dy = Y[i] - Y[j];
    double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    if (DIM == 3) {
      const double dw = W[i] - W[j];
      s = __dadd_rn(s, __dmul_rn(dw, dw));
    }
    const int64_t o = k - k0;
    bool m = true;
    if (use_dir) m = dir_mask(dx, dy, s, dp);
    if (dist) dist[o] = sqrt(s);
    if (diffs) diffs[o] = fabs(values[i] - values[j]);
    if (mask) mask[o] = m ? 1 : 0;
    if (groups) {
      const uint64_t sb = (uint64_t)__double_as_longlong(s);
      int g = 0;
      for (int e = 0; e < tab.nb; ++e) g += (sb >= tab.thr[e]) ? 1 : 0;
      groups[o] = (g >= tab.nb || !m) ? -1 : g;
    }
  }
}

// ------------------------------------------------------------ host side ---

// T(e) thresholds arrive as bit patterns.  cell(x) = (hi32(x) >> shift) - key0,
// clamped to [0, ncell-1]; lut[c] = number of thresholds in strictly lower
// cells; emax = max number of thresholds sharing one cell.
int build_bin_table(const uint64_t* thr_bits, int n_lags, BinTable* out) {
  std::memset(out, 0, sizeof(BinTable));
  for (int k = 0; k < THR_MAX; ++k) out->thr[k] = ~0ull;
  out->nb = n_lags;
  out->shift = 0; out->key0 = 0; out->ncell = 1; out->emax = 0;
  out->lut[0] = 0;
  if (n_lags == 0) return 0;
  if (n_lags < 0 || n_lags > SKG_MAX_LAGS) {
    set_error("n_lags=%d outside [0, %d]", n_lags, SKG_MAX_LAGS);
    return SKG_E_BADARG;
  }
  for (int k = 0; k < n_lags; ++k) {
    if (k && thr_bits[k] < thr_bits[k - 1]) {
      set_error("lag thresholds must be non-decreasing (k=%d)", k);
      return SKG_E_BADARG;
    }
    if (thr_bits[k] > INF_BITS) {
      set_error("lag threshold %d is NaN", k);
      return SKG_E_BADARG;
    }
    out->thr[k] = thr_bits[k];
  }
  const uint32_t hi_last = (uint32_t)(thr_bits[n_lags - 1] >> 32);
  // lowest key worth resolving: at most 48 octaves below the last threshold
  uint32_t hi_first = (uint32_t)(thr_bits[0] >> 32);
  const uint32_t floor_hi = hi_last > (48u << 20) ? hi_last - (48u << 20) : 0u;
  if (hi_first < floor_hi) hi_first = floor_hi;
  int shift = 0;
  for (; shift < 31; ++shift) {
    int64_t cells = (int64_t)(hi_last >> shift) - (int64_t)(hi_first >> shift) + 3;
    if (cells <= LUT_MAX) break;
  }
  const int key0 = (int)(hi_first >> shift) - 1;  // cell 0: everything below the first threshold's cell
  const int ncell = (int)(hi_last >> shift) - key0 + 2;  // top cell: above the last threshold's cell
  out->shift = shift; out->key0 = key0; out->ncell = ncell;
  int per_cell[LUT_MAX];
  std::memset(per_cell, 0, sizeof(per_cell));
  for (int k = 0; k < n_lags; ++k) {
    int c = (int)((uint32_t)(thr_bits[k] >> 32) >> shift) - key0;
    c = std::min(std::max(c, 0), ncell - 1);
    per_cell[c]++;
  }
  int run = 0, emax = 0;
  for (int c = 0; c < ncell; ++c) {
    out->lut[c] = (uint8_t)run;
    run += per_cell[c];
    emax = std::max(emax, per_cell[c]);
  }
  out->emax = emax;
  return 0;
}

int build_dir_params(const double* dir, int dim, DirParams* out) {
  std::memset(out, 0, sizeof(DirParams));
  if (!dir) return 0;
  const int model = (int)dir[0];
  if (model == SKG_DIR_NONE) return 0;
  if (model != SKG_DIR_TRIANGLE && model != SKG_DIR_COMPASS) {
    set_error("unknown directional model id %d", model);
    return SKG_E_BADARG;
  }
  if (dim != 2) {
    set_error("directional mask needs 2-D coordinates (DirectionalVariogram.py:359-364)");
    return SKG_E_UNSUPPORTED;
  }
  out->model = model;
  out->az_rad = dir[1];
  out->half_tol_rad = dir[2];
  out->half_bw = dir[3];
  out->cos_az = std::cos(dir[1]);
  out->sin_az = std::sin(dir[1]);
  const double half_pi = 3.141592653589793 / 2;
  out->tol_all = dir[2] >= half_pi ? 1 : 0;
  out->tan_ht = out->tol_all ? 0.0 : std::tan(dir[2]);
  out->amb_exclude = dir[4] != 0.0 ? 1 : 0;
  return 0;
}

static int check_pair_args(const double* coords, int64_t n, int dim, int64_t t0, int64_t t1,
                           int64_t stride = 1) {
  if (!coords || n < 0) { set_error("coords NULL or n < 0"); return SKG_E_BADARG; }
  if (dim != 2 && dim != 3) {
    set_error("dim=%d unsupported (2 or 3; pad 1-D input with a zero column)", dim);
    return SKG_E_UNSUPPORTED;
  }
  const int64_t ntiles = skg_pair_num_tiles(n);
  if (t0 < 0 || t1 < t0 || t1 > ntiles || stride < 1) {
    set_error("tile range [%lld, %lld) step %lld outside [0, %lld]", (long long)t0, (long long)t1,
              (long long)stride, (long long)ntiles);
    return SKG_E_BADARG;
  }
  if (n > 23000000ll) { set_error("n too large for 32-bit tile ids"); return SKG_E_UNSUPPORTED; }
  return 0;
}

constexpr int64_t LIST_SPLIT_CAP = 1 << 22;  // j-splitting only while tiles * js fits this

static int64_t list_entries(int64_t ntiles) { return std::max<int64_t>(ntiles, std::min<int64_t>(ntiles * 16, LIST_SPLIT_CAP)); }

static int64_t sweep_head_bytes(int64_t n) {
  const int64_t nt = (n + TILE - 1) / TILE;
  const int64_t ntiles = nt * (nt + 1) / 2;
  int64_t b = 256 + nt * 6 * 8 + nt * 8 * 6 * 8 + list_entries(ntiles) * 5;
  return (b + 255) & ~255ll;
}

static int carve_scratch(void* scratch, int64_t scratch_bytes, int64_t n, int64_t need_tail,
                         SweepScratch* out) {
  const int64_t head = sweep_head_bytes(n);
  if (!scratch || scratch_bytes < head + need_tail) {
    set_error("scratch too small: need %lld bytes", (long long)(head + need_tail));
    return SKG_E_SCRATCH;
  }
  const int64_t nt = (n + TILE - 1) / TILE;
  unsigned char* base = reinterpret_cast<unsigned char*>(scratch);
  out->counters = reinterpret_cast<unsigned long long*>(base);
  out->bbox = reinterpret_cast<double*>(base + 256);
  out->bbox32 = reinterpret_cast<double*>(base + 256 + nt * 6 * 8);
  out->list = reinterpret_cast<unsigned int*>(base + 256 + nt * 6 * 8 + nt * 8 * 6 * 8);
  out->flags = reinterpret_cast<unsigned char*>(out->list) + list_entries(nt * (nt + 1) / 2) * 4;
  out->tail = base + head;
  out->tail_bytes = scratch_bytes - head;
  return 0;
}

// bounding boxes of the 256-point blocks + ordered list of the tiles worth sweeping
// j-split: enough (tile, chunk) work items per CTA that culling and the static round-robin
// stay balanced; bounded by the list capacity carved from the scratch
static int choose_js(int64_t tiles) {
  const int64_t want = (int64_t)sm_count() * 6 * 12;
  int js = 1;
  while (js < 16 && tiles * js < want && tiles * (js * 2) <= LIST_SPLIT_CAP) js *= 2;
  if (const char* f = getenv("SKG_FORCE_JS")) {  // developer override
    const int fj = atoi(f);
    if (fj >= 1 && fj <= 16 && (fj & (fj - 1)) == 0 && tiles * fj <= LIST_SPLIT_CAP) js = fj;
  }
  return js;
}

static int launch_cull(const double* coords, int64_t n, int dim, const CullWindows& win,
                       int64_t t0, int64_t t1, int64_t stride, int js, const SweepScratch& S,
                       cudaStream_t st) {
  const int64_t nt = (n + TILE - 1) / TILE;
  const int64_t ntile = t1 > t0 ? (t1 - t0 + stride - 1) / stride : 0;
  const int64_t nitem = ntile * js;
  const unsigned fb = (unsigned)std::min<int64_t>((nitem + 255) / 256, (int64_t)sm_count() * 8);
  const unsigned tb = (unsigned)std::min<int64_t>((ntile + 255) / 256, (int64_t)sm_count() * 8);
  if (win.want_max) SKG_CUDA(cudaMemsetAsync(S.counters + 1, 0, 8, st));
  if (dim == 3) {
    block_bbox_kernel<3><<<(unsigned)nt, 256, 0, st>>>(coords, n, S.bbox, S.bbox32);
    if (win.want_max) tile_need_kernel<3><<<tb, 256, 0, st>>>(S.bbox, nt, t0, ntile, stride, S.counters);
    item_flags_kernel<3><<<fb, 256, 0, st>>>(S.bbox, S.bbox32, n, nt, t0, ntile, stride, js, win,
                                             S.counters, S.flags);
  } else {
    block_bbox_kernel<2><<<(unsigned)nt, 256, 0, st>>>(coords, n, S.bbox, S.bbox32);
    if (win.want_max) tile_need_kernel<2><<<tb, 256, 0, st>>>(S.bbox, nt, t0, ntile, stride, S.counters);
    item_flags_kernel<2><<<fb, 256, 0, st>>>(S.bbox, S.bbox32, n, nt, t0, ntile, stride, js, win,
                                             S.counters, S.flags);
  }
  item_compact_kernel<<<1, 1024, 0, st>>>(S.flags, t0, ntile, stride, js, S.list, S.counters);
  SKG_CUDA(cudaGetLastError());
  return 0;
}

static double bits_to_double(uint64_t b) {
  double d;
  std::memcpy(&d, &b, 8);
  return d;
}

constexpr int MAX_GRID_PER_SM = 16;

// (CTAs per SM, j-split) keeping the static round-robin balanced; shared by all sweeps
static void choose_grid(int64_t tiles, int occ, int64_t* grid_out, int* js_out) {
  const int sms = sm_count();
  double best = -1.0;
  int64_t grid = 1;
  int js = 1;
  for (int k = occ; k >= std::max(1, occ - 1); --k) {
    for (int cand = 1; cand <= 16; cand *= 2) {
      const int64_t items = tiles * cand;
      const int64_t g = std::min<int64_t>((int64_t)sms * k, items);
      const int64_t rounds = (items + g - 1) / g;
      double eff = (double)items / ((double)g * (double)rounds);
      eff *= (g >= (int64_t)sms * k) ? (0.92 + 0.08 * k / occ) : (double)g / ((double)sms * occ);
      eff *= 1.0 - 0.03 * std::log2((double)cand);  // per-item staging overhead (measured)
      if (eff > best + 0.004) { best = eff; grid = g; js = cand; }
    }
  }
  if (const char* f = getenv("SKG_FORCE_JS")) {  // developer override: "<js>,<ctas_per_sm>"
    int fj = 0, fk = 0;
    if (sscanf(f, "%d,%d", &fj, &fk) == 2 && fj >= 1 && fj <= 16 && fk >= 1) {
      js = fj;
      grid = std::min<int64_t>((int64_t)sms * std::min(fk, occ), tiles * js);
    }
  }
  *grid_out = grid;
  *js_out = js;
}

static size_t hist_smem_bytes(int nb, int stats, bool priv) {
  const size_t S = priv ? HB : 1;
  return (size_t)nb * S * (stats == 3 ? 32 : 16);
}

struct HistLaunch {
  const double* coords; const double* values; int64_t n;
  const BinTable* tab; const DirParams* dp; HistPartial* partial;
  const unsigned int* list; const unsigned long long* list_count;
  const double* bbox; const double* bbox32; double wlo, whi; const DzTable* dzt;
  int max_blocks; int64_t ntiles; int js; size_t smem; cudaStream_t st; int* nblocks_out;
};

template <int DIM, int DIR, int STATS, int EMAX, bool PRIV>
static int launch_hist(const HistLaunch& L) {
  auto kern = pair_hist_kernel<DIM, DIR, STATS, EMAX, PRIV>;
  static thread_local size_t cached_smem = (size_t)-1;
  static thread_local int cached_occ = 0;
  if (cached_smem != L.smem) {
    SKG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
    SKG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cached_occ, kern, HB, L.smem));
    cached_smem = L.smem;
  }
  int occ = cached_occ;
  if (occ < 1) { set_error("pair_hist kernel does not fit (smem=%zu)", L.smem); return SKG_E_UNSUPPORTED; }
  occ = std::min(occ, MAX_GRID_PER_SM);
  const int js = L.js;
  int64_t grid = std::min<int64_t>((int64_t)sm_count() * occ, L.ntiles * js);
  grid = std::min<int64_t>(grid, L.max_blocks);
  *L.nblocks_out = (int)grid;
  kern<<<(unsigned)grid, HB, L.smem, L.st>>>(L.coords, L.values, L.n, *L.tab, *L.dp, L.partial,
                                             L.list, L.list_count, L.bbox, L.bbox32, L.wlo, L.whi,
                                             *L.dzt, js);
  SKG_CUDA(cudaGetLastError());
  return 0;
}

template <int DIM, int DIR, int EMAX, bool PRIV>
static int dispatch_hist_stats(int stats, const HistLaunch& L) {
  switch (stats) {
    case 0: return launch_hist<DIM, DIR, 0, EMAX, PRIV>(L);
    case 1: return launch_hist<DIM, DIR, 1, EMAX, PRIV>(L);
    case 2: return launch_hist<DIM, DIR, 2, EMAX, PRIV>(L);
    case 4: return launch_hist<DIM, DIR, 4, EMAX, PRIV>(L);
    default: return launch_hist<DIM, DIR, 3, EMAX, PRIV>(L);
  }
}

template <int EMAX, bool PRIV>
static int dispatch_hist_dim(int dim, int dir, int stats, const HistLaunch& L) {
  if (dim == 3) return dispatch_hist_stats<3, 0, EMAX, PRIV>(stats, L);
  if (dir == 1) return dispatch_hist_stats<2, 1, EMAX, PRIV>(stats, L);
  if (dir == 2) return dispatch_hist_stats<2, 2, EMAX, PRIV>(stats, L);
  return dispatch_hist_stats<2, 0, EMAX, PRIV>(stats, L);
}

static int dispatch_hist(bool priv, int emax, int dim, int dir, int stats, const HistLaunch& L) {
  if (!priv) return dispatch_hist_dim<0, false>(dim, dir, stats, L);
  if (emax <= 1) return dispatch_hist_dim<1, true>(dim, dir, stats, L);
  return dispatch_hist_dim<0, true>(dim, dir, stats, L);
}

template <int DIM, int DIR, bool SEG, int KEY, bool GATHER>
static int launch_select(const double* coords, const double* values, int64_t n, const BinTable& tab,
                         const DirParams& dp, const SegTable& seg_tab, int64_t nbuckets,
                         uint64_t* hist, uint64_t* max_bits, const uint32_t* bitmap,
                         double* cand_key, uint32_t* cand_slot, uint64_t* cand_count,
                         int64_t capacity, int64_t ntiles, const unsigned int* list,
                         const unsigned long long* list_count, const double* bbox,
                         const double* bbox32, double wlo, double whi, int js, cudaStream_t st) {
  const bool e1 = tab.emax <= 1;
  auto kern = e1 ? pair_select_kernel<DIM, DIR, SEG, KEY, GATHER, 1>
                 : pair_select_kernel<DIM, DIR, SEG, KEY, GATHER, 0>;
  int occ = 0;
  SKG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, HB, 0));
  occ = std::max(1, std::min(occ, MAX_GRID_PER_SM));
  const int64_t grid = std::min<int64_t>((int64_t)sm_count() * occ, ntiles * js);
  kern<<<(unsigned)grid, HB, 0, st>>>(
      coords, values, n, tab, dp, seg_tab, (long long)nbuckets,
      reinterpret_cast<unsigned long long*>(hist), reinterpret_cast<unsigned long long*>(max_bits),
      bitmap, cand_key, cand_slot, reinterpret_cast<unsigned long long*>(cand_count),
      (long long)capacity, list, list_count, bbox, bbox32, wlo, whi, js);
  SKG_CUDA(cudaGetLastError());
  return 0;
}

template <bool GATHER, typename... A>
static int dispatch_select(int dim, int dir, bool seg, int key, const A&... a) {
#define SKG_SEL(D, DR, SG, K) return launch_select<D, DR, SG, K, GATHER>(a...)
#define SKG_SEL_KEY(D, DR, SG) do { if (key == SKG_KEY_SQDIST) SKG_SEL(D, DR, SG, 0); else SKG_SEL(D, DR, SG, 1); } while (0)
  if (dim == 3) { if (seg) SKG_SEL_KEY(3, 0, true); else SKG_SEL_KEY(3, 0, false); }
  if (dir == 1) { if (seg) SKG_SEL_KEY(2, 1, true); else SKG_SEL_KEY(2, 1, false); }
  if (dir == 2) { if (seg) SKG_SEL_KEY(2, 2, true); else SKG_SEL_KEY(2, 2, false); }
  if (seg) SKG_SEL_KEY(2, 0, true); else SKG_SEL_KEY(2, 0, false);
#undef SKG_SEL_KEY
#undef SKG_SEL
}

// windows of squared distance in which a select pass can find anything
static void select_windows(int n_lags, int key_kind, const uint64_t* thr_bits,
                           const uint64_t* lo_bits, const uint64_t* hi_bits, bool want_max,
                           CullWindows* w) {
  std::memset(w, 0, sizeof(*w));
  w->want_max = want_max ? 1 : 0;
  const double inf = bits_to_double(INF_BITS);
  auto add = [&](double lo, double hi) {
    if (!(hi > lo)) return;
    if (w->count && lo <= w->hi[w->count - 1]) {  // merge touching / overlapping windows
      w->hi[w->count - 1] = std::max(w->hi[w->count - 1], hi);
      return;
    }
    if (w->count == 64) { w->hi[63] = std::max(w->hi[63], hi); return; }
    w->lo[w->count] = lo;
    w->hi[w->count] = hi;
    w->count++;
  };
  if (n_lags <= 0) {
    if (hi_bits[0] > lo_bits[0]) {
      if (key_kind == SKG_KEY_SQDIST)
        add(bits_to_double(lo_bits[0]), hi_bits[0] > INF_BITS ? inf : bits_to_double(hi_bits[0]));
      else
        add(0.0, inf);
    }
    return;
  }
  for (int g = 0; g < n_lags; ++g) {  // segments are ascending classes of s
    if (hi_bits[g] <= lo_bits[g]) continue;
    double lo = g ? bits_to_double(thr_bits[g - 1]) : 0.0;
    double hi = bits_to_double(thr_bits[g]);
    if (key_kind == SKG_KEY_SQDIST) {
      lo = std::max(lo, bits_to_double(lo_bits[g]));
      hi = std::min(hi, hi_bits[g] > INF_BITS ? inf : bits_to_double(hi_bits[g]));
    }
    add(lo, hi);
  }
}

static int select_common(bool gather, const double* coords, const double* values, int64_t n,
                         int dim, const double* dir, const uint64_t* thr_bits, int n_lags,
                         int key_kind, const uint64_t* seg_lo_bits, const uint64_t* seg_hi_bits,
                         const double* seg_scale, int64_t n_buckets, uint64_t* hist,
                         uint64_t* max_bits, const uint32_t* bitmap, double* cand_key,
                         uint32_t* cand_slot, uint64_t* cand_count, int64_t capacity,
                         void* scratch, int64_t scratch_bytes, int64_t t0, int64_t t1,
                         int64_t stride, void* stream) {
  int rc = check_pair_args(coords, n, dim, t0, t1, stride);
  if (rc) return rc;
  if (key_kind != SKG_KEY_SQDIST && key_kind != SKG_KEY_ABSDZ) { set_error("bad key_kind"); return SKG_E_BADARG; }
  if (key_kind == SKG_KEY_ABSDZ && !values) { set_error("values required for |dz| keys"); return SKG_E_BADARG; }
  if (n_buckets < 1 || !seg_lo_bits || !seg_hi_bits || !seg_scale) { set_error("bad segment window arguments"); return SKG_E_BADARG; }
  if (n_lags > 0 && !thr_bits) { set_error("thr_bits NULL"); return SKG_E_BADARG; }
  const int nseg = n_lags > 0 ? n_lags : 1;
  if ((double)nseg * (double)n_buckets >= 4294967296.0) { set_error("nseg*n_buckets must fit 32 bits"); return SKG_E_BADARG; }
  BinTable tab;
  rc = build_bin_table(thr_bits, n_lags > 0 ? n_lags : 0, &tab);
  if (rc) return rc;
  DirParams dp;
  rc = build_dir_params(dir, dim, &dp);
  if (rc) return rc;
  SweepScratch S;
  rc = carve_scratch(scratch, scratch_bytes, n, 0, &S);
  if (rc) return rc;
  if (t1 == t0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  SegTable host_seg;
  std::memset(&host_seg, 0, sizeof(host_seg));
  for (int k = 0; k < nseg; ++k) {
    host_seg.lo_bits[k] = seg_lo_bits[k];
    host_seg.hi_bits[k] = seg_hi_bits[k];
    host_seg.scale[k] = seg_scale[k];
  }
  CullWindows win;
  select_windows(n_lags, key_kind, thr_bits, seg_lo_bits, seg_hi_bits,
                 !gather && n_lags <= 0 && max_bits != nullptr, &win);
  const int64_t ntiles = (t1 - t0 + stride - 1) / stride;
  const int js = choose_js(ntiles);
  rc = launch_cull(coords, n, dim, win, t0, t1, stride, js, S, st);
  if (rc) return rc;
  // chunk/warp-level culling inside the sweep uses the hull of all windows (and
  // everything when the maximum is wanted, which only the tile list can narrow)
  double wlo = 0.0, whi = bits_to_double(INF_BITS);
  if (!win.want_max) {
    if (win.count) { wlo = win.lo[0]; whi = win.hi[win.count - 1]; }
    else { wlo = 1.0; whi = 0.0; }  // nothing to find
  }
  if (gather)
    rc = dispatch_select<true>(dim, dp.model == SKG_DIR_NONE ? 0 : (dp.amb_exclude ? 1 : 2), n_lags > 0, key_kind, coords, values, n,
                               tab, dp, host_seg, n_buckets, hist, max_bits, bitmap,
                               cand_key, cand_slot, cand_count, capacity, ntiles,
                               (const unsigned int*)S.list, (const unsigned long long*)S.counters,
                               (const double*)S.bbox, (const double*)S.bbox32, wlo, whi, js, st);
  else
    rc = dispatch_select<false>(dim, dp.model == SKG_DIR_NONE ? 0 : (dp.amb_exclude ? 1 : 2), n_lags > 0, key_kind, coords, values, n,
                                tab, dp, host_seg, n_buckets, hist, max_bits, bitmap,
                                cand_key, cand_slot, cand_count, capacity, ntiles,
                                (const unsigned int*)S.list, (const unsigned long long*)S.counters,
                                (const double*)S.bbox, (const double*)S.bbox32, wlo, whi, js, st);
  return rc;
}

}  // namespace skg

using namespace skg;

extern "C" {

int64_t skg_pair_num_tiles(int64_t n) {
  if (n <= 1) return 0;
  const int64_t nt = (n + TILE - 1) / TILE;
  return nt * (nt + 1) / 2;
}

int64_t skg_pair_scratch_bytes(int64_t n, int n_lags) {
  if (n_lags < 1) n_lags = 1;
  if (n < 0) n = 0;
  return sweep_head_bytes(n) +
         (int64_t)sm_count() * MAX_GRID_PER_SM * n_lags * (int64_t)sizeof(HistPartial);
}

int skg_pair_hist(const double* coords, const double* values, int64_t n, int dim, const double* dir,
                  const uint64_t* thr_bits, int n_lags, int stats, const double* dz_lo,
                  uint64_t* count, double* sum_sq, double* sum_sqrt, void* scratch,
                  int64_t scratch_bytes, int64_t tile_begin, int64_t tile_end, int64_t tile_stride,
                  void* stream) {
  int rc = check_pair_args(coords, n, dim, tile_begin, tile_end, tile_stride);
  if (rc) return rc;
  if (!values || !thr_bits || !count || n_lags < 1) { set_error("values/thr_bits/count NULL or n_lags < 1"); return SKG_E_BADARG; }
  if (stats == SKG_STAT_COUNT_BELOW) {
    if (!dz_lo || !sum_sq) { set_error("below-count mode needs dz_lo and sum_sq (receives the counts)"); return SKG_E_BADARG; }
  } else if ((stats & ~3) || ((stats & SKG_STAT_SUM_SQ) && !sum_sq) || ((stats & SKG_STAT_SUM_SQRT) && !sum_sqrt)) {
    set_error("stats flags / output pointers inconsistent");
    return SKG_E_BADARG;
  }
  const int max_blocks = sm_count() * MAX_GRID_PER_SM;
  SweepScratch S;
  rc = carve_scratch(scratch, scratch_bytes, n, (int64_t)max_blocks * n_lags * (int64_t)sizeof(HistPartial), &S);
  if (rc) return rc;
  BinTable tab;
  rc = build_bin_table(thr_bits, n_lags, &tab);
  if (rc) return rc;
  DirParams dp;
  rc = build_dir_params(dir, dim, &dp);
  if (rc) return rc;
  if (tile_end == tile_begin) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  // tiles entirely beyond the last lag edge cannot contribute
  CullWindows win;
  std::memset(&win, 0, sizeof(win));
  win.count = 1;
  win.lo[0] = 0.0;
  win.hi[0] = bits_to_double(thr_bits[n_lags - 1]);
  const int64_t ntl = (tile_end - tile_begin + tile_stride - 1) / tile_stride;
  const int js = choose_js(ntl);
  rc = launch_cull(coords, n, dim, win, tile_begin, tile_end, tile_stride, js, S, st);
  if (rc) return rc;
  const int use_dir = dp.model == SKG_DIR_NONE ? 0 : (dp.amb_exclude ? 1 : 2);
  HistPartial* partial = reinterpret_cast<HistPartial*>(S.tail);
  int nblocks = 0;
  DzTable dzt;
  std::memset(&dzt, 0, sizeof(dzt));
  if (stats == SKG_STAT_COUNT_BELOW)
    for (int k = 0; k < n_lags; ++k) dzt.lo[k] = dz_lo[k];
  HistLaunch L{coords, values, n, &tab, &dp, partial, S.list, S.counters, S.bbox, S.bbox32,
               win.lo[0], win.hi[0], &dzt, max_blocks, ntl, js, 0, st, &nblocks};
  bool priv = true;
  L.smem = hist_smem_bytes(n_lags, stats, true);
  if (L.smem > 180 * 1024) {
    if (stats == SKG_STAT_COUNT_BELOW) { set_error("below-count mode: too many lags for private records"); return SKG_E_UNSUPPORTED; }
    priv = false;
    L.smem = hist_smem_bytes(n_lags, stats, false);
  }
  rc = dispatch_hist(priv, tab.emax, dim, use_dir, stats, L);
  if (rc) return rc;
  pair_hist_reduce_kernel<<<n_lags, 32, 0, st>>>(
      partial, nblocks, n_lags, reinterpret_cast<unsigned long long*>(count),
      ((stats & SKG_STAT_SUM_SQ) || stats == SKG_STAT_COUNT_BELOW) ? sum_sq : nullptr,
      (stats != SKG_STAT_COUNT_BELOW && (stats & SKG_STAT_SUM_SQRT)) ? sum_sqrt : nullptr);
  SKG_CUDA(cudaGetLastError());
  return 0;
}

int skg_pair_select_hist(const double* coords, const double* values, int64_t n, int dim,
                         const double* dir, const uint64_t* thr_bits, int n_lags, int key_kind,
                         const uint64_t* seg_lo_bits, const uint64_t* seg_hi_bits,
                         const double* seg_scale, int64_t n_buckets, uint64_t* hist,
                         uint64_t* max_bits, void* scratch, int64_t scratch_bytes,
                         int64_t tile_begin, int64_t tile_end, int64_t tile_stride, void* stream) {
  if (!hist) { set_error("hist NULL"); return SKG_E_BADARG; }
  return select_common(false, coords, values, n, dim, dir, thr_bits, n_lags, key_kind, seg_lo_bits,
                       seg_hi_bits, seg_scale, n_buckets, hist, max_bits, nullptr, nullptr, nullptr,
                       nullptr, 0, scratch, scratch_bytes, tile_begin, tile_end, tile_stride, stream);
}

int skg_pair_select_gather(const double* coords, const double* values, int64_t n, int dim,
                           const double* dir, const uint64_t* thr_bits, int n_lags, int key_kind,
                           const uint64_t* seg_lo_bits, const uint64_t* seg_hi_bits,
                           const double* seg_scale, int64_t n_buckets, const uint32_t* bitmap,
                           double* cand_key, uint32_t* cand_slot, uint64_t* cand_count,
                           int64_t capacity, void* scratch, int64_t scratch_bytes,
                           int64_t tile_begin, int64_t tile_end, int64_t tile_stride, void* stream) {
  if (!bitmap || !cand_key || !cand_slot || !cand_count || capacity < 0) {
    set_error("gather buffers NULL");
    return SKG_E_BADARG;
  }
  return select_common(true, coords, values, n, dim, dir, thr_bits, n_lags, key_kind, seg_lo_bits,
                       seg_hi_bits, seg_scale, n_buckets, nullptr, nullptr, bitmap, cand_key,
                       cand_slot, cand_count, capacity, scratch, scratch_bytes, tile_begin, tile_end,
                       tile_stride, stream);
}

int skg_pair_dir_ambiguous(const double* coords, int64_t n, int dim, const double* dir,
                           int32_t* pair_i, int32_t* pair_j, uint64_t* count, int64_t capacity,
                           int64_t tile_begin, int64_t tile_end, int64_t tile_stride, void* stream) {
  int rc = check_pair_args(coords, n, dim, tile_begin, tile_end, tile_stride);
  if (rc) return rc;
  if (!dir || !pair_i || !pair_j || !count || capacity < 0) { set_error("NULL argument"); return SKG_E_BADARG; }
  if (n > 2147483647ll) { set_error("n too large for int32 pair indices"); return SKG_E_UNSUPPORTED; }
  DirParams dp;
  rc = build_dir_params(dir, dim, &dp);
  if (rc) return rc;
  if (dp.model == SKG_DIR_NONE) { set_error("directional model required"); return SKG_E_BADARG; }
  if (tile_end == tile_begin) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ntl = (tile_end - tile_begin + tile_stride - 1) / tile_stride;
  const int64_t grid = std::min<int64_t>((int64_t)sm_count() * MAX_GRID_PER_SM, ntl);
  pair_dir_ambiguous_kernel<<<(unsigned)grid, BLOCK, 0, st>>>(
      coords, n, dp, pair_i, pair_j, reinterpret_cast<unsigned long long*>(count),
      (long long)capacity, tile_begin, tile_end, tile_stride);
  SKG_CUDA(cudaGetLastError());
  return 0;
}

int skg_pair_materialize(const double* coords, const double* values, int64_t n, int dim,
                         const double* dir, const uint64_t* thr_bits, int n_lags, double* dist,
                         double* diffs, int64_t* groups, uint8_t* mask, int64_t pair_begin,
                         int64_t pair_end, void* stream) {
  if (!coords || n < 0) { set_error("coords NULL or n < 0"); return SKG_E_BADARG; }
  if (dim != 2 && dim != 3) { set_error("dim=%d unsupported", dim); return SKG_E_UNSUPPORTED; }
  const int64_t npairs = n * (n - 1) / 2;
  if (pair_begin < 0 || pair_end < pair_begin || pair_end > npairs) { set_error("pair range outside [0, n(n-1)/2]"); return SKG_E_BADARG; }
  if (diffs && !values) { set_error("values required for diffs"); return SKG_E_BADARG; }
  if (groups && (!thr_bits || n_lags < 1)) { set_error("thr_bits required for groups"); return SKG_E_BADARG; }
  BinTable tab;
  int rc = build_bin_table(thr_bits, groups ? n_lags : 0, &tab);
  if (rc) return rc;
  DirParams dp;
  rc = build_dir_params(dir, dim, &dp);
  if (rc) return rc;
  if (pair_end == pair_begin) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t cnt = pair_end - pair_begin;
  const int threads = 256;
  const int64_t blocks = std::min<int64_t>((cnt + threads - 1) / threads, (int64_t)sm_count() * 16);
  const int use_dir = dp.model != SKG_DIR_NONE;
  if (dim == 3)
    pair_materialize_kernel<3><<<(unsigned)blocks, threads, 0, st>>>(
        coords, values, n, tab, dp, use_dir, dist, diffs, (long long*)groups, mask, pair_begin, pair_end);
  else
    pair_materialize_kernel<2><<<(unsigned)blocks, threads, 0, st>>>(
        coords, values, n, tab, dp, use_dir, dist, diffs, (long long*)groups, mask, pair_begin, pair_end);
  SKG_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
