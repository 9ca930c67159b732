// Host side of the pair sweeps (K1-K4 of SURVEY.md section 2.2) for sm_100a: culling prelude
// kernels, the fused-pass instantiations for scalar values, guard-band pair collection,
// materialisation and the C ABI.  The sweep kernels themselves live in pair_sweep.cuh; the
// other instantiations in pair_select.cu (select passes, cell counts), pair_mv.cu (multi-column
// values) and pair_batch.cu (many value vectors per sweep).
//
// All kernels walk the upper triangle of the N x N pair matrix in TILE x TILE point tiles.
// Threads keep their "i" points in registers, the "j" tile (or chunk) sits in shared memory and
// is read back as warp-wide broadcasts.  The N(N-1)/2 distance vector the reference
// materialises (MetricSpace.py:151-153, Variogram.py:1202-1208) never exists.
//
// Per pair:  s = fl(fl(dx*dx) + fl(dy*dy)) [+ fl(dw*dw)]  - scipy's euclidean operand order, no
// FMA (explicit __dmul_rn/__dadd_rn), so that every comparison against the squared-space
// thresholds reproduces np.digitize on d = sqrt(s) exactly (see include/skg_b200.h).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "pair_sweep.cuh"

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
                                            TileSmem<DIM>& sm, Op& op,
                                            const unsigned int* __restrict__ tile_list = nullptr,
                                            const unsigned long long* __restrict__ tile_count = nullptr) {
  const int tid = threadIdx.x;
  const int64_t nt = (n + TILE - 1) / TILE;
  const double qnan = __longlong_as_double(NAN_BITS);
  const double* __restrict__ X = coords;
  const double* __restrict__ Y = coords + n;
  const double* __restrict__ W = coords + 2 * n;

  // with a tile list (culling prelude, one work item per tile) only the listed tiles are swept
  const int64_t nlist = tile_list ? (int64_t)(*tile_count) : 0;
  for (int64_t w = blockIdx.x;; w += gridDim.x) {
    int64_t t;
    if (tile_list) {
      if (w >= nlist) break;
      t = tile_list[w];
    } else {
      t = t0 + w * stride;
      if (t >= t1) break;
    }
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



template <int DIM>
__global__ void __launch_bounds__(256)
block_bbox_kernel(const double* __restrict__ coords, int64_t n, double* __restrict__ bbox,
                  double* __restrict__ bbox32, double* __restrict__ bbox8) {
  __shared__ double sLo[DIM][8], sHi[DIM][8];
  const int64_t i = (int64_t)blockIdx.x * TILE + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double inf = __longlong_as_double((long long)INF_BITS);
  bool bad = i >= n;  // NaN coordinate or padding: the point's 8-group gets no usable box
#pragma unroll
  for (int k = 0; k < DIM; ++k) {
    const double v = i < n ? coords[(int64_t)k * n + i] : __longlong_as_double(NAN_BITS);
    bad |= !(v == v);
    double lo = (v == v) ? v : inf, hi = (v == v) ? v : -inf;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
      if (o == 4 && DIM == 2 && bbox8 && (lane & 7) == 0) {  // boxes of the aligned 8-point groups
        const int64_t g8 = (int64_t)blockIdx.x * (TILE / 8) + (threadIdx.x >> 3);
        bbox8[g8 * 4 + k] = lo;
        bbox8[g8 * 4 + 2 + k] = hi;
      }
    }
    if (lane == 0) {
      sLo[k][warp] = lo; sHi[k][warp] = hi;
      bbox32[((int64_t)blockIdx.x * 8 + warp) * 6 + k] = lo;      // 32-point sub-block boxes
      bbox32[((int64_t)blockIdx.x * 8 + warp) * 6 + 3 + k] = hi;
    }
  }
  if (DIM == 2 && bbox8) {
    // a group with a NaN / missing point is flagged by a NaN lower x: every window test on it
    // fails and its pairs take the per-pair path
    const unsigned badmask = __ballot_sync(0xffffffffu, bad);
    if ((lane & 7) == 0 && ((badmask >> lane) & 0xffu)) {
      const int64_t g8 = (int64_t)blockIdx.x * (TILE / 8) + (threadIdx.x >> 3);
      bbox8[g8 * 4] = __longlong_as_double(NAN_BITS);
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

// Flags + ORDERED compaction in one launch.  A CTA takes the next block of CULL_ITEMS work items
// (atomic ticket, so a block only ever waits for blocks that are already running), decides
// their flags, scans them and finds its output offset by decoupled look-back over the
// published per-block aggregates / inclusive prefixes; the list comes out in ascending item
// order, hence the sweep's static round-robin and fixed-order reductions stay deterministic.
constexpr int CULL_THREADS = 256;
constexpr int CULL_PER_THREAD = 8;
constexpr int CULL_ITEMS = CULL_THREADS * CULL_PER_THREAD;
constexpr unsigned long long SCAN_AGG = 1ull << 62, SCAN_PREFIX = 2ull << 62, SCAN_VAL = (1ull << 62) - 1;

// Count-only sweeps: a work item whose whole range of squared distances lies in ONE class is
// counted here in bulk (rows x columns, no pair is ever evaluated) and dropped from the list.
struct BulkTable {
  int nb;              // 0 = off
  double thr[THR_MAX];  // class thresholds as doubles
};

template <int DIM>
__global__ void __launch_bounds__(CULL_THREADS)
item_cull_kernel(const double* __restrict__ bbox, const double* __restrict__ bbox32, int64_t n,
                 int64_t nt, int64_t t_begin, int64_t ntile, int64_t t_stride, int js,
                 const __grid_constant__ CullWindows win, unsigned long long* __restrict__ counters,
                 unsigned long long* __restrict__ scan /* [0] ticket, [1 + b] block status */,
                 unsigned int* __restrict__ list, const __grid_constant__ BulkTable bulk,
                 unsigned long long* __restrict__ bulk_count) {
  __shared__ int warp_tot[CULL_THREADS / 32];
  __shared__ long long s_base;
  __shared__ unsigned int s_block;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t nitem = ntile * js;
  const int jchunk = TILE / js;
  if (tid == 0) s_block = (unsigned int)atomicAdd(&scan[0], 1ull);
  __syncthreads();
  const int64_t blk = s_block;
  const double need = win.want_max
      ? __longlong_as_double((long long)counters[1]) * (1.0 - 1e-12) : 0.0;
  const int64_t k0 = blk * CULL_ITEMS + (int64_t)tid * CULL_PER_THREAD;
  unsigned int items[CULL_PER_THREAD];
  unsigned keepmask = 0u;
#pragma unroll
  for (int q = 0; q < CULL_PER_THREAD; ++q) {
    const int64_t k = k0 + q;
    int keep = 0;
    items[q] = 0u;
    if (k < nitem) {
      const int64_t t = t_begin + (k / js) * t_stride;
      const int h = (int)(k % js);
      items[q] = (unsigned int)(t * js + h);
      int I, J;
      tile_decode(t, nt, I, J);
      if ((int64_t)J * TILE + (int64_t)h * jchunk < n) {  // the chunk holds points
        double jb[6], a, b;
        chunk_box<DIM>(bbox32, J, h, js, jb);
        subblock_range<DIM>(bbox, I, jb, a, b);
        if (a == a && b == b && a <= b) {
          const double lo = a * (1.0 - 1e-12), hi = b * (1.0 + 1e-12);
          for (int w = 0; w < win.count; ++w) keep |= (lo < win.hi[w] && hi >= win.lo[w]) ? 1 : 0;
          if (win.want_max && hi >= need) keep = 1;
          if (DIM == 2 && keep && win.dir_model != SKG_DIR_NONE &&
              !dir_box_may_pass(bbox + (int64_t)I * 6, jb, win))
            keep = 0;
          if (keep && bulk.nb > 0 && I != J) {
            // class(v) = #{k : thr[k] <= v}; the 1e-12 slack covers the rounding of the box arithmetic
            int c0 = 0, c1 = 0;
            for (int sidx = 0, len = bulk.nb; len > 0;) {  // lower class bound by bisection
              const int half = len >> 1;
              if (bulk.thr[sidx + half] <= lo) { sidx += half + 1; len -= half + 1; } else len = half;
              c0 = sidx;
            }
            for (int sidx = 0, len = bulk.nb; len > 0;) {
              const int half = len >> 1;
              if (bulk.thr[sidx + half] <= hi) { sidx += half + 1; len -= half + 1; } else len = half;
              c1 = sidx;
            }
            if (c0 == c1) {
              const int64_t rows = min((int64_t)TILE, n - (int64_t)I * TILE);
              const int64_t cols = min((int64_t)jchunk, n - ((int64_t)J * TILE + (int64_t)h * jchunk));
              if (c0 < bulk.nb) atomicAdd(&bulk_count[c0], (unsigned long long)(rows * cols));
              keep = 0;
            }
          }
        }
      }
    }
    keepmask |= (unsigned)keep << q;
  }
  const int cnt = __popc(keepmask);
  int x = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_tot[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int v = lane < CULL_THREADS / 32 ? warp_tot[lane] : 0;
#pragma unroll
    for (int o = 1; o < CULL_THREADS / 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += y;
    }
    if (lane < CULL_THREADS / 32) warp_tot[lane] = v;
    const long long total = __shfl_sync(0xffffffffu, v, CULL_THREADS / 32 - 1);
    // publish the aggregate, then look back for the exclusive prefix
    volatile unsigned long long* st = scan + 1;
    long long excl = 0;
    if (blk == 0) {
      if (lane == 0) { st[0] = SCAN_PREFIX | (unsigned long long)total; }
    } else {
      if (lane == 0) { st[blk] = SCAN_AGG | (unsigned long long)total; }
      __threadfence();
      int64_t j = blk - 1;
      while (true) {  // 32 predecessors per step, newest first
        const int64_t idx = j - lane;
        unsigned long long v2 = SCAN_PREFIX;  // lanes below block 0 act as an empty prefix
        if (idx >= 0) {
          do { v2 = st[idx]; } while ((v2 >> 62) == 0ull);
        }
        const unsigned pref = __ballot_sync(0xffffffffu, (v2 >> 62) == 2ull);
        const int first = pref ? (__ffs(pref) - 1) : 32;       // nearest block with a full prefix
        long long add = (lane <= first) ? (long long)(v2 & SCAN_VAL) : 0ll;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) add += __shfl_xor_sync(0xffffffffu, add, o);
        excl += add;
        if (pref) break;
        j -= 32;
      }
      if (lane == 0) {
        __threadfence();
        st[blk] = SCAN_PREFIX | (unsigned long long)(excl + total);
      }
    }
    if (lane == 0) {
      s_base = excl;
      if ((blk + 1) * CULL_ITEMS >= nitem) {  // the last block knows the grand total
        counters[0] = (unsigned long long)(excl + total);  // surviving work items
        counters[2] = (unsigned long long)js;              // j-chunks per tile (for reporting)
        counters[3] = (unsigned long long)nitem;           // candidate work items before culling
      }
    }
  }
  __syncthreads();
  long long pos = s_base + (warp ? warp_tot[warp - 1] : 0) + (x - cnt);
#pragma unroll
  for (int q = 0; q < CULL_PER_THREAD; ++q)
    if ((keepmask >> q) & 1u) list[pos++] = items[q];
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
                          long long capacity, int64_t t0, int64_t t1, int64_t stride,
                          const unsigned int* __restrict__ tile_list,
                          const unsigned long long* __restrict__ tile_count) {
  __shared__ TileSmem<2> sm;
  AmbOp op;
  op.dp = &dp; op.out_i = out_i; op.out_j = out_j; op.count = count; op.capacity = capacity;
  sweep_tiles<2>(coords, nullptr, n, t0, t1, stride, sm, op, tile_list, tile_count);
}

// ------------------------------------------------------- materialisation ---

template <int DIM>
__global__ void pair_materialize_kernel(const double* __restrict__ coords,
                                        const double* __restrict__ values, const MvSpec mv,
                                        int64_t n, const __grid_constant__ BinTable tab,
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
    if (diffs) {
      if (mv.mode == SKG_VALUE_CROSS) {
        diffs[o] = __dmul_rn(fabs(values[i] - values[j]), fabs(values[n + i] - values[n + j]));
      } else if (mv.mode == SKG_VALUE_AGGREGATE) {
        double acc = 0.0;
        for (int c = 0; c < mv.nv; ++c) {
          const double d = values[(int64_t)c * n + i] - values[(int64_t)c * n + j];
          acc = __dadd_rn(acc, __dmul_rn(d, d));
        }
        diffs[o] = sqrt(acc);
      } else {
        diffs[o] = fabs(values[i] - values[j]);
      }
    }
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

int check_pair_args(const double* coords, int64_t n, int dim, int64_t t0, int64_t t1,
                    int64_t stride) {
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

// value columns: [n_values][n]; mode SKG_VALUE_*
static int check_values(int n_values, int value_mode, MvSpec* mv) {
  mv->nv = n_values;
  mv->square = (value_mode & SKG_VALUE_SQUARE_FLAG) ? 1 : 0;
  value_mode &= ~SKG_VALUE_SQUARE_FLAG;
  mv->mode = value_mode;
  if (value_mode == SKG_VALUE_SCALAR && n_values == 1) return 0;
  if (value_mode == SKG_VALUE_CROSS && n_values == 2) return 0;
  if (value_mode == SKG_VALUE_AGGREGATE && n_values >= 1 && n_values <= SKG_MAX_VALUE_COLUMNS) return 0;
  if (value_mode >= SKG_VALUE_DIST_S && value_mode <= SKG_VALUE_DIST_D15 && n_values == 0) return 0;
  set_error("n_values=%d / value_mode=%d: scalar needs 1 column, cross 2, aggregate 1..%d", n_values,
            value_mode, SKG_MAX_VALUE_COLUMNS);
  return SKG_E_BADARG;
}

constexpr int64_t LIST_SPLIT_CAP = 1 << 22;  // j-splitting only while tiles * js fits this

static int64_t list_entries(int64_t ntiles) { return std::max<int64_t>(ntiles, std::min<int64_t>(ntiles * 16, LIST_SPLIT_CAP)); }

static int64_t scan_entries(int64_t ntiles) { return list_entries(ntiles) / CULL_ITEMS + 4; }

int64_t sweep_head_bytes(int64_t n) {
  const int64_t nt = (n + TILE - 1) / TILE;
  const int64_t ntiles = nt * (nt + 1) / 2;
  int64_t b = 256 + nt * 6 * 8 + nt * 8 * 6 * 8 + nt * (TILE / 8) * 4 * 8 + scan_entries(ntiles) * 8 +
              list_entries(ntiles) * 4;
  return (b + 255) & ~255ll;
}

int carve_scratch(void* scratch, int64_t scratch_bytes, int64_t n, int64_t need_tail,
                  SweepScratch* out) {
  const int64_t head = sweep_head_bytes(n);
  if (!scratch || scratch_bytes < head + need_tail) {
    set_error("scratch too small: need %lld bytes", (long long)(head + need_tail));
    return SKG_E_SCRATCH;
  }
  const int64_t nt = (n + TILE - 1) / TILE;
  const int64_t ntiles = nt * (nt + 1) / 2;
  unsigned char* base = reinterpret_cast<unsigned char*>(scratch);
  out->counters = reinterpret_cast<unsigned long long*>(base);
  out->bbox = reinterpret_cast<double*>(base + 256);
  out->bbox32 = out->bbox + nt * 6;
  out->bbox8 = out->bbox32 + nt * 8 * 6;
  out->scan = reinterpret_cast<unsigned long long*>(out->bbox8 + nt * (TILE / 8) * 4);
  out->scan_bytes = scan_entries(ntiles) * 8;
  out->list = reinterpret_cast<unsigned int*>(out->scan + scan_entries(ntiles));
  out->tail = base + head;
  out->tail_bytes = scratch_bytes - head;
  return 0;
}

// bounding boxes of the 256-point blocks + ordered list of the tiles worth sweeping
// j-split: enough (tile, chunk) work items per CTA that culling and the static round-robin
// stay balanced; bounded by the list capacity carved from the scratch
int choose_js(int64_t tiles) {
  const int64_t want = (int64_t)sm_count() * 6 * 12;
  int js = 1;
  while (js < 16 && tiles * js < want && tiles * (js * 2) <= LIST_SPLIT_CAP) js *= 2;
  if (const char* f = getenv("SKG_FORCE_JS")) {  // developer override
    const int fj = atoi(f);
    if (fj >= 1 && fj <= 16 && (fj & (fj - 1)) == 0 && tiles * fj <= LIST_SPLIT_CAP) js = fj;
  }
  return js;
}

int launch_cull(const double* coords, int64_t n, int dim, const CullWindows& win,
                int64_t t0, int64_t t1, int64_t stride, int js, const SweepScratch& S,
                cudaStream_t st, const BinTable* bulk_tab, unsigned long long* bulk_count) {
  BulkTable bulk;  // passed to the kernel by value
  bulk.nb = 0;
  if (bulk_tab && bulk_count) {
    bulk.nb = bulk_tab->nb;
    for (int k = 0; k < bulk_tab->nb; ++k) std::memcpy(&bulk.thr[k], &bulk_tab->thr[k], 8);
  }
  const int64_t nt = (n + TILE - 1) / TILE;
  const int64_t ntile = t1 > t0 ? (t1 - t0 + stride - 1) / stride : 0;
  const int64_t nitem = ntile * js;
  const unsigned tb = (unsigned)std::min<int64_t>((ntile + 255) / 256, (int64_t)sm_count() * 8);
  const int64_t cb = (nitem + CULL_ITEMS - 1) / CULL_ITEMS;
  // counters[0..7] (list length, max bits, js, nitem, window group size, ...) + ticket/status
  SKG_CUDA(cudaMemsetAsync(S.counters, 0, 128, st));
  SKG_CUDA(cudaMemsetAsync(S.scan, 0, (size_t)std::min<int64_t>(S.scan_bytes, (cb + 1) * 8), st));
  if (nitem == 0) return 0;
  if (dim == 3) {
    block_bbox_kernel<3><<<(unsigned)nt, 256, 0, st>>>(coords, n, S.bbox, S.bbox32, nullptr);
    if (win.want_max) tile_need_kernel<3><<<tb, 256, 0, st>>>(S.bbox, nt, t0, ntile, stride, S.counters);
    item_cull_kernel<3><<<(unsigned)cb, CULL_THREADS, 0, st>>>(S.bbox, S.bbox32, n, nt, t0, ntile, stride,
                                                              js, win, S.counters, S.scan, S.list, bulk,
                                                              bulk_count);
  } else {
    block_bbox_kernel<2><<<(unsigned)nt, 256, 0, st>>>(coords, n, S.bbox, S.bbox32, S.bbox8);
    if (win.want_max) tile_need_kernel<2><<<tb, 256, 0, st>>>(S.bbox, nt, t0, ntile, stride, S.counters);
    item_cull_kernel<2><<<(unsigned)cb, CULL_THREADS, 0, st>>>(S.bbox, S.bbox32, n, nt, t0, ntile, stride,
                                                              js, win, S.counters, S.scan, S.list, bulk,
                                                              bulk_count);
  }
  SKG_CUDA(cudaGetLastError());
  return 0;
}

static double bits_to_double(uint64_t b) {
  double d;
  std::memcpy(&d, &b, 8);
  return d;
}

static size_t hist_smem_bytes(int nb, int stats, bool priv) {
  if (!priv) return (size_t)nb * (stats == 3 ? 32 : 16);
  const int sumb = stats == 3 ? 16 : (stats == 4 ? 4 : (stats == 0 ? 0 : 8));
  return ((size_t)nb * HB * (sumb + 4) + 15) & ~(size_t)15;
}

template <int DIM, int DIR, int EMAX, bool PRIV>
static int dispatch_hist_stats(int stats, const HistLaunch& L) {
  switch (stats) {
    case 0: return launch_hist<DIM, DIR, 0, EMAX, PRIV, false>(L);
    case 1: return launch_hist<DIM, DIR, 1, EMAX, PRIV, false>(L);
    case 2: return launch_hist<DIM, DIR, 2, EMAX, PRIV, false>(L);
    case 4: return launch_hist<DIM, DIR, 4, EMAX, PRIV, false>(L);
    default: return launch_hist<DIM, DIR, 3, EMAX, PRIV, false>(L);
  }
}

template <int EMAX, bool PRIV>
static int dispatch_hist_dim(int dim, int dir, int stats, const HistLaunch& L) {
  if (dim == 3) return dispatch_hist_stats<3, 0, EMAX, PRIV>(stats, L);
  if (dir == 1) return dispatch_hist_stats<2, 1, EMAX, PRIV>(stats, L);
  if (dir == 2) return dispatch_hist_stats<2, 2, EMAX, PRIV>(stats, L);
  return dispatch_hist_stats<2, 0, EMAX, PRIV>(stats, L);
}

int dispatch_hist(bool priv, int emax, int dim, int dir, int stats, const HistLaunch& L) {
  if (!priv) return dispatch_hist_dim<0, false>(dim, dir, stats, L);
  if (emax <= 1) return dispatch_hist_dim<1, true>(dim, dir, stats, L);
  return dispatch_hist_dim<0, true>(dim, dir, stats, L);
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

static int select_common(bool gather, const double* coords, const double* values, int n_values,
                         int value_mode, int64_t n, int dim, const double* dir, const uint64_t* thr_bits, int n_lags,
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
  MvSpec mv;
  rc = check_values(n_values, value_mode, &mv);
  if (rc) return rc;
  const bool use_mv = key_kind == SKG_KEY_ABSDZ && value_mode != SKG_VALUE_SCALAR;
  if (use_mv && n_lags <= 0) { set_error("multi-column |dz| keys need lag classes"); return SKG_E_UNSUPPORTED; }
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
  cull_set_dir(&win, dp);
  rc = launch_cull(coords, n, dim, win, t0, t1, stride, js, S, st);
  if (rc) return rc;
  // chunk/warp-level culling inside the sweep uses the hull of all windows (and
  // everything when the maximum is wanted, which only the tile list can narrow)
  double wlo = 0.0, whi = bits_to_double(INF_BITS);
  if (!win.want_max) {
    if (win.count) { wlo = win.lo[0]; whi = win.hi[win.count - 1]; }
    else { wlo = 1.0; whi = 0.0; }  // nothing to find
  }
  const int use_dir = dp.model == SKG_DIR_NONE ? 0 : (dp.amb_exclude ? 1 : 2);
  SelectLaunch L{coords, values, n, &tab, &dp, &host_seg, n_buckets, hist, max_bits, bitmap,
                 cand_key, cand_slot, cand_count, capacity, ntiles, S.list, S.counters, S.bbox,
                 S.bbox32, wlo, whi, js, st, mv, getenv("SKG_GROUP_CULL_OFF") ? nullptr : S.bbox8, -1, 0.0, 0.0};
  if (key_kind == SKG_KEY_SQDIST && !use_mv && getenv("SKG_FAST_SEG_OFF") == nullptr) {
    // exactly one segment with a window, lying inside that segment's class (or no classes at all):
    // membership is lo <= s < hi, no lag-class lookup needed
    int only = -1, cnt = 0;
    for (int k = 0; k < nseg; ++k)
      if (seg_hi_bits[k] > seg_lo_bits[k]) { only = k; ++cnt; }
    if (cnt == 1 && seg_hi_bits[only] <= INF_BITS + 1) {
      const uint64_t c_lo = (n_lags > 0 && only > 0) ? thr_bits[only - 1] : 0ull;
      const uint64_t c_hi = n_lags > 0 ? thr_bits[only] : INF_BITS + 1;
      const uint64_t lo = std::max(seg_lo_bits[only], c_lo), hi = std::min(seg_hi_bits[only], c_hi);
      if (hi > lo && !(n_lags <= 0 && max_bits != nullptr && !gather)) {
        L.fast_seg = only;
        L.fast_lo = bits_to_double(lo);
        L.fast_hi = hi > INF_BITS ? bits_to_double(INF_BITS) : bits_to_double(hi);
        if (hi > INF_BITS) L.fast_seg = -1;  // an unbounded window keeps the general path (s = +inf)
      }
    }
  }
  if (gather && L.bbox8 && dz_gather_applicable(dim, use_dir, n_lags > 0, key_kind, use_mv, host_seg, nseg))
    return launch_dz_gather(L);   // per-lag |dz| windows: value test first (pair_dzgather.cu)
  rc = use_mv ? dispatch_select_mv(gather, dim, use_dir, L)
              : dispatch_select(gather, dim, use_dir, n_lags > 0, key_kind, L);
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
         (int64_t)sm_count() * MAX_GRID_PER_SM * n_lags * (int64_t)sizeof(HistPartial) + (int64_t)sizeof(DzTable) + 256;
}

static int pair_hist_impl(const double* coords, const double* values, int n_values, int value_mode,
                          int64_t n, int dim, const double* dir,
                          const uint64_t* thr_bits, int n_lags, int stats, const double* dz_lo,
                          const double* dz_hi, const double* dz_scale, int64_t n_buckets, uint64_t* whist,
                          uint64_t* count, double* sum_sq, double* sum_sqrt, void* scratch,
                          int64_t scratch_bytes, int64_t tile_begin, int64_t tile_end, int64_t tile_stride,
                          void* stream) {
  int rc = check_pair_args(coords, n, dim, tile_begin, tile_end, tile_stride);
  if (rc) return rc;
  const bool dist_mode = (value_mode & ~SKG_VALUE_SQUARE_FLAG) >= SKG_VALUE_DIST_S &&
                         (value_mode & ~SKG_VALUE_SQUARE_FLAG) <= SKG_VALUE_DIST_D15;
  if ((!values && !dist_mode) || !thr_bits || !count || n_lags < 1) { set_error("values/thr_bits/count NULL or n_lags < 1"); return SKG_E_BADARG; }
  if (stats == SKG_STAT_COUNT_BELOW) {
    if (!dz_lo || !sum_sq) { set_error("below-count mode needs dz_lo and sum_sq (receives the counts)"); return SKG_E_BADARG; }
  } else if ((stats & ~3) || ((stats & SKG_STAT_SUM_SQ) && !sum_sq) || ((stats & SKG_STAT_SUM_SQRT) && !sum_sqrt)) {
    set_error("stats flags / output pointers inconsistent");
    return SKG_E_BADARG;
  }
  MvSpec mv;
  rc = check_values(n_values, value_mode, &mv);
  if (rc) return rc;
  const bool use_mv = (value_mode != SKG_VALUE_SCALAR) && stats != 0;  // any flag or multi-column mode
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
  cull_set_dir(&win, dp);
  // count-only sweeps without a mask: whole work items inside one class are counted in bulk
  const bool bulk = stats == 0 && dp.model == SKG_DIR_NONE && getenv("SKG_BULK_OFF") == nullptr;
  rc = launch_cull(coords, n, dim, win, tile_begin, tile_end, tile_stride, js, S, st,
                   bulk ? &tab : nullptr, bulk ? reinterpret_cast<unsigned long long*>(count) : nullptr);
  if (rc) return rc;
  const int use_dir = dp.model == SKG_DIR_NONE ? 0 : (dp.amb_exclude ? 1 : 2);
  HistPartial* partial = reinterpret_cast<HistPartial*>(S.tail);
  int nblocks = 0;
  DzTable dzt;
  std::memset(&dzt, 0, sizeof(dzt));
  if (stats == SKG_STAT_COUNT_BELOW)
    for (int k = 0; k < n_lags; ++k) {
      dzt.lo[k] = dz_lo[k];
      dzt.hi[k] = whist ? dz_hi[k] : 0.0;
      dzt.scale[k] = whist ? dz_scale[k] : 0.0;
    }
  HistLaunch L{coords, values, n, &tab, &dp, partial, S.list, S.counters, S.bbox, S.bbox32,
               win.lo[0], win.hi[0], &dzt, max_blocks, ntl, js, 0, st, &nblocks, mv, S.bbox8,
               reinterpret_cast<unsigned long long*>(whist), (long long)n_buckets, nullptr};
  // register lag windows: 2-D isotropic sweeps over scalar values with an estimator sum
  const bool win_off = getenv("SKG_WIN_OFF") != nullptr;  // developer switch: first kernel only
  if (!win_off && dim == 2 && dp.model == SKG_DIR_NONE && !use_mv && stats >= 0 && stats <= 4 &&
      !(stats == 4 && S.tail_bytes < (int64_t)max_blocks * n_lags * (int64_t)sizeof(HistPartial) + (int64_t)sizeof(DzTable))) {
    L.smem = hist_win_smem_bytes(n_lags, stats);
    if (L.smem <= 160 * 1024) {
      if (stats == 4) {
        // the |dz| window table travels through the scratch tail (6 KB: too large next to the lag
        // table for the kernel-parameter space); stream-ordered copy from the caller-visible struct
        unsigned char* dst = S.tail + (int64_t)max_blocks * n_lags * (int64_t)sizeof(HistPartial);
        SKG_CUDA(cudaMemcpyAsync(dst, &dzt, sizeof(DzTable), cudaMemcpyHostToDevice, st));
        L.dzt_dev = reinterpret_cast<const DzTable*>(dst);
      }
      rc = dispatch_hist_win(tab.emax, stats, L);
      if (rc) return rc;
      pair_hist_reduce_kernel<<<n_lags, 32, 0, st>>>(
          partial, nblocks, n_lags, reinterpret_cast<unsigned long long*>(count),
          ((stats & SKG_STAT_SUM_SQ) || stats == 4) ? sum_sq : nullptr,
          (stats != 4 && (stats & SKG_STAT_SUM_SQRT)) ? sum_sqrt : nullptr);
      SKG_CUDA(cudaGetLastError());
      return 0;
    }
  }
  bool priv = true;
  L.smem = hist_smem_bytes(n_lags, stats, true) + (use_mv ? mv_smem_bytes(mv) : 0);
  if (L.smem > 180 * 1024) {
    if (stats == SKG_STAT_COUNT_BELOW) { set_error("below-count mode: too many lags for private records"); return SKG_E_UNSUPPORTED; }
    if (use_mv) { set_error("multi-column values: too many lags for private records"); return SKG_E_UNSUPPORTED; }
    priv = false;
    L.smem = hist_smem_bytes(n_lags, stats, false);
  }
  rc = use_mv ? dispatch_hist_mv(dim, use_dir, stats, L)
              : dispatch_hist(priv, tab.emax, dim, use_dir, stats, L);
  if (rc) return rc;
  pair_hist_reduce_kernel<<<n_lags, 32, 0, st>>>(
      partial, nblocks, n_lags, reinterpret_cast<unsigned long long*>(count),
      ((stats & SKG_STAT_SUM_SQ) || stats == SKG_STAT_COUNT_BELOW) ? sum_sq : nullptr,
      (stats != SKG_STAT_COUNT_BELOW && (stats & SKG_STAT_SUM_SQRT)) ? sum_sqrt : nullptr);
  SKG_CUDA(cudaGetLastError());
  return 0;
}

int skg_pair_hist(const double* coords, const double* values, int n_values, int value_mode,
                  int64_t n, int dim, const double* dir,
                  const uint64_t* thr_bits, int n_lags, int stats, const double* dz_lo,
                  uint64_t* count, double* sum_sq, double* sum_sqrt, void* scratch,
                  int64_t scratch_bytes, int64_t tile_begin, int64_t tile_end, int64_t tile_stride,
                  void* stream) {
  return pair_hist_impl(coords, values, n_values, value_mode, n, dim, dir, thr_bits, n_lags, stats, dz_lo,
                        nullptr, nullptr, 0, nullptr, count, sum_sq, sum_sqrt, scratch, scratch_bytes,
                        tile_begin, tile_end, tile_stride, stream);
}

int skg_pair_dz_windows(const double* coords, const double* values, int n_values, int value_mode,
                        int64_t n, int dim, const double* dir, const uint64_t* thr_bits, int n_lags,
                        const uint64_t* dz_lo_bits, const uint64_t* dz_hi_bits, const double* dz_scale,
                        int64_t n_buckets, uint64_t* count, double* below, uint64_t* hist, void* scratch,
                        int64_t scratch_bytes, int64_t tile_begin, int64_t tile_end, int64_t tile_stride,
                        void* stream) {
  if (!dz_lo_bits || !dz_hi_bits || !dz_scale || !hist || !below || n_buckets < 1 || n_lags < 1 ||
      n_lags > SKG_MAX_LAGS) {
    set_error("window arguments NULL, n_buckets < 1 or n_lags outside [1, %d]", SKG_MAX_LAGS);
    return SKG_E_BADARG;
  }
  double lo[SKG_MAX_LAGS], hi[SKG_MAX_LAGS];
  for (int k = 0; k < n_lags; ++k) {
    std::memcpy(&lo[k], &dz_lo_bits[k], 8);
    std::memcpy(&hi[k], &dz_hi_bits[k], 8);
    if (dz_hi_bits[k] > INF_BITS) hi[k] = bits_to_double(INF_BITS);
  }
  return pair_hist_impl(coords, values, n_values, value_mode, n, dim, dir, thr_bits, n_lags,
                        SKG_STAT_COUNT_BELOW, lo, hi, dz_scale, n_buckets, hist, count, below, nullptr,
                        scratch, scratch_bytes, tile_begin, tile_end, tile_stride, stream);
}

int skg_pair_select_hist(const double* coords, const double* values, int n_values, int value_mode,
                         int64_t n, int dim,
                         const double* dir, const uint64_t* thr_bits, int n_lags, int key_kind,
                         const uint64_t* seg_lo_bits, const uint64_t* seg_hi_bits,
                         const double* seg_scale, int64_t n_buckets, uint64_t* hist,
                         uint64_t* max_bits, void* scratch, int64_t scratch_bytes,
                         int64_t tile_begin, int64_t tile_end, int64_t tile_stride, void* stream) {
  if (!hist) { set_error("hist NULL"); return SKG_E_BADARG; }
  return select_common(false, coords, values, n_values, value_mode, n, dim, dir, thr_bits, n_lags, key_kind, seg_lo_bits,
                       seg_hi_bits, seg_scale, n_buckets, hist, max_bits, nullptr, nullptr, nullptr,
                       nullptr, 0, scratch, scratch_bytes, tile_begin, tile_end, tile_stride, stream);
}

int skg_pair_select_gather(const double* coords, const double* values, int n_values,
                           int value_mode, int64_t n, int dim,
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
  return select_common(true, coords, values, n_values, value_mode, n, dim, dir, thr_bits, n_lags, key_kind, seg_lo_bits,
                       seg_hi_bits, seg_scale, n_buckets, nullptr, nullptr, bitmap, cand_key,
                       cand_slot, cand_count, capacity, scratch, scratch_bytes, tile_begin, tile_end,
                       tile_stride, stream);
}

int skg_pair_dir_ambiguous(const double* coords, int64_t n, int dim, const double* dir,
                           int32_t* pair_i, int32_t* pair_j, uint64_t* count, int64_t capacity,
                           void* scratch, int64_t scratch_bytes,
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
  const unsigned int* list = nullptr;
  const unsigned long long* list_count = nullptr;
  if (scratch) {  // optional: only the tiles whose difference vectors can come near the mask boundary
    SweepScratch S;
    rc = carve_scratch(scratch, scratch_bytes, n, 0, &S);
    if (rc) return rc;
    CullWindows win;
    std::memset(&win, 0, sizeof(win));
    win.count = 1;
    win.lo[0] = 0.0;
    win.hi[0] = bits_to_double(INF_BITS);
    cull_set_dir(&win, dp);
    rc = launch_cull(coords, n, dim, win, tile_begin, tile_end, tile_stride, 1, S, st);
    if (rc) return rc;
    list = S.list;
    list_count = S.counters;
  }
  pair_dir_ambiguous_kernel<<<(unsigned)grid, BLOCK, 0, st>>>(
      coords, n, dp, pair_i, pair_j, reinterpret_cast<unsigned long long*>(count),
      (long long)capacity, tile_begin, tile_end, tile_stride, list, list_count);
  SKG_CUDA(cudaGetLastError());
  return 0;
}

int skg_pair_materialize(const double* coords, const double* values, int n_values, int value_mode,
                         int64_t n, int dim,
                         const double* dir, const uint64_t* thr_bits, int n_lags, double* dist,
                         double* diffs, int64_t* groups, uint8_t* mask, int64_t pair_begin,
                         int64_t pair_end, void* stream) {
  if (!coords || n < 0) { set_error("coords NULL or n < 0"); return SKG_E_BADARG; }
  if (dim != 2 && dim != 3) { set_error("dim=%d unsupported", dim); return SKG_E_UNSUPPORTED; }
  const int64_t npairs = n * (n - 1) / 2;
  if (pair_begin < 0 || pair_end < pair_begin || pair_end > npairs) { set_error("pair range outside [0, n(n-1)/2]"); return SKG_E_BADARG; }
  if (diffs && !values) { set_error("values required for diffs"); return SKG_E_BADARG; }
  MvSpec mv;
  {
    const int rcv = check_values(n_values, value_mode, &mv);
    if (rcv) return rcv;
  }
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
        coords, values, mv, n, tab, dp, use_dir, dist, diffs, (long long*)groups, mask, pair_begin, pair_end);
  else
    pair_materialize_kernel<2><<<(unsigned)blocks, threads, 0, st>>>(
        coords, values, mv, n, tab, dp, use_dir, dist, diffs, (long long*)groups, mask, pair_begin, pair_end);
  SKG_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
