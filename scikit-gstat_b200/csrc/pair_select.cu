// Instantiations of the select-pass kernels (K1/K3) for scalar values.
#include <cstring>

#include "pair_sweep.cuh"

namespace skg {

template <int DIM, int DIR, bool SEG, int KEY, bool GATHER>
static int select_emax(const SelectLaunch& L) {
  if (L.tab->emax <= 1) return launch_select<DIM, DIR, SEG, KEY, GATHER, 1, false>(L);
  return launch_select<DIM, DIR, SEG, KEY, GATHER, 0, false>(L);
}

template <int DIM, int DIR, bool GATHER>
static int select_seg_key(bool seg, int key, const SelectLaunch& L) {
  if (seg) {
    if (key == SKG_KEY_SQDIST) return select_emax<DIM, DIR, true, SKG_KEY_SQDIST, GATHER>(L);
    return select_emax<DIM, DIR, true, SKG_KEY_ABSDZ, GATHER>(L);
  }
  if (key == SKG_KEY_SQDIST) return select_emax<DIM, DIR, false, SKG_KEY_SQDIST, GATHER>(L);
  return select_emax<DIM, DIR, false, SKG_KEY_ABSDZ, GATHER>(L);
}

template <bool GATHER>
static int select_dim(int dim, int dir, bool seg, int key, const SelectLaunch& L) {
  if (dim == 3) return select_seg_key<3, 0, GATHER>(seg, key, L);
  if (dir == 1) return select_seg_key<2, 1, GATHER>(seg, key, L);
  if (dir == 2) return select_seg_key<2, 2, GATHER>(seg, key, L);
  return select_seg_key<2, 0, GATHER>(seg, key, L);
}

int dispatch_select(bool gather, int dim, int dir, bool seg, int key, const SelectLaunch& L) {
  return gather ? select_dim<true>(dim, dir, seg, key, L) : select_dim<false>(dim, dir, seg, key, L);
}

}  // namespace skg

// ------------------------------------------------------- cell counts ---
//
// Exact pair counts on a log-spaced grid of squared distance: cell(s) = (hi32(bits(s)) >> shift)
// - key0, i.e. NC cells per sweep with 2^(20-shift) cells per octave and NO threshold compares.
// The first sweep of every distance order statistic (maxlag='median', uniform-count bins,
// bandwidth percentiles): the cell that holds a wanted rank is so narrow (0.07 % of d at
// shift 11) that its keys can be gathered directly or refined by sweeps that cull nearly all
// tiles.  Counters are shared-memory u32 atomics (lanes of a warp hold neighbouring points, so
// they spread over neighbouring cells), flushed as per-CTA partials and folded in a fixed order.

namespace skg {

constexpr int CELLS = SKG_CELL_COUNT;

template <int DIM, int DIR>
__global__ void __launch_bounds__(HB)
pair_cellcount_kernel(const double* __restrict__ coords, int64_t n,
                      const __grid_constant__ DirParams dp, unsigned long long lo_bits,
                      unsigned long long hi_bits, int shift, int key0,
                      unsigned long long* __restrict__ count, const unsigned int* __restrict__ tile_list,
                      const unsigned long long* __restrict__ tile_count,
                      const double* __restrict__ bbox32, double wlo, double whi, int js,
                      unsigned long long* __restrict__ amb_count) {
  __shared__ double2 sXY[TILE];
  __shared__ double sW[DIM == 3 ? TILE : 1];
  __shared__ unsigned int sCnt[CELLS];
  const int tid = threadIdx.x;
  const double qnan = __longlong_as_double(NAN_BITS);
  for (int k = tid; k < CELLS; k += HB) sCnt[k] = 0u;
  __syncthreads();
  const int64_t nt = (n + TILE - 1) / TILE;
  const double* __restrict__ X = coords;
  const double* __restrict__ Y = coords + n;
  const double* __restrict__ W = coords + 2 * n;
  const int jchunk = TILE / js;
  const int64_t nwork = (int64_t)(*tile_count);
  unsigned int namb = 0u;  // guard-band pairs seen (DIR == 1: they are excluded here and resolved by the caller)
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
    double xi[HR], yi[HR], wi[HR];
#pragma unroll
    for (int r = 0; r < HR; ++r) {
      const int64_t i = (int64_t)I * TILE + r * HB + tid;
      const bool v = i < n;
      xi[r] = v ? X[i] : qnan;
      yi[r] = v ? Y[i] : qnan;
      wi[r] = (DIM == 3) ? (v ? W[i] : qnan) : 0.0;
    }
    __syncthreads();
    for (int k = tid; k < jcount; k += HB) {
      const int64_t j = (int64_t)J * TILE + jbeg + k;
      sXY[k] = make_double2(X[j], Y[j]);
      if (DIM == 3) sW[k] = W[j];
    }
    __syncthreads();
    if (jcount == 0 || !wact) continue;
    const bool diag = (I == J);
#pragma unroll 2
    for (int jj = 0; jj < jcount; ++jj) {
      const double2 pj = sXY[jj];
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
        const unsigned long long sb = (unsigned long long)__double_as_longlong(s);
        bool in = sb >= lo_bits && sb < hi_bits;             // NaN bit patterns lie above hi_bits
        if (diag) in = in && (jbeg + jj > r * HB + tid);     // keep i < j only
        if (DIR == 1) {
          bool amb = false;
          const bool ok = (s > 0.0) && dir_fast(dx, dy, dp, &amb);
          if (amb && s > 0.0 && (!diag || jbeg + jj > r * HB + tid)) ++namb;
          in = in && ok && !amb;
        } else if (DIR) {
          in = in && dir_mask_t<true>(dx, dy, s, dp);
        }
        if (in) {
          const int c = (int)((unsigned)(sb >> 32) >> shift) - key0;
          atomicAdd(&sCnt[min(max(c, 0), CELLS - 1)], 1u);
        }
      }
    }
  }
  __syncthreads();
  // integer counts: the order of the global atomics does not matter; a CTA touches only the
  // cells its tiles' distances fall into
  for (int k = tid; k < CELLS; k += HB) {
    const unsigned int v = sCnt[k];
    if (v) atomicAdd(&count[k], (unsigned long long)v);
  }
  if (DIR == 1 && namb) atomicAdd(amb_count, (unsigned long long)namb);
}

template <int DIM, int DIR>
static int launch_cellcount(const double* coords, int64_t n, const DirParams& dp, uint64_t lo_bits,
                            uint64_t hi_bits, int shift, int key0, unsigned long long* count,
                            int64_t ntiles, int js, const SweepScratch& S, double wlo, double whi,
                            cudaStream_t st) {
  auto kern = pair_cellcount_kernel<DIM, DIR>;
  int occ = 0;
  SKG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, HB, 0));
  occ = occ < 1 ? 1 : (occ < MAX_GRID_PER_SM ? occ : MAX_GRID_PER_SM);
  int64_t grid = (int64_t)sm_count() * occ;
  if (grid > ntiles * js) grid = ntiles * js;
  kern<<<(unsigned)grid, HB, 0, st>>>(coords, n, dp, lo_bits, hi_bits, shift, key0, count, S.list,
                                      S.counters, S.bbox32, wlo, whi, js, S.counters + 6);
  SKG_CUDA(cudaGetLastError());
  return 0;
}

// Order statistics of gathered candidates without a sort: candidate i of a wanted slot finds its
// rank by counting the candidates of the same slot below it (ties broken by position), and the
// one whose rank is wanted writes its key.  O(m^2) compares over the whole grid: for the few
// thousand keys a narrowed select window leaves this is a few microseconds and replaces a dozen
// library launches (two sorts, two searches, index arithmetic) of the host-driven pick.
constexpr int PICK_THREADS = 256;
constexpr int PICK_MAX_WANT = 32;
constexpr int PICK_MAX_RANKS = 64;
struct PickPlan {
  int n_want;
  int slot[PICK_MAX_WANT];
  int off[PICK_MAX_WANT + 1];
  long long rank[PICK_MAX_RANKS];
};

__global__ void __launch_bounds__(PICK_THREADS)
pick_ranks_kernel(const double* __restrict__ keys, const int* __restrict__ slots,
                  const unsigned long long* __restrict__ count, long long cap, PickPlan plan,
                  double* __restrict__ out) {
  __shared__ double sK[PICK_THREADS];
  __shared__ int sS[PICK_THREADS];
  long long m = (long long)*count;
  if (m > cap) m = cap;
  const long long i = (long long)blockIdx.x * PICK_THREADS + threadIdx.x;
  if ((long long)blockIdx.x * PICK_THREADS >= m) return;
  const bool live = i < m;
  const int my_slot = live ? slots[i] : 0;
  const double my_key = live ? keys[i] : 0.0;
  int w = -1;
  if (live)
    for (int k = 0; k < plan.n_want; ++k)
      if (plan.slot[k] == my_slot) w = k;
  long long below = 0, total = 0;
  for (long long base = 0; base < m; base += PICK_THREADS) {
    const long long j = base + threadIdx.x;
    sK[threadIdx.x] = j < m ? keys[j] : 0.0;
    sS[threadIdx.x] = j < m ? slots[j] : 0;
    __syncthreads();
    if (w >= 0) {
      const int lim = (int)(m - base < PICK_THREADS ? m - base : PICK_THREADS);
      for (int t = 0; t < lim; ++t) {
        const bool same = sS[t] == my_slot;
        const double k = sK[t];
        total += same;
        below += same && (k < my_key || (k == my_key && base + t < i));
      }
    }
    __syncthreads();
  }
  if (w < 0) return;
  if (below == 0) out[w] = (double)total;   // exactly one candidate per slot has rank 0
  for (int r = plan.off[w]; r < plan.off[w + 1]; ++r)
    if (plan.rank[r] == below) out[plan.n_want + r] = my_key;
}

// More candidates than ranking by counting can take: one linear-in-bits histogram per wanted
// slot narrows every wanted rank to one of PICK_BUCKETS buckets, the keys of those buckets are
// appended to short lists, and the rank is found by counting inside its list.  Four launches, all
// decisions on the device; a bucket that overflows its list (heavy ties) reports -1 as the slot
// count and the caller sorts instead.
constexpr int PICK_BUCKETS = 2048;
constexpr int PICK_LIST = 4096;
struct PickWide {
  unsigned long long lo[PICK_MAX_WANT];
  int shift[PICK_MAX_WANT];
};

__device__ __forceinline__ int pick_bucket(double key, unsigned long long lo, int shift) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(key);
  if (b <= lo) return 0;
  const unsigned long long q = (b - lo) >> shift;
  return q < (unsigned long long)PICK_BUCKETS ? (int)q : PICK_BUCKETS - 1;
}

__global__ void __launch_bounds__(PICK_THREADS)
pick_hist_kernel(const double* __restrict__ keys, const int* __restrict__ slots,
                 const unsigned long long* __restrict__ count, long long cap, PickPlan plan, PickWide wide,
                 unsigned int* __restrict__ hist) {
  long long m = (long long)*count;
  if (m > cap) m = cap;
  for (long long i = (long long)blockIdx.x * PICK_THREADS + threadIdx.x; i < m;
       i += (long long)gridDim.x * PICK_THREADS) {
    const int sl = slots[i];
    for (int k = 0; k < plan.n_want; ++k)
      if (plan.slot[k] == sl) atomicAdd(&hist[k * PICK_BUCKETS + pick_bucket(keys[i], wide.lo[k], wide.shift[k])], 1u);
  }
}

// one CTA per wanted rank r: bucket holding it and the number of keys below that bucket
__global__ void __launch_bounds__(PICK_THREADS)
pick_locate_kernel(PickPlan plan, const unsigned int* __restrict__ hist, int* __restrict__ q_bucket,
                   long long* __restrict__ q_below, double* __restrict__ out) {
  __shared__ long long sPart[PICK_THREADS];
  const int r = blockIdx.x;
  int w = 0;
  while (w + 1 < plan.n_want && plan.off[w + 1] <= r) ++w;
  const unsigned int* h = hist + (size_t)w * PICK_BUCKETS;
  constexpr int PER = PICK_BUCKETS / PICK_THREADS;
  long long mine = 0;
  for (int t = 0; t < PER; ++t) mine += h[threadIdx.x * PER + t];
  sPart[threadIdx.x] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long run = 0;
    for (int t = 0; t < PICK_THREADS; ++t) { const long long v = sPart[t]; sPart[t] = run; run += v; }
    if (r == plan.off[w]) out[w] = (double)run;
    q_bucket[r] = -1;
    q_below[r] = 0;
  }
  __syncthreads();
  long long below = sPart[threadIdx.x];
  const long long want = plan.rank[r];
  for (int t = 0; t < PER; ++t) {
    const long long c = h[threadIdx.x * PER + t];
    if (want >= below && want < below + c) { q_bucket[r] = threadIdx.x * PER + t; q_below[r] = below; }
    below += c;
  }
}

__global__ void __launch_bounds__(PICK_THREADS)
pick_collect_kernel(const double* __restrict__ keys, const int* __restrict__ slots,
                    const unsigned long long* __restrict__ count, long long cap, PickPlan plan, PickWide wide,
                    const int* __restrict__ q_bucket, double* __restrict__ lists, unsigned int* __restrict__ q_len) {
  long long m = (long long)*count;
  if (m > cap) m = cap;
  for (long long i = (long long)blockIdx.x * PICK_THREADS + threadIdx.x; i < m;
       i += (long long)gridDim.x * PICK_THREADS) {
    const int sl = slots[i];
    for (int k = 0; k < plan.n_want; ++k) {
      if (plan.slot[k] != sl) continue;
      const double key = keys[i];
      const int b = pick_bucket(key, wide.lo[k], wide.shift[k]);
      for (int r = plan.off[k]; r < plan.off[k + 1]; ++r) {
        if (q_bucket[r] != b) continue;
        const unsigned int p = atomicAdd(&q_len[r], 1u);
        if (p < (unsigned)PICK_LIST) lists[(size_t)r * PICK_LIST + p] = key;
      }
    }
  }
}

__global__ void __launch_bounds__(PICK_THREADS)
pick_list_kernel(PickPlan plan, const double* __restrict__ lists, const unsigned int* __restrict__ q_len,
                 const long long* __restrict__ q_below, const int* __restrict__ q_bucket,
                 double* __restrict__ out) {
  __shared__ double sK[PICK_LIST];
  const int r = blockIdx.x;
  int w = 0;
  while (w + 1 < plan.n_want && plan.off[w + 1] <= r) ++w;
  if (q_bucket[r] < 0) return;                      // rank beyond the slot's keys: value stays 0
  const unsigned int len = q_len[r];
  if (len > (unsigned)PICK_LIST) {                  // too many keys share the bucket: caller sorts
    if (threadIdx.x == 0) out[w] = -1.0;
    return;
  }
  for (unsigned int t = threadIdx.x; t < len; t += PICK_THREADS) sK[t] = lists[(size_t)r * PICK_LIST + t];
  __syncthreads();
  const long long want = plan.rank[r] - q_below[r];
  for (unsigned int i = threadIdx.x; i < len; i += PICK_THREADS) {
    const double mine = sK[i];
    long long below = 0;
    for (unsigned int t = 0; t < len; ++t) {
      const double k = sK[t];
      below += (k < mine) || (k == mine && t < i);
    }
    if (below == want) out[plan.n_want + r] = mine;
  }
}

}  // namespace skg

using namespace skg;

extern "C" {

int64_t skg_pair_cell_scratch_bytes(int64_t n) {
  if (n < 0) n = 0;
  return sweep_head_bytes(n);
}

int skg_pair_cell_counts(const double* coords, int64_t n, int dim, const double* dir,
                         uint64_t lo_bits, uint64_t hi_bits, int shift, int key0, uint64_t* count,
                         void* scratch, int64_t scratch_bytes, int64_t tile_begin, int64_t tile_end,
                         int64_t tile_stride, void* stream) {
  int rc = check_pair_args(coords, n, dim, tile_begin, tile_end, tile_stride);
  if (rc) return rc;
  if (!count || shift < 0 || shift > 20 || hi_bits > INF_BITS + 1) {
    set_error("count NULL, shift outside [0, 20] or hi_bits beyond +inf");
    return SKG_E_BADARG;
  }
  SweepScratch S;
  rc = carve_scratch(scratch, scratch_bytes, n, 0, &S);
  if (rc) return rc;
  DirParams dp;
  rc = build_dir_params(dir, dim, &dp);
  if (rc) return rc;
  if (tile_end == tile_begin || hi_bits <= lo_bits) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  CullWindows win;
  std::memset(&win, 0, sizeof(win));
  win.count = 1;
  std::memcpy(&win.lo[0], &lo_bits, 8);
  if (hi_bits > INF_BITS) {
    const uint64_t inf = INF_BITS;
    std::memcpy(&win.hi[0], &inf, 8);
  } else {
    std::memcpy(&win.hi[0], &hi_bits, 8);
  }
  cull_set_dir(&win, dp);
  const int64_t ntl = (tile_end - tile_begin + tile_stride - 1) / tile_stride;
  const int js = choose_js(ntl);
  rc = launch_cull(coords, n, dim, win, tile_begin, tile_end, tile_stride, js, S, st);
  if (rc) return rc;
  const int use_dir = dp.model == SKG_DIR_NONE ? 0 : (dp.amb_exclude ? 1 : 2);
  unsigned long long* cnt = reinterpret_cast<unsigned long long*>(count);
  if (dim == 3) rc = launch_cellcount<3, 0>(coords, n, dp, lo_bits, hi_bits, shift, key0, cnt, ntl, js, S, win.lo[0], win.hi[0], st);
  else if (use_dir == 1) rc = launch_cellcount<2, 1>(coords, n, dp, lo_bits, hi_bits, shift, key0, cnt, ntl, js, S, win.lo[0], win.hi[0], st);
  else if (use_dir == 2) rc = launch_cellcount<2, 2>(coords, n, dp, lo_bits, hi_bits, shift, key0, cnt, ntl, js, S, win.lo[0], win.hi[0], st);
  else rc = launch_cellcount<2, 0>(coords, n, dp, lo_bits, hi_bits, shift, key0, cnt, ntl, js, S, win.lo[0], win.hi[0], st);
  if (rc) return rc;
  SKG_CUDA(cudaGetLastError());
  return 0;
}

int64_t skg_pick_scratch_bytes(void) {
  return (int64_t)PICK_MAX_WANT * PICK_BUCKETS * 4 + (int64_t)PICK_MAX_RANKS * (PICK_LIST * 8 + 4 + 8 + 4) + 256;
}

int skg_pick_order_stats(const double* keys, const uint32_t* slots, const uint64_t* count, int64_t capacity,
                         const uint32_t* want_slot, const int32_t* want_off, const int64_t* want_rank,
                         const uint64_t* want_lo_bits, const uint64_t* want_hi_bits, int n_want,
                         double* out, void* scratch, int64_t scratch_bytes, void* stream) {
  if (!keys || !slots || !count || !want_slot || !want_off || !want_rank || !out || capacity < 0) {
    set_error("skg_pick_order_stats: NULL argument or negative capacity");
    return SKG_E_BADARG;
  }
  if (n_want < 0 || n_want > PICK_MAX_WANT || want_off[0] != 0) {
    set_error("skg_pick_order_stats: n_want outside [0, 32] or want_off[0] != 0");
    return SKG_E_BADARG;
  }
  PickPlan plan;
  std::memset(&plan, 0, sizeof(plan));
  plan.n_want = n_want;
  for (int k = 0; k < n_want; ++k) {
    if (want_off[k + 1] < want_off[k] || want_off[k + 1] > PICK_MAX_RANKS) {
      set_error("skg_pick_order_stats: want_off not ascending or more than 64 ranks");
      return SKG_E_BADARG;
    }
    plan.slot[k] = (int)want_slot[k];
    plan.off[k + 1] = want_off[k + 1];
  }
  const int n_ranks = plan.off[n_want];
  for (int r = 0; r < n_ranks; ++r) plan.rank[r] = want_rank[r];
  const bool wide = capacity > SKG_PICK_DIRECT_MAX;
  if (wide && (!want_lo_bits || !want_hi_bits || !scratch || scratch_bytes < skg_pick_scratch_bytes())) {
    set_error("skg_pick_order_stats: capacity > %d needs the key windows and skg_pick_scratch_bytes() of scratch",
              SKG_PICK_DIRECT_MAX);
    return SKG_E_BADARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int n_out = n_want + n_ranks;
  if (n_out == 0) return 0;
  SKG_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * n_out, st));
  if (capacity == 0) return 0;
  const unsigned long long* cnt = reinterpret_cast<const unsigned long long*>(count);
  const int* sl = reinterpret_cast<const int*>(slots);
  if (!wide) {
    const int64_t grid = (capacity + PICK_THREADS - 1) / PICK_THREADS;
    pick_ranks_kernel<<<(unsigned)grid, PICK_THREADS, 0, st>>>(keys, sl, cnt, (long long)capacity, plan, out);
    SKG_CUDA(cudaGetLastError());
    return 0;
  }
  if (n_ranks == 0) {
    set_error("skg_pick_order_stats: the wide path needs at least one wanted rank");
    return SKG_E_BADARG;
  }
  PickWide W;
  std::memset(&W, 0, sizeof(W));
  for (int k = 0; k < n_want; ++k) {
    const uint64_t lo = want_lo_bits[k], hi = want_hi_bits[k] > want_lo_bits[k] ? want_hi_bits[k] : want_lo_bits[k] + 1;
    int shift = 0;
    while (((hi - lo) >> shift) >= (uint64_t)PICK_BUCKETS) ++shift;
    W.lo[k] = lo;
    W.shift[k] = shift;
  }
  char* base = static_cast<char*>(scratch);
  unsigned int* hist = reinterpret_cast<unsigned int*>(base);
  size_t off = (size_t)PICK_MAX_WANT * PICK_BUCKETS * 4;
  long long* q_below = reinterpret_cast<long long*>(base + off);
  off += (size_t)PICK_MAX_RANKS * 8;
  int* q_bucket = reinterpret_cast<int*>(base + off);
  off += (size_t)PICK_MAX_RANKS * 4;
  unsigned int* q_len = reinterpret_cast<unsigned int*>(base + off);
  off += (size_t)PICK_MAX_RANKS * 4;
  off = (off + 255) & ~(size_t)255;
  double* lists = reinterpret_cast<double*>(base + off);
  SKG_CUDA(cudaMemsetAsync(base, 0, off, st));
  int64_t grid = (capacity + PICK_THREADS - 1) / PICK_THREADS;
  const int64_t max_grid = (int64_t)sm_count() * 8;
  if (grid > max_grid) grid = max_grid;
  pick_hist_kernel<<<(unsigned)grid, PICK_THREADS, 0, st>>>(keys, sl, cnt, (long long)capacity, plan, W, hist);
  pick_locate_kernel<<<n_ranks, PICK_THREADS, 0, st>>>(plan, hist, q_bucket, q_below, out);
  pick_collect_kernel<<<(unsigned)grid, PICK_THREADS, 0, st>>>(keys, sl, cnt, (long long)capacity, plan, W,
                                                               q_bucket, lists, q_len);
  pick_list_kernel<<<n_ranks, PICK_THREADS, 0, st>>>(plan, lists, q_len, q_below, q_bucket, out);
  SKG_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
