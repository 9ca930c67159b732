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

}  // extern "C"
