// K1d - gather pass for per-lag |dz| order statistics (Dowd's median, estimators.py:178, and the
// percentile estimators) with the VALUE test first.
//
// The gather pass of a windowed median looks for the few pairs whose |dz| falls into a narrow
// per-lag window [lo_g, hi_g) (one bucket of the window histogram: ~10^-5 of the class).  The
// general select kernel classifies every pair by distance first (distance, LUT look-up, window
// look-up: ~26 issue slots per pair).  Here the "j" chunk is walked in groups of G consecutive
// (Z-ordered, compact) points; for one "i" point and one group the squared distances lie in the
// point-to-box range [smin, smax] (formed with the pair arithmetic, rigorous by monotone
// rounding), so only the lag classes c(smin) .. c(smax) can occur and
//     A = min lo_c,  B = max hi_c   over those classes
// bound every window the pair could fall into.  The pair loop is then  |z_i - z_j| in [A, B) ?
// - one subtraction and two compares.  A hit (~1 %) is looked at again against the (up to three)
// class windows themselves, held in registers, and only what passes that (~10^-4) evaluates the
// distance, finds its class and tests that class's own window (out of line: one copy of the code).  Results are identical to the general kernel's
// (same keys, same window tests on the same bit patterns); the order of candidates differs.
#include <cstdlib>
#include <cstring>

#include "pair_sweep.cuh"

namespace skg {

constexpr int DG_THREADS = 128;
constexpr int DG_R = TILE / DG_THREADS;   // "i" points per thread

__device__ __forceinline__ void dg_point_box(double x, double y, double lox, double loy, double hix,
                                             double hiy, double& smin, double& smax) {
  const double ax = x - lox, bx = x - hix, ay = y - loy, by = y - hiy;
  const double ax2 = __dmul_rn(ax, ax), bx2 = __dmul_rn(bx, bx);
  const double ay2 = __dmul_rn(ay, ay), by2 = __dmul_rn(by, by);
  const bool gx = ax2 > bx2, gy = ay2 > by2;
  const double fx = gx ? ax2 : bx2, fy = gy ? ay2 : by2;
  double nx = gx ? bx2 : ax2, ny = gy ? by2 : ay2;
  if ((__double2hiint(ax) ^ __double2hiint(bx)) < 0) nx = 0.0;
  if ((__double2hiint(ay) ^ __double2hiint(by)) < 0) ny = 0.0;
  smin = __dadd_rn(nx, ny);
  smax = __dadd_rn(fx, fy);
}

struct DgHit {   // what the rare full test needs (one copy per thread, built once per kernel)
  const uint64_t* sThr; const uint8_t* lutb; int shift, klo, khi, emax, nb;
  const SegTable* seg; long long nbuckets; const uint32_t* bitmap;
  double* cand_key; uint32_t* cand_slot; unsigned long long* cand_count; long long capacity;
};

// the pair's |dz| lies in a window one of its possible classes has: now the distance decides
__device__ __noinline__ void dg_full_test(const DgHit& h, double xi, double yi, double xj, double yj, double ad) {
  const double dx = xi - xj;
  const double dy = yi - yj;
  const double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
  const int c = lag_of<0>((uint64_t)__double_as_longlong(s), h.sThr, h.lutb, h.shift, h.klo, h.khi, h.emax);
  if (c >= h.nb) return;
  const uint64_t kb = (uint64_t)__double_as_longlong(ad);
  if (kb < h.seg->lo_bits[c] || kb >= h.seg->hi_bits[c]) return;
  const long long slot = (long long)c * h.nbuckets;   // scale 0: the window is bucket 0
  if (!((h.bitmap[slot >> 5] >> (slot & 31)) & 1u)) return;
  const unsigned long long pos = atomicAdd(h.cand_count, 1ull);
  if ((long long)pos < h.capacity) {
    h.cand_key[pos] = ad;
    h.cand_slot[pos] = (uint32_t)slot;
  }
}

template <int EMAX>
__global__ void __launch_bounds__(DG_THREADS, 6)
pair_dz_gather_kernel(const double* __restrict__ coords, const double* __restrict__ values, int64_t n,
                      const __grid_constant__ BinTable tab, const __grid_constant__ SegTable seg_tab,
                      long long nbuckets, const uint32_t* __restrict__ bitmap, double* __restrict__ cand_key,
                      uint32_t* __restrict__ cand_slot, unsigned long long* __restrict__ cand_count,
                      long long capacity, const unsigned int* __restrict__ tile_list,
                      const unsigned long long* __restrict__ tile_count, const double* __restrict__ bbox32,
                      const double* __restrict__ bbox8, double wlo, double whi, int js, int G) {
  __shared__ double2 sXY[TILE];
  __shared__ double sZ[TILE];
  __shared__ double sThrD[THR_MAX + 4];      // thresholds as doubles, +inf beyond the last class
  __shared__ uint8_t sLut[LUT_MAX];
  __shared__ double sLo[THR_MAX + 4];        // class windows as doubles; no window: [+inf, -inf)
  __shared__ double sHi[THR_MAX + 4];
  __shared__ double2 sHull2[THR_MAX + 2];    // (min lo, max hi) over the classes c, c+1
  __shared__ double2 sHull3[THR_MAX + 2];    // ... over c, c+1, c+2
  __shared__ double sGB[TILE / 8][4];
  __shared__ double sAll[2];                 // bounds over all classes
  const int tid = threadIdx.x;
  const int nb = tab.nb;
  const double qnan = __longlong_as_double(NAN_BITS);
  const double inf = __longlong_as_double((long long)INF_BITS);
  for (int k = tid; k < THR_MAX + 4; k += DG_THREADS) {
    sThrD[k] = k < nb ? __longlong_as_double((long long)tab.thr[k]) : inf;
    const bool has = k < nb && seg_tab.hi_bits[k] > seg_tab.lo_bits[k];
    sLo[k] = has ? __longlong_as_double((long long)seg_tab.lo_bits[k]) : inf;
    sHi[k] = has ? __longlong_as_double((long long)seg_tab.hi_bits[k]) : -inf;
  }
  for (int k = tid; k < tab.ncell; k += DG_THREADS) sLut[k] = tab.lut[k];
  __syncthreads();
  for (int k = tid; k < THR_MAX + 2; k += DG_THREADS) {
    sHull2[k] = make_double2(fmin(sLo[k], sLo[k + 1]), fmax(sHi[k], sHi[k + 1]));
    sHull3[k] = make_double2(fmin(fmin(sLo[k], sLo[k + 1]), sLo[k + 2]), fmax(fmax(sHi[k], sHi[k + 1]), sHi[k + 2]));
  }
  if (tid == 0) {
    double a = inf, b = -inf;
    for (int k = 0; k < nb; ++k) { a = fmin(a, sLo[k]); b = fmax(b, sHi[k]); }
    sAll[0] = a;
    sAll[1] = b;
  }
  __syncthreads();
  const double allA = sAll[0], allB = sAll[1];
  const int shift = tab.shift, klo = tab.key0, khi = tab.key0 + tab.ncell - 1, emax = tab.emax;
  const uint8_t* lutb = sLut - klo;
  const uint64_t* sThr = reinterpret_cast<const uint64_t*>(sThrD);
  const DgHit hit{sThr, lutb, shift, klo, khi, emax, nb, &seg_tab, nbuckets, bitmap,
                  cand_key, cand_slot, cand_count, capacity};
  const int64_t nt = (n + TILE - 1) / TILE;
  const double* __restrict__ X = coords;
  const double* __restrict__ Y = coords + n;
  const int jchunk = TILE / js;
  if (G > jchunk) G = jchunk;
  const int gshift = G == 32 ? 5 : (G == 16 ? 4 : 3);
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
      chunk_box<2>(bbox32, J, jbeg / jchunk, js, jb);
#pragma unroll
      for (int r = 0; r < DG_R; ++r) {
        double a, b;
        const int64_t sub = (int64_t)I * 8 + r * (DG_THREADS / 32) + (tid >> 5);
        subblock_range<2>(bbox32, sub, jb, a, b);
        wact |= (a * (1.0 - 1e-12) < whi && b * (1.0 + 1e-12) >= wlo);
      }
    }
    const int64_t jn = n - ((int64_t)J * TILE + jbeg);
    const int jcount = jn < jchunk ? (jn > 0 ? (int)jn : 0) : jchunk;
    double xi[DG_R], yi[DG_R], zi[DG_R];
#pragma unroll
    for (int r = 0; r < DG_R; ++r) {
      const int64_t i = (int64_t)I * TILE + r * DG_THREADS + tid;
      const bool v = i < n;
      xi[r] = v ? X[i] : qnan;
      yi[r] = v ? Y[i] : qnan;
      zi[r] = v ? values[i] : qnan;   // padded point: |dz| = NaN fails every window test
    }
    __syncthreads();
    for (int k = tid; k < jcount; k += DG_THREADS) {
      const int64_t j = (int64_t)J * TILE + jbeg + k;
      sXY[k] = make_double2(X[j], Y[j]);
      sZ[k] = values[j];
    }
    {
      const int per = G >> 3;
      for (int g = tid; g < (jchunk >> gshift); g += DG_THREADS) {
        const int64_t b8 = ((int64_t)J * TILE + jbeg) / 8 + (int64_t)g * per;
        double lx = bbox8[b8 * 4], ly = bbox8[b8 * 4 + 1], hx = bbox8[b8 * 4 + 2], hy = bbox8[b8 * 4 + 3];
        bool bad = !(lx == lx);   // flagged 8-box: NaN coordinate or ragged end
        for (int u = 1; u < per; ++u) {
          const double l2 = bbox8[(b8 + u) * 4];
          bad |= !(l2 == l2);
          lx = fmin(lx, l2);
          ly = fmin(ly, bbox8[(b8 + u) * 4 + 1]);
          hx = fmax(hx, bbox8[(b8 + u) * 4 + 2]);
          hy = fmax(hy, bbox8[(b8 + u) * 4 + 3]);
        }
        if (bad) lx = ly = hx = hy = qnan;
        sGB[g][0] = lx; sGB[g][1] = ly; sGB[g][2] = hx; sGB[g][3] = hy;
      }
    }
    __syncthreads();
    if (jcount == 0 || !wact) continue;
    const bool diag = I == J;

    for (int g0 = 0; g0 < jcount; g0 += G) {
      const int gi = g0 >> gshift;
      const double2 blo = *reinterpret_cast<const double2*>(&sGB[gi][0]);
      const double2 bhi = *reinterpret_cast<const double2*>(&sGB[gi][2]);
      // per point: the hull [A, B) of the windows of the classes its pairs with this group can
      // fall into, and (up to three of) those windows themselves for the second look
      double A[DG_R], B[DG_R], w0l[DG_R], w0h[DG_R], w1l[DG_R], w1h[DG_R], w2l[DG_R], w2h[DG_R];
      bool idle = true;
#pragma unroll
      for (int r = 0; r < DG_R; ++r) {
        double smin, smax;
        dg_point_box(xi[r], yi[r], blo.x, blo.y, bhi.x, bhi.y, smin, smax);
        const int c0 = lag_of<EMAX>((uint64_t)__double_as_longlong(smin), sThr, lutb, shift, klo, khi, emax);
        const int cb = min(c0, nb);
        // smax < thr[cb + k]: classes cb .. cb+k only (NaN - flagged box, padded point - fails all)
        const bool fit1 = smax < sThrD[cb], fit2 = smax < sThrD[cb + 1], fit3 = smax < sThrD[cb + 2];
        const double2 h2 = sHull2[cb], h3 = sHull3[cb];
        w0l[r] = sLo[cb]; w0h[r] = sHi[cb];
        w1l[r] = fit1 ? inf : sLo[cb + 1]; w1h[r] = fit1 ? -inf : sHi[cb + 1];
        w2l[r] = fit2 ? inf : sLo[cb + 2]; w2h[r] = fit2 ? -inf : sHi[cb + 2];
        double a = fit1 ? w0l[r] : (fit2 ? h2.x : h3.x);
        double b = fit1 ? w0h[r] : (fit2 ? h2.y : h3.y);
        if (!fit3) {          // wide or flagged group (rare): hull over its classes, hits take the full test
          a = allA; b = allB;
          if (smax == smax && smin == smin) {
            int c1 = lag_of<0>((uint64_t)__double_as_longlong(smax), sThr, lutb, shift, klo, khi, emax);
            c1 = c1 < nb ? c1 : nb;
            a = inf; b = -inf;
            for (int c = cb; c <= c1; ++c) { a = fmin(a, sLo[c]); b = fmax(b, sHi[c]); }
          }
          w0l[r] = a; w0h[r] = b;
        }
        A[r] = a;
        B[r] = b;
        idle &= !(a < b);
      }
      if (__all_sync(0xffffffffu, idle)) continue;   // no class of this group has a window
      const int g1 = min(g0 + G, jcount);
      auto look = [&](int r, int jj, double ad) {   // inside the hull: the class windows themselves
        const bool in = (ad >= w0l[r] && ad < w0h[r]) || (ad >= w1l[r] && ad < w1h[r]) ||
                        (ad >= w2l[r] && ad < w2h[r]);
        if (in && (!diag || (jbeg + jj > r * DG_THREADS + tid))) {   // ~10^-4: keep i < j only
          const double2 pj = sXY[jj];
          dg_full_test(hit, xi[r], yi[r], pj.x, pj.y, ad);
        }
      };
      int jj = g0;
      for (; jj + 1 < g1; jj += 2) {   // two j points x DG_R i points per branch
        const double za = sZ[jj], zb = sZ[jj + 1];
        double ada[DG_R], adb[DG_R];
        bool ha[DG_R], hb[DG_R], any = false;
#pragma unroll
        for (int r = 0; r < DG_R; ++r) {
          ada[r] = fabs(zi[r] - za);
          adb[r] = fabs(zi[r] - zb);
          ha[r] = ada[r] >= A[r] && ada[r] < B[r];
          hb[r] = adb[r] >= A[r] && adb[r] < B[r];
          any |= ha[r] | hb[r];
        }
        if (any) {
#pragma unroll
          for (int r = 0; r < DG_R; ++r) {
            if (ha[r]) look(r, jj, ada[r]);
            if (hb[r]) look(r, jj + 1, adb[r]);
          }
        }
      }
      if (jj < g1) {
        const double za = sZ[jj];
#pragma unroll
        for (int r = 0; r < DG_R; ++r) {
          const double ad = fabs(zi[r] - za);
          if (ad >= A[r] && ad < B[r]) look(r, jj, ad);
        }
      }
    }
  }
}

// Takes the gather when the select's windows are per-lag |dz| windows of the kind exact_select's
// gather stage builds (2-D, isotropic, scalar values, every segment scale 0, finite windows).
bool dz_gather_applicable(int dim, int use_dir, bool seg, int key, bool use_mv, const SegTable& sg, int nseg) {
  if (dim != 2 || use_dir != 0 || !seg || key != SKG_KEY_ABSDZ || use_mv) return false;
  if (getenv("SKG_DZ_GATHER_OFF")) return false;
  for (int k = 0; k < nseg; ++k) {
    if (sg.hi_bits[k] <= sg.lo_bits[k]) continue;
    if (sg.scale[k] != 0.0 || sg.hi_bits[k] > INF_BITS) return false;
  }
  return true;
}

template <int EMAX>
static int launch_dz_gather_t(const SelectLaunch& L) {
  auto kern = pair_dz_gather_kernel<EMAX>;
  static thread_local int occ = 0;
  if (occ == 0) {
    SKG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, DG_THREADS, 0));
    occ = occ < 1 ? 1 : (occ < MAX_GRID_PER_SM ? occ : MAX_GRID_PER_SM);
  }
  int64_t grid = (int64_t)sm_count() * occ;
  if (grid > L.ntiles * L.js) grid = L.ntiles * L.js;
  if (grid < 1) grid = 1;
  const int jchunk = TILE / L.js;
  const int G = jchunk >= 32 ? 32 : 16;
  kern<<<(unsigned)grid, DG_THREADS, 0, L.st>>>(
      L.coords, L.values, L.n, *L.tab, *L.seg, (long long)L.nbuckets, L.bitmap, L.cand_key, L.cand_slot,
      reinterpret_cast<unsigned long long*>(L.cand_count), (long long)L.capacity, L.list, L.list_count,
      L.bbox32, L.bbox8, L.wlo, L.whi, L.js, G);
  SKG_CUDA(cudaGetLastError());
  return 0;
}

int launch_dz_gather(const SelectLaunch& L) {
  return L.tab->emax <= 1 ? launch_dz_gather_t<1>(L) : launch_dz_gather_t<0>(L);
}

}  // namespace skg
