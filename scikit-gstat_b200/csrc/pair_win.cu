// K2w - the fused pass with REGISTER lag windows (2-D, isotropic, scalar values: the
// configurations BASELINE.json quotes its variogram metric on).
//
// The per-pair read-modify-write of a private shared-memory accumulator is what bounded the
// first fused kernel (6 of ~7.7 shared-memory wavefronts per warp step, FP64 pipe 36 %).  Here
// the "j" chunk is walked in GROUPS of G = 8/16/32 consecutive (Z-order, hence compact) points.
// For one "i" point and one group the squared distances of all G pairs lie in
// [smin, smax], the point-to-box range, which is formed with the pair arithmetic itself
// (same operand order, no FMA) so that monotone rounding makes it a rigorous bound.  If that
// range spans at most W = 2 (or 3) lag classes, the thread keeps W running sums and counts in
// REGISTERS, classifies each pair with W-1 compares against thresholds held in registers, and
// touches shared memory once per group instead of once per pair.  Windows are per THREAD and
// per "i" point (a lane's window only has to cover the group's box as seen from its own
// point), the choice fast/per-pair is per warp and group (vote), so there is no divergence.
// Groups that do not fit (wide boxes, diagonal tiles, ragged ends, NaN coordinates) take the
// per-pair path of the first kernel, unchanged.  Results are identical in counts; sums differ
// from the first kernel only in summation order.
#include <cstdlib>
#include <cstring>

#include "pair_sweep.cuh"

namespace skg {

// ------------------------------------------------------------ group size ---
//
// Mean diagonal of the boxes of 8/16/32 consecutive points against the mean lag-class width:
// the largest group whose boxes usually fit a 2- or 3-class window wins (0 = windows off).
__global__ void __launch_bounds__(256)
win_plan_kernel(const double* __restrict__ bbox8, int64_t ngroups8, double last_edge_sq, int nb,
                int force_g, unsigned long long* __restrict__ counters) {
  __shared__ double s_sum[3][8];
  __shared__ int s_cnt[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t nquad = ngroups8 / 4;  // aligned runs of 32 points
  const int64_t step = nquad > 8192 ? nquad / 8192 : 1;
  double d8 = 0.0, d16 = 0.0, d32 = 0.0;
  int cnt = 0;
  for (int64_t q = (int64_t)tid * step; q < nquad; q += 256 * step) {
    double lx[4], ly[4], hx[4], hy[4];
    bool ok = true;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      lx[u] = bbox8[(q * 4 + u) * 4 + 0]; ly[u] = bbox8[(q * 4 + u) * 4 + 1];
      hx[u] = bbox8[(q * 4 + u) * 4 + 2]; hy[u] = bbox8[(q * 4 + u) * 4 + 3];
      ok &= (lx[u] == lx[u]);
    }
    if (!ok) continue;
    auto diag = [](double ax, double ay, double bx, double by) {
      return sqrt((bx - ax) * (bx - ax) + (by - ay) * (by - ay));
    };
#pragma unroll
    for (int u = 0; u < 4; ++u) d8 += 0.25 * diag(lx[u], ly[u], hx[u], hy[u]);
#pragma unroll
    for (int u = 0; u < 4; u += 2)
      d16 += 0.5 * diag(fmin(lx[u], lx[u + 1]), fmin(ly[u], ly[u + 1]), fmax(hx[u], hx[u + 1]),
                        fmax(hy[u], hy[u + 1]));
    d32 += diag(fmin(fmin(lx[0], lx[1]), fmin(lx[2], lx[3])), fmin(fmin(ly[0], ly[1]), fmin(ly[2], ly[3])),
                fmax(fmax(hx[0], hx[1]), fmax(hx[2], hx[3])), fmax(fmax(hy[0], hy[1]), fmax(hy[2], hy[3])));
    ++cnt;
  }
  d8 = warp_sum(d8); d16 = warp_sum(d16); d32 = warp_sum(d32); cnt = warp_sum(cnt);
  if (lane == 0) { s_sum[0][warp] = d8; s_sum[1][warp] = d16; s_sum[2][warp] = d32; s_cnt[warp] = cnt; }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0, b = 0.0, c = 0.0;
    int m = 0;
    for (int w = 0; w < 8; ++w) { a += s_sum[0][w]; b += s_sum[1][w]; c += s_sum[2][w]; m += s_cnt[w]; }
    int G = 0;
    if (force_g >= 0) {
      G = force_g;
    } else if (m > 0 && nb >= 2 && last_edge_sq > 0.0 && last_edge_sq < 1e300) {
      const double width = sqrt(last_edge_sq) / nb;  // mean lag-class width
      a /= m; b /= m; c /= m;
      // measured (B200, uniform points, 25 even lags up to the median): 32-point groups win up to a
      // box-diagonal / class-width ratio of ~1.6, 16-point groups up to ~2.2 (8-point groups never
      // pay for their set-up); beyond that the per-pair path is faster
      if (c <= 1.6 * width) G = 32;
      else if (b <= 2.2 * width) G = 16;
    }
    counters[4] = (unsigned long long)G;
    // for inspection (tools/win_bench.py): mean box diagonals of 16 / 32 points, mean class width
    counters[5] = (unsigned long long)__double_as_longlong(m > 0 ? (force_g >= 0 ? b / m : b) : 0.0);
    counters[6] = (unsigned long long)__double_as_longlong(m > 0 ? (force_g >= 0 ? c / m : c) : 0.0);
    counters[7] = (unsigned long long)__double_as_longlong(nb > 0 ? sqrt(last_edge_sq) / nb : 0.0);
  }
}

// ---------------------------------------------------------------- kernel ---

#ifndef SKG_WIN_THREADS
#define SKG_WIN_THREADS 128
#endif
constexpr int WB = SKG_WIN_THREADS;  // threads per CTA
constexpr int WR = TILE / WB;        // "i" points per thread
constexpr int WIN_MIN_CTAS = WB == 64 ? 6 : 4;  // register cap: 168 / 128 per thread
constexpr int WIN_COUNT_CTAS = WB == 64 ? 8 : 6; // count-only sweeps (STATS == 0) need fewer registers: 80 per thread
#ifndef SKG_WIN_UNROLL2
#define SKG_WIN_UNROLL2 (SKG_WIN_THREADS == 64 ? 2 : 4)
#endif
constexpr int WIN_UNROLL2 = SKG_WIN_UNROLL2;  // j points in flight per thread in the 2-class loop


template <int STATS>
__device__ __forceinline__ void flush_row(unsigned saddr, unsigned caddr, double sq, double rt, unsigned cnt) {
  if (STATS == 0) {
    asm volatile(
        "{\n\t.reg .b32 c;\n\t"
        "ld.shared.u32 c, [%0];\n\t"
        "add.u32 c, c, %1;\n\t"
        "st.shared.u32 [%0], c;\n\t}"
        :: "r"(caddr), "r"(cnt));
  } else if (STATS == 4) {  // sq carries the below-window count of the class (exact: < 2^32)
    const unsigned below = (unsigned)__double2int_rn(sq);
    asm volatile(
        "{\n\t.reg .b32 b, c;\n\t"
        "ld.shared.u32 b, [%0];\n\t"
        "ld.shared.u32 c, [%1];\n\t"
        "add.u32 b, b, %2;\n\tadd.u32 c, c, %3;\n\t"
        "st.shared.u32 [%0], b;\n\t"
        "st.shared.u32 [%1], c;\n\t}"
        :: "r"(saddr), "r"(caddr), "r"(below), "r"(cnt));
  } else if (STATS == 3) {
    asm volatile(
        "{\n\t.reg .b64 tb, rb;\n\t.reg .f64 t, u;\n\t.reg .b32 c;\n\t"
        "ld.shared.v2.b64 {tb, rb}, [%0];\n\t"
        "ld.shared.u32 c, [%1];\n\t"
        "mov.b64 t, tb;\n\tmov.b64 u, rb;\n\t"
        "add.rn.f64 t, t, %2;\n\t"
        "add.rn.f64 u, u, %3;\n\t"
        "add.u32 c, c, %4;\n\t"
        "mov.b64 tb, t;\n\tmov.b64 rb, u;\n\t"
        "st.shared.v2.b64 [%0], {tb, rb};\n\t"
        "st.shared.u32 [%1], c;\n\t}"
        :: "r"(saddr), "r"(caddr), "d"(sq), "d"(rt), "r"(cnt));
  } else {
    const double v = (STATS == 2) ? rt : sq;
    asm volatile(
        "{\n\t.reg .b64 tb;\n\t.reg .f64 t;\n\t.reg .b32 c;\n\t"
        "ld.shared.b64 tb, [%0];\n\t"
        "ld.shared.u32 c, [%1];\n\t"
        "mov.b64 t, tb;\n\t"
        "add.rn.f64 t, t, %2;\n\t"
        "add.u32 c, c, %3;\n\t"
        "mov.b64 tb, t;\n\t"
        "st.shared.b64 [%0], tb;\n\t"
        "st.shared.u32 [%1], c;\n\t}"
        :: "r"(saddr), "r"(caddr), "d"(v), "r"(cnt));
  }
}

// squared-distance range between a point and a box, with the pair arithmetic (monotone
// rounding: smin <= s <= smax for every point of the box, see the header of this file)
__device__ __forceinline__ void point_box_range(double x, double y, double lox, double loy,
                                                double hix, double hiy, double& smin, double& smax) {
  const double ax = x - lox, bx = x - hix, ay = y - loy, by = y - hiy;
  const double ax2 = __dmul_rn(ax, ax), bx2 = __dmul_rn(bx, bx);
  const double ay2 = __dmul_rn(ay, ay), by2 = __dmul_rn(by, by);
  const bool gx = ax2 > bx2, gy = ay2 > by2;
  const double fx = gx ? ax2 : bx2, fy = gy ? ay2 : by2;
  double nx = gx ? bx2 : ax2, ny = gy ? by2 : ay2;
  // lo < x < hi (differences of opposite sign): the nearest point of the box has the same x
  if ((__double2hiint(ax) ^ __double2hiint(bx)) < 0) nx = 0.0;
  if ((__double2hiint(ay) ^ __double2hiint(by)) < 0) ny = 0.0;
  smin = __dadd_rn(nx, ny);
  smax = __dadd_rn(fx, fy);
}

template <int STATS, int EMAX>
__global__ void __launch_bounds__(WB, STATS == 0 ? WIN_COUNT_CTAS : WIN_MIN_CTAS)
pair_hist_win_kernel(const double* __restrict__ coords, const double* __restrict__ values, int64_t n,
                     const __grid_constant__ BinTable tab, HistPartial* __restrict__ partial,
                     const unsigned int* __restrict__ tile_list,
                     const unsigned long long* __restrict__ counters,
                     const double* __restrict__ bbox32, const double* __restrict__ bbox8,
                     double wlo, double whi, int js, const DzTable* __restrict__ dzt,
                     unsigned long long* __restrict__ whist, long long wbuckets) {
  // STATS == 4 (Dowd's windowed median, skg_pair_dz_windows): per class the pair count, the count of
  // |dz| below the class's window [lo, hi) and - rare - an atomic into the window's bucket histogram
  __shared__ double2 sXY[TILE + 2];
  __shared__ double sZ[TILE + 2];
  __shared__ double sDzLo[STATS == 4 ? THR_MAX + 2 : 1];
  __shared__ double sDzHi[STATS == 4 ? THR_MAX + 2 : 1];
  __shared__ double sDzSc[STATS == 4 ? THR_MAX + 2 : 1];
  __shared__ double sThrD[THR_MAX + 2];  // thresholds as doubles, +inf beyond the last class
  __shared__ uint16_t sLut[LUT_MAX];     // lut[cell] * 8
  __shared__ double sGB[TILE / 8][4];    // boxes of the current chunk's j groups
  extern __shared__ __align__(16) unsigned char acc_raw[];

  using PL = PrivLayout<STATS>;
  constexpr int SUMB = PL::SUMB;
  const int nb = tab.nb;
  const int NR = nb + 3;  // rows nb .. nb+2 collect what lies beyond the last edge (never read)
  const int tid = threadIdx.x;
  const double qnan = __longlong_as_double(NAN_BITS);
  const double inf = __longlong_as_double((long long)INF_BITS);

  for (int k = tid; k < THR_MAX + 2; k += WB)
    sThrD[k] = k < nb ? __longlong_as_double((long long)tab.thr[k]) : inf;
  for (int k = tid; k < tab.ncell; k += WB) sLut[k] = (uint16_t)(tab.lut[k] * 8);
  if (STATS == 4) {
    // classes beyond the last edge repeat the last class's window, so that a group straddling the
    // last edge still sees ONE window (shared-window loop below); their rows are never read and
    // their in-window pairs are kept out of the histogram by the lg < nb tests
    for (int k = tid; k < THR_MAX + 2; k += WB) {
      const int kk = k < nb ? k : nb - 1;
      sDzLo[k] = dzt->lo[kk];
      sDzHi[k] = dzt->hi[kk];
      sDzSc[k] = dzt->scale[kk];
    }
  }
  {
    unsigned long long* z = reinterpret_cast<unsigned long long*>(acc_raw);
    for (int k = tid; k < NR * (WB * (SUMB + 4) / 8); k += WB) z[k] = 0ull;
  }
  __syncthreads();

  const int shift = tab.shift, klo = tab.key0, khi = tab.key0 + tab.ncell - 1, emax = tab.emax;
  const unsigned char* lutb = reinterpret_cast<const unsigned char*>(sLut) - 2 * klo;
  const uint64_t* sThr = reinterpret_cast<const uint64_t*>(sThrD);
  const char* thrBytes = reinterpret_cast<const char*>(sThrD);
  constexpr int GS = 8;
  const unsigned nbS = (unsigned)nb * GS;
  const double last_thr = nb > 0 ? sThrD[nb - 1] : 0.0;
  const unsigned aBase = (unsigned)__cvta_generic_to_shared(acc_raw);
  const unsigned aSum = aBase + tid * SUMB;
  const unsigned aCnt = aBase + (unsigned)NR * (WB * SUMB) + tid * 4;
  constexpr unsigned SROW = SUMB * WB / 8, CROW = WB * 4 / 8;  // row strides per unit of g8
  const int64_t nt = (n + TILE - 1) / TILE;
  const double* __restrict__ X = coords;
  const double* __restrict__ Y = coords + n;
  const int jchunk = TILE / js;  // >= 16
  int G = (int)counters[4];
  if (G > jchunk) G = jchunk;  // a group never straddles two work items
  const int gshift = G == 32 ? 5 : (G == 16 ? 4 : 3);

#ifdef SKG_WIN_DEBUG
  unsigned long long dbg[6] = {0, 0, 0, 0, 0, 0};  // warp groups: dead, W=2, W=3, per-pair, diagonal, culled items
#define SKG_DBG(k) do { if ((tid & 31) == 0) dbg[k]++; } while (0)
#else
#define SKG_DBG(k) do { } while (0)
#endif
  const bool hist_on = STATS == 4 && whist != nullptr;
  const int64_t nwork = (int64_t)counters[0];
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
      for (int r = 0; r < WR; ++r) {
        double a, b;
        const int64_t sub = (int64_t)I * 8 + r * (WB / 32) + (tid >> 5);
        subblock_range<2>(bbox32, sub, jb, a, b);
        wact |= (a * (1.0 - 1e-12) < whi && b * (1.0 + 1e-12) >= wlo);
      }
    }
    const int64_t jn = n - ((int64_t)J * TILE + jbeg);
    const int jcount = jn < jchunk ? (jn > 0 ? (int)jn : 0) : jchunk;
    double xi[WR], yi[WR], zi[WR];
#pragma unroll
    for (int r = 0; r < WR; ++r) {
      const int64_t i = (int64_t)I * TILE + r * WB + tid;
      const bool v = i < n;
      xi[r] = v ? X[i] : qnan;
      yi[r] = v ? Y[i] : qnan;
      zi[r] = (STATS != 0 && v) ? values[i] : 0.0;
    }
    __syncthreads();  // every read of the previous j chunk is done
    for (int k = tid; k < jcount + 2; k += WB) {  // two NaN entries pad the pipeline tail
      const int64_t j = (int64_t)J * TILE + jbeg + k;
      const bool v = k < jcount;
      sXY[k] = make_double2(v ? X[j] : qnan, v ? Y[j] : qnan);
      sZ[k] = (STATS != 0 && v) ? values[j] : 0.0;
    }
    if (G) {
      const int per = G >> 3;  // 8-point boxes per group
      for (int g = tid; g < (jchunk >> gshift); g += WB) {
        const int64_t b8 = ((int64_t)J * TILE + jbeg) / 8 + (int64_t)g * per;
        double lx = bbox8[b8 * 4], ly = bbox8[b8 * 4 + 1], hx = bbox8[b8 * 4 + 2], hy = bbox8[b8 * 4 + 3];
        bool bad = !(lx == lx);  // a flagged 8-box (NaN / missing point) flags the whole group
        for (int u = 1; u < per; ++u) {
          const double l2 = bbox8[(b8 + u) * 4];
          bad |= !(l2 == l2);
          lx = fmin(lx, l2);
          ly = fmin(ly, bbox8[(b8 + u) * 4 + 1]);
          hx = fmax(hx, bbox8[(b8 + u) * 4 + 2]);
          hy = fmax(hy, bbox8[(b8 + u) * 4 + 3]);
        }
        if (bad) lx = ly = hx = hy = qnan;  // every window test on the group fails
        sGB[g][0] = lx; sGB[g][1] = ly; sGB[g][2] = hx; sGB[g][3] = hy;
      }
    }
    __syncthreads();
    if (jcount == 0 || !wact) { SKG_DBG(5); continue; }  // warp-uniform; both barriers are already passed

    // ---- per-pair path (first kernel): pairs j0 .. j1-1 of the chunk, software-pipelined ----
    auto per_pair = [&](int j0, int j1, auto diag_tag) {
      constexpr bool DIAG = decltype(diag_tag)::value;
      auto front = [&](int jj, unsigned (&g)[WR], double (&dz)[WR]) {
        const double2 pj = sXY[jj];
        const double zj = sZ[jj];
#pragma unroll
        for (int r = 0; r < WR; ++r) {
          const double dx = xi[r] - pj.x;
          const double dy = yi[r] - pj.y;
          const double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
          const uint64_t sb = (uint64_t)__double_as_longlong(s);
          unsigned gg = lag_off8<EMAX>(sb, sThr, lutb, shift, klo, khi, emax);
          if (DIAG) gg = (jbeg + jj <= r * WB + tid) ? nbS : gg;  // keep i < j only
          g[r] = gg;
          dz[r] = zi[r] - zj;
          if (STATS == 4) {
            // flag for rmw_split<4> (+1: below the window) with the bucket of an in-window key
            // folded in as -(slot + 1): the atomic itself waits for accumulate(), because the
            // pipeline's look-ahead entry is evaluated here but never accumulated
            const double ad = fabs(dz[r]);
            const unsigned lg = min(gg, nbS) / GS;
            const double lo = sDzLo[lg];
            double flag = (ad < lo) ? 1.0 : 0.0;
            if (whist && gg < nbS && ad >= lo && ad < sDzHi[lg]) {
              long long bk = __double2ll_rz(__dmul_rn(__dsub_rn(ad, lo), sDzSc[lg]));
              bk = bk < wbuckets - 1 ? bk : wbuckets - 1;
              flag = -(double)((long long)lg * wbuckets + bk + 1);
            }
            dz[r] = flag;
          }
        }
      };
      auto accumulate = [&](const unsigned (&g)[WR], const double (&dz)[WR]) {
#pragma unroll
        for (int r = 0; r < WR; ++r) {
          double v = dz[r];
          if (STATS == 4 && v < 0.0) {  // in-window key: its bucket (exact in a double: < 2^32 slots)
            atomicAdd(&whist[(long long)(-v) - 1], 1ull);
            v = 0.0;
          }
          rmw_split<STATS>(aSum + g[r] * SROW, aCnt + g[r] * CROW, v, g[r] < nbS ? 1 : 0);
        }
      };
      unsigned gA[WR], gB[WR];
      double dA[WR], dB[WR];
      front(j0, gA, dA);
      int jj = j0 + 1;
      for (; jj + 1 <= j1; jj += 2) {  // the entry at j1 is a later point or NaN padding: never accumulated
        front(jj, gB, dB);
        accumulate(gA, dA);
        front(jj + 1, gA, dA);
        accumulate(gB, dB);
      }
      if (jj <= j1) {
        front(jj, gB, dB);
        accumulate(gA, dA);
      }
    };

    if (I == J) { SKG_DBG(4); per_pair(0, jcount, std::true_type{}); continue; }

    // ---- register windows ----
    const int Gs = G ? G : jchunk;  // windows off: the chunk is one per-pair "group"
    for (int g0 = 0; g0 < jcount; g0 += Gs) {
      unsigned gb[WR];
      double T1[WR], T2[WR];
      bool all1 = false, all2 = false, all3 = false;
      if (G) {
        const int gi = g0 >> gshift;
        const double2 blo = *reinterpret_cast<const double2*>(&sGB[gi][0]);
        const double2 bhi = *reinterpret_cast<const double2*>(&sGB[gi][2]);
        bool fit1 = true, fit2 = true, fit3 = true, dead = true;
#pragma unroll
        for (int r = 0; r < WR; ++r) {
          double smin, smax;
          point_box_range(xi[r], yi[r], blo.x, blo.y, bhi.x, bhi.y, smin, smax);
          const unsigned g8 = lag_off8<EMAX>((uint64_t)__double_as_longlong(smin), sThr, lutb, shift, klo, khi, emax);
          gb[r] = g8;
          T1[r] = *reinterpret_cast<const double*>(thrBytes + g8);
          T2[r] = *reinterpret_cast<const double*>(thrBytes + g8 + 8);
          const double T3 = *reinterpret_cast<const double*>(thrBytes + g8 + 16);
          fit1 &= smax < T1[r];   // the whole group in the point's own class
          fit2 &= smax < T2[r];   // NaN (flagged box, padded point): no window fits
          fit3 &= smax < T3;
          dead &= smin >= last_thr;
        }
        if (__all_sync(0xffffffffu, dead)) { SKG_DBG(0); continue; }  // the whole group lies beyond the last edge
        all1 = __all_sync(0xffffffffu, fit1);
        all2 = all1 || __all_sync(0xffffffffu, fit2);
        all3 = all2 || __all_sync(0xffffffffu, fit3);
      }
      if (!all3) {
        SKG_DBG(3);
        per_pair(g0, min(g0 + Gs, jcount), std::false_type{});
        continue;
      }
      SKG_DBG(all2 ? 1 : 2);
      if (STATS == 4) {
        // |dz| windows of the (up to three) classes of each point's lag window, in registers
        double wl[WR][3], wh[WR][3];
        unsigned cb[WR][3], c1[WR], c2[WR];
#pragma unroll
        for (int r = 0; r < WR; ++r) {
          const unsigned b = gb[r] / GS;
#pragma unroll
          for (int k = 0; k < 3; ++k) { wl[r][k] = sDzLo[b + k]; wh[r][k] = sDzHi[b + k]; cb[r][k] = 0u; }
          c1[r] = c2[r] = 0u;
        }
        // The classes of the window share ONE |dz| window (lag classes past the range of the field
        // have the same median; the host merges their windows): "below" and "inside" no longer
        // depend on the class, only their attribution does - packed cumulative counters
        // (pairs | below << 16, at most G = 32 each) under the class predicates, no selects.
        bool same = true;
#pragma unroll
        for (int r = 0; r < WR; ++r) {
          const bool s01 = wl[r][0] == wl[r][1] && wh[r][0] == wh[r][1];
          const bool s12 = wl[r][1] == wl[r][2] && wh[r][1] == wh[r][2];
          same &= all1 || (s01 && (all2 || s12));
        }
        if (__all_sync(0xffffffffu, same)) {
          unsigned a0[WR], a1[WR], a2[WR];
          double lo[WR], hi[WR], sc[WR], T2x[WR];
          unsigned long long* hrow[WR];
#pragma unroll
          for (int r = 0; r < WR; ++r) {
            a0[r] = a1[r] = a2[r] = 0u;
            lo[r] = wl[r][0];
            hi[r] = hist_on ? wh[r][0] : lo[r];   // no histogram wanted: nothing is "inside"
            sc[r] = sDzSc[gb[r] / GS];
            T2x[r] = all2 ? inf : T2[r];          // two classes only: the third never fires
            hrow[r] = whist + (long long)(gb[r] / GS) * wbuckets;
          }
          const int wb1 = (int)(wbuckets - 1);
          auto inside = [&](int r, double ad, unsigned k) {   // ~7 % of the pairs
            if (gb[r] / GS + k >= (unsigned)nb) return;         // beyond the last edge: not a class
            int bk = __double2int_rz(__dmul_rn(__dsub_rn(ad, lo[r]), sc[r]));   // saturates, then clamped
            bk = bk < wb1 ? bk : wb1;
            atomicAdd(hrow[r] + ((long long)k * wbuckets + bk), 1ull);
          };
          if (all1) {
#pragma unroll 4
            for (int jj = g0; jj < g0 + G; ++jj) {
              const double zj = sZ[jj];
#pragma unroll
              for (int r = 0; r < WR; ++r) {
                const double ad = fabs(zi[r] - zj);
                unsigned inw;
                asm("{\n\t.reg .pred x, w;\n\t.reg .u32 v;\n\t"
                    "setp.lt.f64 x, %2, %3;\n\t"
                    "selp.u32 v, 0x10001, 1, x;\n\t"
                    "add.u32 %0, %0, v;\n\t"
                    "setp.lt.and.f64 w, %2, %4, !x;\n\t"
                    "selp.u32 %1, 1, 0, w;\n\t}"
                    : "+r"(a0[r]), "=r"(inw) : "d"(ad), "d"(lo[r]), "d"(hi[r]));
                if (inw) inside(r, ad, 0u);
              }
            }
          } else {
#pragma unroll 2
            for (int jj = g0; jj < g0 + G; ++jj) {
              const double2 pj = sXY[jj];
              const double zj = sZ[jj];
#pragma unroll
              for (int r = 0; r < WR; ++r) {
                const double dx = xi[r] - pj.x;
                const double dy = yi[r] - pj.y;
                const double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                const double ad = fabs(zi[r] - zj);
                unsigned inw;
                asm("{\n\t.reg .pred x, p, q, w;\n\t.reg .u32 v;\n\t"
                    "setp.lt.f64 x, %4, %5;\n\t"
                    "selp.u32 v, 0x10001, 1, x;\n\t"
                    "add.u32 %0, %0, v;\n\t"
                    "setp.ge.f64 p, %7, %8;\n\t"
                    "setp.ge.f64 q, %7, %9;\n\t"
                    "@p add.u32 %1, %1, v;\n\t"
                    "@q add.u32 %2, %2, v;\n\t"
                    "setp.lt.and.f64 w, %4, %6, !x;\n\t"
                    "selp.u32 %3, 1, 0, w;\n\t}"
                    : "+r"(a0[r]), "+r"(a1[r]), "+r"(a2[r]), "=r"(inw)
                    : "d"(ad), "d"(lo[r]), "d"(hi[r]), "d"(s), "d"(T1[r]), "d"(T2x[r]));
                // the class offset is only needed inside the window: two more compares there
                if (inw) inside(r, ad, (s >= T1[r] ? 1u : 0u) + (s >= T2x[r] ? 1u : 0u));
              }
            }
          }
#pragma unroll
          for (int r = 0; r < WR; ++r) {
            const unsigned g8 = gb[r];
            const unsigned t0 = a0[r] - a1[r], t1 = a1[r] - a2[r], t2 = a2[r];  // field-wise: no borrow
            flush_row<4>(aSum + g8 * SROW, aCnt + g8 * CROW, (double)(t0 >> 16), 0.0, t0 & 0xffffu);
            if (!all1)
              flush_row<4>(aSum + (g8 + 8) * SROW, aCnt + (g8 + 8) * CROW, (double)(t1 >> 16), 0.0, t1 & 0xffffu);
            if (!all2)
              flush_row<4>(aSum + (g8 + 16) * SROW, aCnt + (g8 + 16) * CROW, (double)(t2 >> 16), 0.0, t2 & 0xffffu);
          }
          continue;
        }
#pragma unroll 2
        for (int jj = g0; jj < g0 + G; ++jj) {
          const double2 pj = sXY[jj];
          const double zj = sZ[jj];
#pragma unroll
          for (int r = 0; r < WR; ++r) {
            bool p1 = false, p2 = false;
            if (!all1) {  // one class for the whole group: no distance needed
              const double dx = xi[r] - pj.x;
              const double dy = yi[r] - pj.y;
              const double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
              p1 = s >= T1[r];
              p2 = !all2 && s >= T2[r];
            }
            const double ad = fabs(zi[r] - zj);
            const double lo = p2 ? wl[r][2] : (p1 ? wl[r][1] : wl[r][0]);
            const double hi = p2 ? wh[r][2] : (p1 ? wh[r][1] : wh[r][0]);
            const bool below = ad < lo;
            c1[r] += p1 ? 1u : 0u;
            c2[r] += p2 ? 1u : 0u;
            cb[r][0] += (below && !p1) ? 1u : 0u;
            cb[r][1] += (below && p1 && !p2) ? 1u : 0u;
            cb[r][2] += (below && p2) ? 1u : 0u;
            const unsigned lg = gb[r] / GS + (p2 ? 2u : (p1 ? 1u : 0u));
            if (whist && !below && ad < hi && lg < (unsigned)nb) {  // ~8 % of the in-range pairs
              long long bk = __double2ll_rz(__dmul_rn(__dsub_rn(ad, lo), sDzSc[lg]));
              bk = bk < wbuckets - 1 ? bk : wbuckets - 1;
              atomicAdd(&whist[(long long)lg * wbuckets + bk], 1ull);
            }
          }
        }
#pragma unroll
        for (int r = 0; r < WR; ++r) {
          const unsigned g8 = gb[r];
          flush_row<4>(aSum + g8 * SROW, aCnt + g8 * CROW, (double)cb[r][0], 0.0, (unsigned)G - c1[r]);
          if (!all1)
            flush_row<4>(aSum + (g8 + 8) * SROW, aCnt + (g8 + 8) * CROW, (double)cb[r][1], 0.0, c1[r] - c2[r]);
          if (!all2)
            flush_row<4>(aSum + (g8 + 16) * SROW, aCnt + (g8 + 16) * CROW, (double)cb[r][2], 0.0, c2[r]);
        }
        continue;
      }
      if (all1) {
        // one class for every pair of the group: no distance is evaluated at all
        double q[WR], rs[WR];
#pragma unroll
        for (int r = 0; r < WR; ++r) q[r] = rs[r] = 0.0;
        if (STATS != 0) {
#pragma unroll 4
          for (int jj = g0; jj < g0 + G; ++jj) {
            const double zj = sZ[jj];
#pragma unroll
            for (int r = 0; r < WR; ++r) {
              const double dz = zi[r] - zj;
              if (STATS & SKG_STAT_SUM_SQ) q[r] = fma(dz, dz, q[r]);
              if (STATS & SKG_STAT_SUM_SQRT) rs[r] += sqrt_abs(dz);
            }
          }
        }
#pragma unroll
        for (int r = 0; r < WR; ++r)
          flush_row<STATS>(aSum + gb[r] * SROW, aCnt + gb[r] * CROW, q[r], rs[r], (unsigned)G);
        continue;
      }
      // cumulative sums: A0 over all pairs of the group, A1 over s >= T1, A2 over s >= T2
      double q0[WR], q1[WR], q2[WR], r0[WR], r1[WR], r2[WR];
      unsigned c1[WR], c2[WR];
#pragma unroll
      for (int r = 0; r < WR; ++r) {
        q0[r] = q1[r] = q2[r] = r0[r] = r1[r] = r2[r] = 0.0;
        c1[r] = c2[r] = 0u;
      }
      if (all2) {
#pragma unroll WIN_UNROLL2
        for (int jj = g0; jj < g0 + G; ++jj) {
          const double2 pj = sXY[jj];
          const double zj = sZ[jj];
#pragma unroll
          for (int r = 0; r < WR; ++r) {
            const double dx = xi[r] - pj.x;
            const double dy = yi[r] - pj.y;
            const double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            const double dz = zi[r] - zj;
            if (STATS == 0) {
              asm("{\n\t.reg .pred p;\n\tsetp.ge.f64 p, %1, %2;\n\t@p add.u32 %0, %0, 1;\n\t}"
                  : "+r"(c1[r]) : "d"(s), "d"(T1[r]));
            }
            if (STATS & SKG_STAT_SUM_SQ) {
              q0[r] = fma(dz, dz, q0[r]);
              asm("{\n\t.reg .pred p;\n\tsetp.ge.f64 p, %3, %4;\n\t"
                  "@p fma.rn.f64 %0, %2, %2, %0;\n\t@p add.u32 %1, %1, 1;\n\t}"
                  : "+d"(q1[r]), "+r"(c1[r]) : "d"(dz), "d"(s), "d"(T1[r]));
            }
            if (STATS & SKG_STAT_SUM_SQRT) {
              const double rt = sqrt_abs(dz);
              r0[r] += rt;
              if (STATS & SKG_STAT_SUM_SQ) {
                asm("{\n\t.reg .pred p;\n\tsetp.ge.f64 p, %2, %3;\n\t@p add.rn.f64 %0, %0, %1;\n\t}"
                    : "+d"(r1[r]) : "d"(rt), "d"(s), "d"(T1[r]));
              } else {
                asm("{\n\t.reg .pred p;\n\tsetp.ge.f64 p, %3, %4;\n\t"
                    "@p add.rn.f64 %0, %0, %2;\n\t@p add.u32 %1, %1, 1;\n\t}"
                    : "+d"(r1[r]), "+r"(c1[r]) : "d"(rt), "d"(s), "d"(T1[r]));
              }
            }
          }
        }
      } else {
#pragma unroll 2
        for (int jj = g0; jj < g0 + G; ++jj) {
          const double2 pj = sXY[jj];
          const double zj = sZ[jj];
#pragma unroll
          for (int r = 0; r < WR; ++r) {
            const double dx = xi[r] - pj.x;
            const double dy = yi[r] - pj.y;
            const double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            const double dz = zi[r] - zj;
            if (STATS == 0) {
              asm("{\n\t.reg .pred p, q;\n\tsetp.ge.f64 p, %2, %3;\n\tsetp.ge.f64 q, %2, %4;\n\t"
                  "@p add.u32 %0, %0, 1;\n\t@q add.u32 %1, %1, 1;\n\t}"
                  : "+r"(c1[r]), "+r"(c2[r]) : "d"(s), "d"(T1[r]), "d"(T2[r]));
            }
            if (STATS & SKG_STAT_SUM_SQ) {
              q0[r] = fma(dz, dz, q0[r]);
              asm("{\n\t.reg .pred p, q;\n\tsetp.ge.f64 p, %5, %6;\n\tsetp.ge.f64 q, %5, %7;\n\t"
                  "@p fma.rn.f64 %0, %4, %4, %0;\n\t@q fma.rn.f64 %1, %4, %4, %1;\n\t"
                  "@p add.u32 %2, %2, 1;\n\t@q add.u32 %3, %3, 1;\n\t}"
                  : "+d"(q1[r]), "+d"(q2[r]), "+r"(c1[r]), "+r"(c2[r])
                  : "d"(dz), "d"(s), "d"(T1[r]), "d"(T2[r]));
            }
            if (STATS & SKG_STAT_SUM_SQRT) {
              const double rt = sqrt_abs(dz);
              r0[r] += rt;
              if (STATS & SKG_STAT_SUM_SQ) {
                asm("{\n\t.reg .pred p, q;\n\tsetp.ge.f64 p, %3, %4;\n\tsetp.ge.f64 q, %3, %5;\n\t"
                    "@p add.rn.f64 %0, %0, %2;\n\t@q add.rn.f64 %1, %1, %2;\n\t}"
                    : "+d"(r1[r]), "+d"(r2[r]) : "d"(rt), "d"(s), "d"(T1[r]), "d"(T2[r]));
              } else {
                asm("{\n\t.reg .pred p, q;\n\tsetp.ge.f64 p, %5, %6;\n\tsetp.ge.f64 q, %5, %7;\n\t"
                    "@p add.rn.f64 %0, %0, %4;\n\t@q add.rn.f64 %1, %1, %4;\n\t"
                    "@p add.u32 %2, %2, 1;\n\t@q add.u32 %3, %3, 1;\n\t}"
                    : "+d"(r1[r]), "+d"(r2[r]), "+r"(c1[r]), "+r"(c2[r])
                    : "d"(rt), "d"(s), "d"(T1[r]), "d"(T2[r]));
              }
            }
          }
        }
      }
      // one shared-memory update per class of the window; rows >= nb are never read
#pragma unroll
      for (int r = 0; r < WR; ++r) {
        const unsigned g8 = gb[r];
        flush_row<STATS>(aSum + g8 * SROW, aCnt + g8 * CROW, q0[r] - q1[r], r0[r] - r1[r], (unsigned)G - c1[r]);
        flush_row<STATS>(aSum + (g8 + 8) * SROW, aCnt + (g8 + 8) * CROW, q1[r] - q2[r], r1[r] - r2[r],
                         c1[r] - c2[r]);
        if (!all2)
          flush_row<STATS>(aSum + (g8 + 16) * SROW, aCnt + (g8 + 16) * CROW, q2[r], r2[r], c2[r]);
      }
    }
  }
  __syncthreads();
#ifdef SKG_WIN_DEBUG
  if ((tid & 31) == 0)
    for (int k = 0; k < 6; ++k)
      atomicAdd(const_cast<unsigned long long*>(&counters[8 + k]), dbg[k]);
#endif

  // fixed-order block reduction of the per-thread records
  HistPartial* out = partial + (int64_t)blockIdx.x * nb;
  const int warp = tid >> 5, lane = tid & 31;
  for (int gI = warp; gI < nb; gI += WB / 32) {
    unsigned long long c = 0;
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int k = 0; k < WB / 32; ++k) {
      const int th = lane + 32 * k;
      const unsigned char* sp = acc_raw + (size_t)(gI * WB + th) * SUMB;
      c += *reinterpret_cast<const unsigned int*>(acc_raw + (size_t)NR * (WB * SUMB) + (size_t)(gI * WB + th) * 4);
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
}

// ------------------------------------------------------------------ host ---

size_t hist_win_smem_bytes(int nb, int stats) {
  const int sumb = stats == 3 ? 16 : (stats == 0 ? 0 : (stats == 4 ? 4 : 8));
  return ((size_t)(nb + 3) * WB * (sumb + 4) + 15) & ~(size_t)15;
}

template <int STATS, int EMAX>
static int launch_hist_win_t(const HistLaunch& L) {
  auto kern = pair_hist_win_kernel<STATS, EMAX>;
  static thread_local size_t cached_smem = (size_t)-1;
  static thread_local int cached_occ = 0;
  if (cached_smem != L.smem) {
    SKG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
    SKG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cached_occ, kern, WB, L.smem));
    cached_smem = L.smem;
  }
  int occ = cached_occ;
  if (occ < 1) { set_error("pair_hist_win kernel does not fit (smem=%zu)", L.smem); return SKG_E_UNSUPPORTED; }
  occ = occ < MAX_GRID_PER_SM ? occ : MAX_GRID_PER_SM;
  int64_t grid = (int64_t)sm_count() * occ;
  if (grid > L.ntiles * L.js) grid = L.ntiles * L.js;
  if (grid > L.max_blocks) grid = L.max_blocks;
  *L.nblocks_out = (int)grid;
  kern<<<(unsigned)grid, WB, L.smem, L.st>>>(L.coords, L.values, L.n, *L.tab, L.partial, L.list,
                                             L.list_count, L.bbox32, L.bbox8, L.wlo, L.whi, L.js, L.dzt_dev,
                                             L.whist, L.wbuckets);
  SKG_CUDA(cudaGetLastError());
  return 0;
}

// 2-D, no directional mask, scalar values, stats in {0 (counts only), 1, 2, 3}.  The group size is planned on
// the device from the boxes the culling prelude just wrote (no host round trip).
int dispatch_hist_win(int emax, int stats, const HistLaunch& L) {
  int force_g = -1;
  if (const char* f = getenv("SKG_WIN_G")) {  // developer override: 0 (off), 8, 16, 32
    const int v = atoi(f);
    if (v == 0 || v == 8 || v == 16 || v == 32) force_g = v;
  }
  const int64_t nt = (L.n + TILE - 1) / TILE;
  double last;
  std::memcpy(&last, &L.tab->thr[L.tab->nb - 1], 8);
  win_plan_kernel<<<1, 256, 0, L.st>>>(L.bbox8, nt * (TILE / 8), last, L.tab->nb, force_g,
                                       const_cast<unsigned long long*>(L.list_count));
  SKG_CUDA(cudaGetLastError());
  if (stats == 4) return emax <= 1 ? launch_hist_win_t<4, 1>(L) : launch_hist_win_t<4, 0>(L);
  if (emax <= 1) {
    if (stats == 0) return launch_hist_win_t<0, 1>(L);
    if (stats == 1) return launch_hist_win_t<1, 1>(L);
    if (stats == 2) return launch_hist_win_t<2, 1>(L);
    return launch_hist_win_t<3, 1>(L);
  }
  if (stats == 0) return emax == 2 ? launch_hist_win_t<0, 2>(L) : launch_hist_win_t<0, 0>(L);  // count sweeps of the
  // select cluster their thresholds: two per table cell is common there
  if (stats == 1) return launch_hist_win_t<1, 0>(L);
  if (stats == 2) return launch_hist_win_t<2, 0>(L);
  return launch_hist_win_t<3, 0>(L);
}

}  // namespace skg
