/* ORACLE - TEST INFRASTRUCTURE, NOT PRODUCT.  Nothing under scikit-gstat_b200/ may call this.
 *
 * Plain-C restatement (gcc; rows of the pair triangle are dealt to threads) of the per-pair
 * arithmetic of scikit-gstat 1.0.23's experimental-variogram path, for sizes at which the
 * reference's own O(N^2) arrays do not fit (BASELINE.json configs 3 and 4).  It follows, pair
 * by pair, what the reference computes on its materialised vectors:
 *
 *   d  = sqrt(dx*dx + dy*dy), dx = x_i - x_j, i < j     scipy pdist, MetricSpace.py:151-153
 *        (no FMA: compile with -ffp-contract=off; operand order as scipy's euclidean kernel)
 *   dz = |z_i - z_j|                                     Variogram.py:1938-1944
 *   g  = #{k : edges[k] <= d}; g == n_lags -> dropped    np.digitize, Variogram.py:2005-2006
 *   mask (optional)                                      DirectionalVariogram.py:398-413 (angle),
 *                                                        :704-714 (_triangle), :776-782 (_compass)
 * and returns exact integer counts, long-double sums and key histograms from which the
 * Python side (oracle/make_fullsize_golden.py) selects exact order statistics (np.median,
 * np.percentile, np.nanpercentile, np.nanmedian: Variogram.py:1288-1293, binning.py:70-79,
 * DirectionalVariogram.py:509, estimators.py:178).
 *
 * The lag class is found on d = sqrt(s) against the edges themselves (binary search), NOT in
 * squared space: it is independent of the product's threshold trick and therefore checks it.
 * The mask uses libm acos/sin while the reference uses numpy's; pairs within 1e-12 of a decision
 * boundary are counted in *n_guard so the caller can see whether that could matter (it is 0
 * for the generic float inputs of the golden cases).
 *
 * Every function sweeps the rows i = row_begin, row_begin + row_step, ... of the pair triangle
 * single-threaded; oracle/pairs.py runs row_step such calls on Python threads (ctypes drops
 * the GIL) and adds the integer / long-double partials.  (This image's gcc has no libgomp.)
 *
 * Build: oracle/Makefile  ->  oracle/_build/libpairs_oracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int model;        /* 0 none, 1 triangle, 2 compass */
  double az;        /* np.radians(azimuth)            */
  double half_tol;  /* np.radians(tolerance / 2)      */
  double half_bw;   /* bandwidth / 2                  */
} dir_t;

static inline uint64_t bits_of(double v) {
  uint64_t b;
  memcpy(&b, &v, 8);
  return b;
}

/* DirectionalVariogram.py:398-413 + :704-714 / :776-782, literal */
static inline int mask_pass(double dx, double dy, double d, const dir_t* p, int* near) {
  const double PI = 3.141592653589793;
  double pos = acos(dx / d); /* NaN for coincident points: every test false */
  double ang = (dy >= 0.0) ? pos : -pos;
  double t = fabs(ang + p->az);
  double a = t;
  if (a > PI) a = a - PI;
  if (a > PI / 2) a = PI - a;
  int in_tol = a <= p->half_tol;
  if (fabs(a - p->half_tol) <= 1e-12) *near = 1;
  if (p->model == 2) return in_tol;
  double w = fabs(d * sin(t));
  int in_band = p->half_bw >= w;
  if (fabs(p->half_bw - w) <= 1e-12 * (p->half_bw + 1.0)) *near = 1;
  return in_tol && in_band;
}

static inline int lag_of(double d, const double* edges, int n_lags) {
  /* np.digitize(d, edges): number of edges <= d (edges ascending); NaN -> n_lags */
  if (!(d == d)) return n_lags;
  int lo = 0, hi = n_lags;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (edges[mid] <= d) lo = mid + 1; else hi = mid;
  }
  return lo;
}

/* Exact order statistics need the keys (d, key_kind 0, or |dz|, key_kind 1) of a (mask-passing)
 * pair population split by lag class g < n_lags (n_lags == 0: one class holding every pair with a
 * non-NaN distance).  Three sweeps narrow the keys of a wanted rank down to a handful:
 *   level 1: per class, histogram of b1 = bucket(key, lo1[g], sc1[g], nb1)
 *   level 2: per TARGET k = (class seg[k], level-1 bucket want1[k]): histogram of
 *            b2 = bucket(key, lo2[k], sc2[k], nb2) over the keys with b1 == want1[k]
 *   collect: per target k = (seg[k], want1[k], want2[k]): the keys themselves
 * bucket(key, lo, sc, nb) = clamp((int64)((key - lo) * sc), 0, nb - 1): the same C expression in
 * every sweep, so bucket membership is consistent whatever its rounding. */
static inline int64_t bucket_of(double key, double lo, double sc, int64_t nb) {
  double t = (key - lo) * sc;
  if (!(t > 0.0)) return 0;
  if (t >= (double)nb) return nb - 1;
  return (int64_t)t;
}

/* mode 0: level-1 histograms hist[g*nb1 + b1]; mode 1: level-2 histograms hist[k*nb2 + b2];
 * mode 2: collect, out[k*cap + pos], count[k] (keys beyond cap are counted, not stored). */
int oracle_key_sweep(const double* x, const double* y, const double* z, int64_t n, const double* dirp,
                     const double* edges, int n_lags, int key_kind, int mode,
                     const double* lo1, const double* sc1, int64_t nb1,
                     int n_targets, const int* seg, const int64_t* want1,
                     const double* lo2, const double* sc2, int64_t nb2, const int64_t* want2,
                     uint64_t* hist, double* out, int64_t cap, int64_t* count,
                     uint64_t* n_guard, int64_t row_begin, int64_t row_step) {
  dir_t dp = {0, 0, 0, 0};
  if (dirp) { dp.model = (int)dirp[0]; dp.az = dirp[1]; dp.half_tol = dirp[2]; dp.half_bw = dirp[3]; }
  const int nseg = n_lags > 0 ? n_lags : 1;
  double far2 = INFINITY;
  if (n_lags > 0 && !dp.model) {
    const double last = edges[n_lags - 1];
    far2 = last * last * (1.0 + 1e-9); /* beyond every edge: no sqrt needed */
  }
  uint64_t guard = 0;
  for (int64_t i = row_begin; i < n - 1; i += row_step) {
    const double xi = x[i], yi = y[i], zi = z ? z[i] : 0.0;
    for (int64_t j = i + 1; j < n; ++j) {
      const double dx = xi - x[j], dy = yi - y[j];
      const double s = dx * dx + dy * dy;
      if (s > far2) continue;
      const double d = sqrt(s);
      int g = n_lags > 0 ? lag_of(d, edges, n_lags) : ((d == d) ? 0 : 1);
      if (g >= nseg) continue;
      if (dp.model) {
        int near = 0;
        const int pass = mask_pass(dx, dy, d, &dp, &near);
        guard += near;
        if (!pass) continue;
      }
      const double key = key_kind == 0 ? d : fabs(zi - z[j]);
      if (!(key == key)) continue; /* NaN observation: np.nanmedian ignores it */
      const int64_t b1 = bucket_of(key, lo1[g], sc1[g], nb1);
      if (mode == 0) { hist[(size_t)g * nb1 + b1]++; continue; }
      for (int k = 0; k < n_targets; ++k) {
        if (seg[k] != g || want1[k] != b1) continue;
        const int64_t b2 = bucket_of(key, lo2[k], sc2[k], nb2);
        if (mode == 1) { hist[(size_t)k * nb2 + b2]++; continue; }
        if (want2[k] != b2) continue;
        const int64_t p = count[k]++;
        if (p < cap) out[(size_t)k * cap + p] = key;
      }
    }
  }
  if (n_guard) *n_guard = guard;
  return 0;
}

/* Per lag class: exact pair count, sum dz^2 and sum sqrt(dz) ADDED to long double outputs (64-bit mantissa:
 * the rounding of the sums is far below the 1e-12 tolerance of the parity tests; the reference
 * itself sums each class sequentially in double, estimators.py:63-66, :123). */
int oracle_lag_hist(const double* x, const double* y, const double* z, int64_t n, const double* dirp,
                    const double* edges, int n_lags, uint64_t* count, long double* sum_sq,
                    long double* sum_sqrt, uint64_t* n_masked, uint64_t* n_guard, int64_t row_begin,
                    int64_t row_step) {
  dir_t dp = {0, 0, 0, 0};
  if (dirp) { dp.model = (int)dirp[0]; dp.az = dirp[1]; dp.half_tol = dirp[2]; dp.half_bw = dirp[3]; }
  const double last = edges[n_lags - 1];
  const double far2 = last * last * (1.0 + 1e-9); /* beyond every edge: no sqrt needed */
  uint64_t guard_total = 0, masked_total = 0;
  {
    long double* lsq = sum_sq;
    long double* lrt = sum_sqrt;
    uint64_t* lc = count;
    uint64_t guard = 0, masked = 0;
    for (int64_t i = row_begin; i < n - 1; i += row_step) {
      const double xi = x[i], yi = y[i], zi = z[i];
      for (int64_t j = i + 1; j < n; ++j) {
        const double dx = xi - x[j], dy = yi - y[j];
        const double s = dx * dx + dy * dy;
        if (!dp.model && s > far2) continue;
        const double d = sqrt(s);
        int pass = 1;
        if (dp.model) {
          int near = 0;
          pass = mask_pass(dx, dy, d, &dp, &near);
          guard += near;
          masked += pass;
        }
        if (!pass) continue;
        const int g = lag_of(d, edges, n_lags);
        if (g >= n_lags) continue;
        const double dz = fabs(zi - z[j]);
        lc[g]++;
        lsq[g] += (long double)dz * (long double)dz;
        lrt[g] += (long double)sqrt(dz);
      }
    }
    guard_total += guard;
    masked_total += masked;
  }
  if (n_guard) *n_guard = guard_total;
  if (n_masked) *n_masked = masked_total;
  return 0;
}
