// K5: batched ordinary kriging for sm_100a - one warp per target point.
//
// Restates OrdinaryKriging.transform (Kriging.py:405-650) for mode='exact':
//   neighbour search  = MetricSpacePair.find_closest (MetricSpace.py:26-71):
//                       d <= radius, the max_points smallest by (d, sample id);
//   system            = Kriging.py:594-614 (gamma off-diagonal, 0 diagonal,
//                       unit border, zero corner; rhs gamma(d_0i), 1);
//   solve             = partial-pivot LU in fp64 (the reference multiplies by
//                       numpy.linalg.inv, i.e. LAPACK dgetrf/dgetri);
//   outputs           = Kriging.py:644-647.
// The M x N target-sample distance matrix (MetricSpace.py:248-253) and the
// N x N sample matrix (MetricSpace.py:158-187) are never formed: samples are
// bucketed in a uniform cell grid and each warp expands Chebyshev rings of
// cells around its target until the max_points nearest are provably inside.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace skg {

struct GridParams {
  double ox, oy, oz, h;
  int nx, ny, nz;
};

struct ModelParams {
  int model;
  double r, c0, s, b;
  double a;  // derived range parameter (models.py: r/1, r/3, r/2, r/3^(1/s))
  double gnorm;  // matern: 2 / Gamma(s)
  // mode='estimate' (Kriging.py:656-679): semivariances of the kriging MATRIX come from a table
  // of `lut_size` values over [0, lut_range]; beyond the table: lut_out (the sill)
  const double* lut;
  int lut_size;
  double lut_range, lut_out;
};

__device__ __forceinline__ int cell_coord(double p, double o, double inv_h) {
  double f = floor((p - o) * inv_h);
  f = fmin(fmax(f, -1.0e9), 1.0e9);
  return (int)f;
}

// ------------------------------------------------------------ grid build ---

template <int DIM>
__global__ void grid_count_kernel(const double* __restrict__ coords, int64_t n, GridParams g,
                                  int* __restrict__ cellid, int* __restrict__ counts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double inv_h = 1.0 / g.h;
  int cx = min(max(cell_coord(coords[i], g.ox, inv_h), 0), g.nx - 1);
  int cy = min(max(cell_coord(coords[n + i], g.oy, inv_h), 0), g.ny - 1);
  int cz = 0;
  if (DIM == 3) cz = min(max(cell_coord(coords[2 * n + i], g.oz, inv_h), 0), g.nz - 1);
  const int c = (cz * g.ny + cy) * g.nx + cx;
  cellid[i] = c;
  atomicAdd(&counts[c], 1);
}

// single-CTA exclusive scan: cell_start[0..ncell], cursor[c] = cell_start[c]
__global__ void grid_scan_kernel(int* __restrict__ counts, int ncell, int* __restrict__ cell_start) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < ncell; base += blockDim.x) {
    const int idx = base + threadIdx.x;
    const int v = idx < ncell ? counts[idx] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int t = lane < (int)(blockDim.x >> 5) ? warp_tot[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      warp_tot[lane] = t;
    }
    __syncthreads();
    const int before = carry + (warp ? warp_tot[warp - 1] : 0) + (x - v);
    if (idx < ncell) {
      cell_start[idx] = before;
      counts[idx] = before;  // becomes the scatter cursor
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) cell_start[ncell] = carry;
}

__global__ void grid_scatter_kernel(const int* __restrict__ cellid, int64_t n,
                                    int* __restrict__ cursor, int* __restrict__ order) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int pos = atomicAdd(&cursor[cellid[i]], 1);
  order[pos] = (int)i;
}

// deterministic order inside each cell: ascending sample id
__global__ void grid_sort_cells_kernel(const int* __restrict__ cell_start, int ncell,
                                       int* __restrict__ order) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  const int a = cell_start[c], b = cell_start[c + 1];
  for (int i = a + 1; i < b; ++i) {
    const int v = order[i];
    int j = i - 1;
    while (j >= a && order[j] > v) {
      order[j + 1] = order[j];
      --j;
    }
    order[j + 1] = v;
  }
}

// ---------------------------------------------------------------- models ---

// Modified Bessel function of the second kind K_nu(x), real nu >= 0, x > 0 (scipy.special.kv in
// models.py:419-422).  Temme's series for x < 2, Steed's continued fraction CF2 above, upward
// recurrence from |mu| <= 1/2 (the classical bessik scheme).
__device__ __noinline__ double bessel_k(double nu, double x) {
  const double EPS = 1e-16, PI = 3.141592653589793;
  const int nl = (int)(nu + 0.5);
  const double xmu = nu - nl, xmu2 = xmu * xmu;
  const double xi = 1.0 / x, xi2 = 2.0 * xi;
  double rkmu, rk1;
  if (x < 2.0) {
    const double x2 = 0.5 * x, pimu = PI * xmu;
    const double fact = fabs(pimu) < 1e-8 ? 1.0 : pimu / sin(pimu);
    double d = -log(x2);
    double e = xmu * d;
    const double fact2 = fabs(e) < 1e-8 ? 1.0 : sinh(e) / e;
    const double gampl = 1.0 / tgamma(1.0 + xmu), gammi = 1.0 / tgamma(1.0 - xmu);
    const double gam2 = 0.5 * (gammi + gampl);
    // (1/Gamma(1-mu) - 1/Gamma(1+mu)) / (2 mu); odd part of the 1/Gamma(1+z) series near 0
    const double gam1 = fabs(xmu) < 0.01
        ? -(0.5772156649015329 + xmu2 * (-0.0420026350340952 + xmu2 * -0.0421977345555443))
        : (gammi - gampl) / (2.0 * xmu);
    double ff = fact * (gam1 * cosh(e) + gam2 * fact2 * d);
    double sum = ff;
    e = exp(e);
    double p = 0.5 * e / gampl, q = 0.5 / (e * gammi), c = 1.0;
    d = x2 * x2;
    double sum1 = p;
    for (int i = 1; i <= 500; ++i) {
      ff = (i * ff + p + q) / ((double)i * i - xmu2);
      c *= d / i;
      p /= (i - xmu);
      q /= (i + xmu);
      const double del = c * ff;
      sum += del;
      sum1 += c * (p - i * ff);
      if (fabs(del) < fabs(sum) * EPS) break;
    }
    rkmu = sum;
    rk1 = sum1 * xi2;
  } else {
    double b = 2.0 * (1.0 + x), d = 1.0 / b, h = d, delh = d;
    double q1 = 0.0, q2 = 1.0;
    const double a1 = 0.25 - xmu2;
    double q = a1, c = a1, a = -a1;
    double sacc = 1.0 + q * delh;
    for (int i = 2; i <= 500; ++i) {
      a -= 2 * (i - 1);
      c = -a * c / i;
      const double qnew = (q1 - b * q2) / a;
      q1 = q2;
      q2 = qnew;
      q += c * qnew;
      b += 2.0;
      d = 1.0 / (b + a * d);
      delh = (b * d - 1.0) * delh;
      h += delh;
      const double dels = q * delh;
      sacc += dels;
      if (fabs(dels / sacc) < EPS) break;
    }
    h = a1 * h;
    rkmu = sqrt(PI / (2.0 * x)) * exp(-x) / sacc;
    rk1 = rkmu * (xmu + x + 0.5 - h) * xi;
  }
  for (int i = 1; i <= nl; ++i) {
    const double t = (xmu + i) * xi2 * rk1 + rkmu;
    rkmu = rk1;
    rk1 = t;
  }
  return rkmu;
}

// models.py:413-422; kept out of line so that the common models do not pay its registers
__device__ __noinline__ double gamma_matern(double h, double a, double s, double b, double c0, double gnorm) {
  if (h == 0.0) return b;
  const double u = (h * sqrt(s)) / a;
  const double kv = bessel_k(s, 2.0 * u);
  if (kv == 0.0) return b + c0;  // K underflows far outside the range: gamma = sill
  return b + c0 * (1.0 - gnorm * pow(u, s) * kv);
}

__device__ __forceinline__ double gamma_model(double h, const ModelParams& m) {
  switch (m.model) {
    case SKG_MODEL_MATERN:
      return gamma_matern(h, m.a, m.s, m.b, m.c0, m.gnorm);
    case SKG_MODEL_SPHERICAL: {  // models.py:77-82
      if (h <= m.r) {
        const double q = h / m.a;
        return m.b + m.c0 * ((1.5 * q) - (0.5 * (q * q * q)));
      }
      return m.b + m.c0;
    }
    case SKG_MODEL_EXPONENTIAL:  // models.py:142-144
      return m.b + m.c0 * (1.0 - exp(-(h / m.a)));
    case SKG_MODEL_GAUSSIAN:  // models.py:207-209
      return m.b + m.c0 * (1.0 - exp(-((h * h) / (m.a * m.a))));
    case SKG_MODEL_CUBIC: {  // models.py:264-272
      if (h < m.r) {
        const double a2 = m.a * m.a, h2 = h * h;
        const double a3 = a2 * m.a, h3 = h2 * h;
        const double a5 = a3 * a2, h5 = h3 * h2;
        const double a7 = a5 * a2, h7 = h5 * h2;
        return m.b + m.c0 * ((7.0 * (h2 / a2)) - (8.75 * (h3 / a3)) + (3.5 * (h5 / a5)) -
                             (0.75 * (h7 / a7)));
      }
      return m.b + m.c0;
    }
    default: {  // stable, models.py:336-344
      if (h == 0.0) return m.b;
      return m.b + m.c0 * (1.0 - exp(-pow(h / m.a, m.s)));
    }
  }
}

// semivariance of a kriging-matrix entry: exact model, or the 'estimate' table
// (Kriging.py:664-679: index = int((d / range) * precision), no FMA)
__device__ __forceinline__ double gamma_matrix(double h, const ModelParams& m) {
  if (m.lut) {
    const int idx = (int)__dmul_rn(__ddiv_rn(h, m.lut_range), (double)m.lut_size);
    return idx >= m.lut_size ? m.lut_out : m.lut[idx];
  }
  return gamma_model(h, m);
}

// ------------------------------------------------------------ the kernel ---

struct KrigeArgs {
  const double* coords;
  const double* values;
  int64_t n;
  GridParams grid;
  const int* cell_start;
  const int* order;
  const double* targets;
  int64_t m;
  ModelParams model;
  double radius;
  int min_points, max_points;
  double* z;
  double* sigma;
  int* status;
  int64_t t0, t1;
  int cap;           // candidate list capacity per warp
  int region_bytes;  // per-warp shared region (list, later the matrix)
  int retry_only;    // LU kernel: only targets the Cholesky kernel marked KRIGE_RETRY
  const int* exclude;  // per target: sample id that must not be a neighbour (-1: none), or NULL
  const double* sample_cov;  // [n][n] c_t - gamma_matrix(d_ij) of all sample pairs (skg_krige_sample_cov) or NULL
};

constexpr int KRIGE_RETRY = 3;  // internal: covariance matrix not positive definite in fp64

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w < v ? w : v;
  }
  return v;
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  return v;
}
__device__ __forceinline__ int warp_sum_int(int v) { return __reduce_add_sync(0xffffffffu, v); }

// Keep the k smallest (key, id) of the warp's list; returns new length (k).
// Bisection on the key bit pattern, ties on the k-th key broken by ascending
// sample id (np.argsort(kind='stable') on index-ordered input, MetricSpace.py:69).
__device__ int prune_list(unsigned long long* keys, int* ids, int len, int k, int lane) {
  unsigned long long lo = ~0ull, hi = 0ull;
  for (int e = lane; e < len; e += 32) {
    const unsigned long long v = keys[e];
    lo = v < lo ? v : lo;
    hi = v > hi ? v : hi;
  }
  lo = warp_min_u64(lo);
  hi = warp_max_u64(hi);
  bool exact = false;
  unsigned long long T = hi;
  while (lo < hi) {
    const unsigned long long mid = lo + ((hi - lo) >> 1);
    int c = 0;
    for (int e = lane; e < len; e += 32) c += keys[e] <= mid ? 1 : 0;
    c = warp_sum_int(c);
    if (c == k) { T = mid; exact = true; break; }
    if (c > k) hi = mid; else lo = mid + 1;
  }
  int idcut = 0x7fffffff;
  if (!exact) {
    T = lo;
    int below = 0;
    for (int e = lane; e < len; e += 32) below += keys[e] < T ? 1 : 0;
    below = warp_sum_int(below);
    const int need = k - below;  // >= 1 ties to take, smallest ids first
    int ilo = 0, ihi = 0x7fffffff;
    while (ilo < ihi) {
      const int mid = ilo + ((ihi - ilo) >> 1);
      int c = 0;
      for (int e = lane; e < len; e += 32) c += (keys[e] == T && ids[e] <= mid) ? 1 : 0;
      c = warp_sum_int(c);
      if (c >= need) ihi = mid; else ilo = mid + 1;
    }
    idcut = ilo;
  }
  // stable in-place compaction (reads of a chunk complete before its writes; writes
  // only go to positions <= the chunk's first read position)
  int out = 0;
  for (int base = 0; base < len; base += 32) {
    const int e = base + lane;
    unsigned long long v = 0ull;
    int id = 0;
    bool keep = false;
    if (e < len) {
      v = keys[e];
      id = ids[e];
      keep = (v < T) || (v == T && (exact || id <= idcut));
    }
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) {
      const int pos = out + __popc(mask & ((1u << lane) - 1u));
      keys[pos] = v;
      ids[pos] = id;
    }
    out += __popc(mask);
    __syncwarp();
  }
  return out;
}

// Neighbour search of one warp for target (tx, ty, tw): fills keys[]/ids[] (shared, per
// warp) with the <= max_points nearest samples within the radius and returns their number.
template <int DIM>
__device__ __forceinline__ int find_neighbours(const KrigeArgs& a, double tx, double ty, double tw,
                                               int excl, unsigned long long* keys, int* ids,
                                               int lane) {
  const double* __restrict__ X = a.coords;
  const double* __restrict__ Y = a.coords + a.n;
  const double* __restrict__ W = a.coords + 2 * a.n;
  const GridParams g = a.grid;
  const double inv_h = 1.0 / g.h;
  const int rmax = (int)ceil(a.radius * inv_h) + 1;
  const int maxp = a.max_points;
  const int cx = cell_coord(tx, g.ox, inv_h);
    const int cy = cell_coord(ty, g.oy, inv_h);
    const int cz = DIM == 3 ? cell_coord(tw, g.oz, inv_h) : 0;
    int len = 0;
    double rad = a.radius;  // shrinks to the k-th distance after a prune
    // rings that can contain anything: up to the farthest grid cell
    int far = max(max(cx, g.nx - 1 - cx), max(cy, g.ny - 1 - cy));
    if (DIM == 3) far = max(far, max(cz, g.nz - 1 - cz));
    const int rstop = min(rmax, max(far, 0));
    bool nan_target = !(tx == tx) || !(ty == ty) || (DIM == 3 && !(tw == tw));

    for (int r = 0; r <= rstop && !nan_target; ++r) {
      const int z0 = DIM == 3 ? cz - r : 0, z1 = DIM == 3 ? cz + r : 0;
      for (int zz = z0; zz <= z1; ++zz) {
        if (zz < 0 || zz >= g.nz) continue;
        const bool zface = DIM == 3 && (zz == cz - r || zz == cz + r);
        for (int yy = cy - r; yy <= cy + r; ++yy) {
          if (yy < 0 || yy >= g.ny) continue;
          const bool full_row = zface || yy == cy - r || yy == cy + r;
          // full row: cells cx-r..cx+r (one contiguous run); else only cx-r and cx+r
          const int nruns = full_row ? 1 : (r == 0 ? 1 : 2);
          for (int run = 0; run < nruns; ++run) {
            int xa, xb;
            if (full_row) { xa = cx - r; xb = cx + r; }
            else { xa = xb = (run == 0 ? cx - r : cx + r); }
            xa = max(xa, 0);
            xb = min(xb, g.nx - 1);
            if (xa > xb) continue;
            const int rowbase = (zz * g.ny + yy) * g.nx;
            const int s0 = a.cell_start[rowbase + xa];
            const int s1 = a.cell_start[rowbase + xb + 1];
            for (int sb = s0; sb < s1; sb += 32) {
              if (len + 32 > a.cap) {  // make room: keep the best max_points so far
                __syncwarp();
                len = prune_list(keys, ids, len, maxp, lane);
                unsigned long long kmax = 0ull;
                for (int e = lane; e < len; e += 32) kmax = keys[e] > kmax ? keys[e] : kmax;
                kmax = warp_max_u64(kmax);
                rad = fmin(rad, __longlong_as_double((long long)kmax));
                __syncwarp();
              }
              const int sidx = sb + lane;
              bool hit = false;
              double d = 0.0;
              int id = 0;
              if (sidx < s1) {
                id = a.order[sidx];
                const double dx = tx - X[id];
                const double dy = ty - Y[id];
                double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                if (DIM == 3) {
                  const double dw = tw - W[id];
                  s = __dadd_rn(s, __dmul_rn(dw, dw));
                }
                d = sqrt(s);  // cdist value, MetricSpace.py:248-253
                hit = d <= rad && id != excl;  // MetricSpace.py:61 (excl: leave-one-out)
              }
              const unsigned mask = __ballot_sync(0xffffffffu, hit);
              if (hit) {
                const int pos = len + __popc(mask & ((1u << lane) - 1u));
                keys[pos] = (unsigned long long)__double_as_longlong(d);
                ids[pos] = id;
              }
              len += __popc(mask);
            }
          }
        }
      }
      // everything within r*h of the target has been seen after ring r
      __syncwarp();
      if (len >= maxp) {
        const double safe = (double)r * g.h * (1.0 - 9.5367431640625e-07);
        int c = 0;
        for (int e = lane; e < len; e += 32)
          c += __longlong_as_double((long long)keys[e]) <= safe ? 1 : 0;
        c = warp_sum_int(c);
        if (c >= maxp) break;
      }
    }
    __syncwarp();
    if (len > maxp) len = prune_list(keys, ids, len, maxp, lane);
    __syncwarp();
    return len;
}

template <int DIM, int NQ>  // NQ: neighbour slots per lane (max_points <= 32 * NQ)
__global__ void __launch_bounds__(256)
krige_kernel(const __grid_constant__ KrigeArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const int maxp = a.max_points;
  // per-warp layout: region | nbr coords [DIM][maxp] | nbr values [maxp] | rhs copy [maxp+1] | mult [maxp+1]
  const int per_warp = a.region_bytes + (DIM + 1) * maxp * 8 + 2 * (maxp + 1) * 8;
  unsigned char* base = smem_raw + (size_t)warp * per_warp;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(base);
  int* ids = reinterpret_cast<int*>(base + (size_t)a.cap * 8);
  double* A = reinterpret_cast<double*>(base);
  double* ncoord = reinterpret_cast<double*>(base + a.region_bytes);
  double* nval = ncoord + DIM * maxp;
  double* rhs = nval + maxp;
  double* mult = rhs + (maxp + 1);

  const double* __restrict__ X = a.coords;
  const double* __restrict__ Y = a.coords + a.n;
  const double* __restrict__ W = a.coords + 2 * a.n;
  const GridParams g = a.grid;
  const double inv_h = 1.0 / g.h;
  const int rmax = (int)ceil(a.radius * inv_h) + 1;
  const double qnan = __longlong_as_double(NAN_BITS);

  for (int64_t t = a.t0 + (int64_t)blockIdx.x * wpb + warp; t < a.t1; t += (int64_t)gridDim.x * wpb) {
    if (a.retry_only && a.status[t] != KRIGE_RETRY) continue;
    const double tx = a.targets[t];
    const double ty = a.targets[a.m + t];
    const double tw = DIM == 3 ? a.targets[2 * a.m + t] : 0.0;
    const int len = find_neighbours<DIM>(a, tx, ty, tw, a.exclude ? a.exclude[t] : -1, keys, ids, lane);
    __syncwarp();
    const int nn = len;
    if (nn < a.min_points || nn < 1) {  // Kriging.py:579-580 LessPointsError
      if (lane == 0) {
        a.z[t] = qnan;
        a.sigma[t] = qnan;
        a.status[t] = SKG_KRIGE_LESS_POINTS;
      }
      __syncwarp();
      continue;
    }
    // pull the neighbours out of the list region before it becomes the matrix
    double kd[NQ];
    int kid[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int e = lane + 32 * q;
      kd[q] = e < nn ? __longlong_as_double((long long)keys[e]) : 0.0;
      kid[q] = e < nn ? ids[e] : 0;
    }
    __syncwarp();
    const int N = nn + 1;   // system size
    const int LD = (N + 1) | 1;  // augmented with the rhs column (index N); odd stride keeps column walks bank-conflict free
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int e = lane + 32 * q;
      if (e < nn) {
        ncoord[e] = X[kid[q]];
        ncoord[maxp + e] = Y[kid[q]];
        if (DIM == 3) ncoord[2 * maxp + e] = W[kid[q]];
        nval[e] = a.values[kid[q]];
        const double gv = gamma_model(kd[q], a.model);  // Kriging.py:611-613
        rhs[e] = gv;
        A[e * LD + N] = gv;
        A[e * LD + e] = 0.0;       // squareform diagonal, Kriging.py:600
        A[e * LD + nn] = 1.0;      // Kriging.py:601
        A[nn * LD + e] = 1.0;      // Kriging.py:602
      }
    }
    if (lane == 0) {
      A[nn * LD + nn] = 0.0;  // Kriging.py:604
      A[nn * LD + N] = 1.0;
      rhs[nn] = 1.0;
    }
    __syncwarp();
    const int npairs = nn * (nn - 1) / 2;
    for (int p0 = lane; p0 < npairs; p0 += 64) {
      // two independent pairs per trip: their loads, sqrt and exp chains overlap
      int pi[2], pj[2];
      double gv[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int p = min(p0 + 32 * q, npairs - 1);
        // p -> (i, j), i < j, row-major upper triangle
        int i = (int)((2.0 * nn - 1.0 - sqrt((2.0 * nn - 1.0) * (2.0 * nn - 1.0) - 8.0 * p)) * 0.5);
        while (i > 0 && i * nn - i * (i + 1) / 2 > p) --i;
        while ((i + 1) * nn - (i + 1) * (i + 2) / 2 <= p) ++i;
        pi[q] = i;
        pj[q] = p - (i * nn - i * (i + 1) / 2) + i + 1;
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int i = pi[q], j = pj[q];
        const double dx = ncoord[i] - ncoord[j];
        const double dy = ncoord[maxp + i] - ncoord[maxp + j];
        double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        if (DIM == 3) {
          const double dw = ncoord[2 * maxp + i] - ncoord[2 * maxp + j];
          s = __dadd_rn(s, __dmul_rn(dw, dw));
        }
        gv[q] = gamma_matrix(sqrt(s), a.model);  // Kriging.py:594, 652-679
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (p0 + 32 * q < npairs) {
          A[pi[q] * LD + pj[q]] = gv[q];
          A[pj[q] * LD + pi[q]] = gv[q];
        }
      }
    }
    __syncwarp();

    // ---- partial-pivot Gaussian elimination on [A | b] -------------------
    bool singular = false;
    for (int k = 0; k < N; ++k) {
      double best = -1.0;
      int brow = k;
      for (int i = k + lane; i < N; i += 32) {
        const double v = fabs(A[i * LD + k]);
        if (v > best) { best = v; brow = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int orow = __shfl_xor_sync(0xffffffffu, brow, o);
        if (ob > best || (ob == best && orow < brow)) { best = ob; brow = orow; }
      }
      if (!(best > 0.0)) { singular = true; break; }  // zero (or NaN) pivot column
      if (brow != k) {
        for (int j = k + lane; j < LD; j += 32) {
          const double u = A[k * LD + j];
          A[k * LD + j] = A[brow * LD + j];
          A[brow * LD + j] = u;
        }
      }
      __syncwarp();
      const double rp = 1.0 / A[k * LD + k];
      for (int i = k + 1 + lane; i < N; i += 32) mult[i] = A[i * LD + k] * rp;
      __syncwarp();
      // rank-1 update of the trailing block.  Lane owns columns j0, j0+32, j0+64 (those
      // that exist) and walks the rows in batches of 4: all loads of a batch are issued
      // before its FMAs/stores, so the shared-memory latency is paid once per batch, not
      // once per element.
      if (NQ == 2) {
        const int NC = N + 1;  // real columns (rhs included); padding columns are skipped
        const int j0 = k + 1 + lane;
        const bool c0 = j0 < NC, c1 = j0 + 32 < NC, c2 = j0 + 64 < NC;
        const double u0 = c0 ? A[k * LD + j0] : 0.0;
        const double u1 = c1 ? A[k * LD + j0 + 32] : 0.0;
        const double u2 = c2 ? A[k * LD + j0 + 64] : 0.0;
        if (c0) {
          for (int i0 = k + 1; i0 < N; i0 += 4) {
            double m[4], a0[4], a1[4], a2[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int i = min(i0 + q, N - 1);
              m[q] = mult[i];
              a0[q] = A[i * LD + j0];
              a1[q] = c1 ? A[i * LD + j0 + 32] : 0.0;
              a2[q] = c2 ? A[i * LD + j0 + 64] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int i = i0 + q;
              if (i < N) {
                A[i * LD + j0] = fma(-m[q], u0, a0[q]);
                if (c1) A[i * LD + j0 + 32] = fma(-m[q], u1, a1[q]);
                if (c2) A[i * LD + j0 + 64] = fma(-m[q], u2, a2[q]);
              }
            }
          }
        }
      } else {
        // wide systems (max_points > 64): any number of column slots per lane
        const int NC = N + 1;
        for (int j = k + 1 + lane; j < NC; j += 32) {
          const double u = A[k * LD + j];
          for (int i = k + 1; i < N; ++i) A[i * LD + j] = fma(-mult[i], u, A[i * LD + j]);
        }
      }
      __syncwarp();
    }
    if (singular) {
      if (lane == 0) {
        a.z[t] = qnan;
        a.sigma[t] = qnan;
        a.status[t] = SKG_KRIGE_SINGULAR;
      }
      __syncwarp();
      continue;
    }
    // back substitution; solution overwrites the rhs column A[i*LD + N]
    for (int i = N - 1; i >= 0; --i) {
      double acc = 0.0;
      for (int j = i + 1 + lane; j < N; j += 32) acc = fma(A[i * LD + j], A[j * LD + N], acc);
      acc = warp_sum(acc);
      if (lane == 0) A[i * LD + N] = (A[i * LD + N] - acc) / A[i * LD + i];
      __syncwarp();
    }
    double sz = 0.0, ss = 0.0;
    for (int i = lane; i < nn; i += 32) {
      const double w = A[i * LD + N];
      sz = fma(w, nval[i], sz);   // Kriging.py:647
      ss = fma(w, rhs[i], ss);    // Kriging.py:644
    }
    sz = warp_sum(sz);
    ss = warp_sum(ss);
    if (lane == 0) {
      a.z[t] = sz;
      a.sigma[t] = ss + A[nn * LD + N];
      a.status[t] = SKG_KRIGE_OK;
    }
    __syncwarp();
  }
}

// ---------------------------------------------- covariance-form fast path ---
//
// The ordinary-kriging system  [G 1; 1' 0] [w; mu] = [g; 1]  with G_ij = gamma(d_ij),
// G_ii = 0 (Kriging.py:594-614) is rewritten with C = c_t 11' - G, c = c_t 1 - g,
// c_t = nugget + sill:   C w = c + mu 1,  1'w = 1.  C is the covariance matrix of the
// (valid) model, symmetric positive definite, so it factors as L L' WITHOUT pivoting in
// half the flops and half the memory of the LU of the indefinite system:
//     y1 = L^-1 c,  y2 = L^-1 1,  mu = (1 - y2.y1) / (y2.y2),  w = L^-T (y1 + mu y2).
// Same weights, Lagrange multiplier, z and sigma as the reference's inv(A).dot(b) up to
// cond * eps.  If a pivot is not positive in fp64 (invalid / degenerate model) the
// target is marked KRIGE_RETRY and the partial-pivot LU kernel solves the original system.
// Packed storage: row i of L starts at i(i+1)/2; rows n and n+1 hold c and 1 and are
// carried through the elimination, which performs the two forward solves for free.

__device__ __forceinline__ int prow(int i, int n) {  // offset of packed row i (rows n, n+1: n entries)
  return i <= n ? i * (i + 1) / 2 : n * (n + 1) / 2 + n;
}

template <int DIM, int NQ>  // NQ: neighbour slots per lane (max_points <= 32 * NQ)
__global__ void __maxnreg__(NQ == 2 ? 104 : 128)
krige_chol_kernel(const __grid_constant__ KrigeArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const int maxp = a.max_points;
  // per-warp layout: region | values | gamma(d_0i) | x [maxp] (its space first holds the neighbour
  // ids) | nbr coords [DIM][maxp] (only without the sample-covariance cache)
  const bool cached = a.sample_cov != nullptr;
  const int per_warp = a.region_bytes + 3 * maxp * 8 + (cached ? 0 : DIM * maxp * 8);
  unsigned char* base = smem_raw + (size_t)warp * per_warp;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(base);
  int* ids = reinterpret_cast<int*>(base + (size_t)a.cap * 8);
  double* L = reinterpret_cast<double*>(base);
  double* nval = reinterpret_cast<double*>(base + a.region_bytes);
  double* gvec = nval + maxp;
  double* xv = gvec + maxp;
  int* nid = reinterpret_cast<int*>(xv);  // neighbour sample ids: dead before x is written
  double* ncoord = xv + maxp;

  const double* __restrict__ X = a.coords;
  const double* __restrict__ Y = a.coords + a.n;
  const double* __restrict__ W = a.coords + 2 * a.n;
  const double qnan = __longlong_as_double(NAN_BITS);
  const double ct = a.model.b + a.model.c0;

  for (int64_t t = a.t0 + (int64_t)blockIdx.x * wpb + warp; t < a.t1; t += (int64_t)gridDim.x * wpb) {
    const double tx = a.targets[t];
    const double ty = a.targets[a.m + t];
    const double tw = DIM == 3 ? a.targets[2 * a.m + t] : 0.0;
    const int nn = find_neighbours<DIM>(a, tx, ty, tw, a.exclude ? a.exclude[t] : -1, keys, ids, lane);
    __syncwarp();
    if (nn < a.min_points || nn < 1) {  // Kriging.py:579-580 LessPointsError
      if (lane == 0) {
        a.z[t] = qnan;
        a.sigma[t] = qnan;
        a.status[t] = SKG_KRIGE_LESS_POINTS;
      }
      __syncwarp();
      continue;
    }
    double kd[NQ];
    int kid[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int e = lane + 32 * q;
      kd[q] = e < nn ? __longlong_as_double((long long)keys[e]) : 0.0;
      kid[q] = e < nn ? ids[e] : 0;
    }
    __syncwarp();
    const int n = nn;
    const int rowc = prow(n, n), rowo = prow(n + 1, n);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int e = lane + 32 * q;
      if (e < n) {
        if (!cached) {
          ncoord[e] = X[kid[q]];
          ncoord[maxp + e] = Y[kid[q]];
          if (DIM == 3) ncoord[2 * maxp + e] = W[kid[q]];
        }
        nval[e] = a.values[kid[q]];
        const double gv = gamma_model(kd[q], a.model);  // Kriging.py:611-613
        gvec[e] = gv;
        nid[e] = kid[q];
        L[prow(e, n) + e] = ct;  // C_ii = c_t - 0 (zero diagonal of the gamma matrix, :600)
        L[rowc + e] = ct - gv;
        L[rowo + e] = 1.0;
      }
    }
    __syncwarp();
    const int npairs = n * (n - 1) / 2;
    if (a.sample_cov) {
      // the sample-sample entries do not depend on the target: gather them from the cached
      // covariance matrix (the reference caches the N x N sample distances the same way,
      // MetricSpace.py:158-187) instead of 2016 distance + model evaluations per target.  Rows
      // of adjacent targets' neighbours stay L2-resident.
      const double* __restrict__ cov = a.sample_cov;
      // four rows per step: up to four independent L2 loads in flight per lane
      for (int j = 1; j < n; j += 4) {
        int jj[4], oo[4];
        const double* rr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          jj[u] = j + u < n ? j + u : n - 1;
          rr[u] = cov + (int64_t)nid[jj[u]] * a.n;
          oo[u] = prow(jj[u], n);
        }
        for (int i = lane; i < jj[3]; i += 32) {
          const int id = nid[i];
          double cv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) cv[u] = rr[u][id];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (i < jj[u] && j + u < n) L[oo[u] + i] = cv[u];
        }
      }
    } else
    for (int p0 = lane; p0 < npairs; p0 += 64) {
      int pi[2], pj[2];
      double cv[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int p = min(p0 + 32 * q, npairs - 1);
        // p -> (j, i), i < j: packed strictly-lower index p = j(j-1)/2 + i
        int j = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)p)) * 0.5f);  // exact after the fix-ups
        while (j * (j - 1) / 2 > p) --j;
        while ((j + 1) * j / 2 <= p) ++j;
        pj[q] = j;
        pi[q] = p - j * (j - 1) / 2;
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int i = pi[q], j = pj[q];
        const double dx = ncoord[i] - ncoord[j];
        const double dy = ncoord[maxp + i] - ncoord[maxp + j];
        double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        if (DIM == 3) {
          const double dw = ncoord[2 * maxp + i] - ncoord[2 * maxp + j];
          s = __dadd_rn(s, __dmul_rn(dw, dw));
        }
        cv[q] = ct - gamma_matrix(sqrt(s), a.model);  // Kriging.py:594, 652-679
      }
#pragma unroll
      for (int q = 0; q < 2; ++q)
        if (p0 + 32 * q < npairs) L[prow(pj[q], n) + pi[q]] = cv[q];
    }
    __syncwarp();

    // ---- blocked Cholesky, rows n and n+1 (rhs) carried along --------------
    // Right-looking with 8-column panels.  The panel is factored with plain fp64 FMAs
    // (lanes over rows); the trailing update  A22 -= P P'  runs on the FP64 tensor cores
    // (mma.sync m8n8k4, 8x8 output tiles, two k=4 steps per panel), which replaces ~30
    // warp-FMAs per tile by two instructions.
    bool bad = false;
    const int R = n + 2;                 // rows incl. the two rhs rows
    const int ncb = (n + 7) >> 3;        // column blocks
    const int nrb = (R + 7) >> 3;        // row blocks
    const int grp = lane >> 2, tig = lane & 3;
    for (int bk = 0; bk < ncb && !bad; ++bk) {
      const int kb = bk << 3;
      const int ke = min(kb + 8, n);
      // Tall panels (many rows below the 8 x 8 diagonal block): factor the block alone, then solve
      // every row below it with its 8 panel entries in registers.  Short panels (small systems, the
      // last panels): the plain right-looking column loop over all rows is cheaper.
      const bool tall = R - ke >= 20;
      for (int k = kb; k < ke; ++k) {
        const int ok = k * (k + 1) / 2;
        const double akk = L[ok + k];
        __syncwarp();  // every lane holds the pivot before lane 0 overwrites it below
        if (!(akk > 0.0)) { bad = true; break; }
        const double rd = rsqrt(akk);
        const int rend = tall ? ke : R;           // rows this loop handles
        for (int i = k + 1 + lane; i < rend; i += 32) {
          const int o = prow(i, n) + k;
          L[o] *= rd;
        }
        if (lane == 0) L[ok + k] = akk * rd;
        __syncwarp();
        // remaining panel columns c in (k, ke): L[i][c] -= L[i][k] * L[c][k]
        if (k + 1 < ke) {
          for (int i = k + 1 + lane; i < rend; i += 32) {
            const int o = prow(i, n);
            const double li = L[o + k];
            const int cmax = i < n ? min(i, ke - 1) : ke - 1;
            int oc = (k + 1) * (k + 2) / 2 + k;   // offset of L[c][k], advanced by c + 1 per column
            for (int c = k + 1; c <= cmax; ++c) {
              L[o + c] = fma(-li, L[oc], L[o + c]);
              oc += c + 1;
            }
          }
        }
        __syncwarp();
      }
      if (bad) break;
      if (tall) {
        // rows below the block (the two rhs rows included):  a[c] = (a[c] - sum_{m<c} a[m] D[c][m]) / D[c][c]
        // with the row's 8 panel entries in registers and the block D read as warp broadcasts:
        // 2.5 instructions per FMA instead of 12
        double rdl = 1.0;
        if (lane < 8) rdl = 1.0 / L[(kb + lane) * (kb + lane + 1) / 2 + kb + lane];  // tall => full 8-column panel
        double rdc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) rdc[c] = __shfl_sync(0xffffffffu, rdl, c);
        for (int i = ke + lane; i < R; i += 32) {
          const int o = prow(i, n) + kb;
          double a[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) a[c] = L[o + c];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int od = (kb + c) * (kb + c + 1) / 2 + kb;
            double acc = a[c];
#pragma unroll
            for (int m = 0; m < c; ++m) acc = fma(-a[m], L[od + m], acc);
            a[c] = acc * rdc[c];
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) L[o + c] = a[c];
        }
        __syncwarp();
      }
      if (bad) break;
      // trailing update on the tensor cores: tiles (I, J), J > bk, I >= J
      for (int I = bk + 1; I < nrb; ++I) {
        const int r = (I << 3) + grp;           // this lane's row of the A and C fragments
        const bool rv = r < R;
        const int orow = prow(rv ? r : 0, n);
        double a0 = 0.0, a1 = 0.0;              // -P[r][kb + tig], -P[r][kb + 4 + tig]
        if (rv && kb + tig < ke) a0 = -L[orow + kb + tig];
        if (rv && kb + 4 + tig < ke) a1 = -L[orow + kb + 4 + tig];
        const int jhi = min(I, ncb - 1);
        for (int J = bk + 1; J <= jhi; ++J) {
          const int rj = (J << 3) + grp;        // row of P feeding column rj of the tile
          const bool jv = rj < n;
          const int oj = prow(jv ? rj : 0, n);
          double b0 = 0.0, b1 = 0.0;
          if (jv && kb + tig < ke) b0 = L[oj + kb + tig];
          if (jv && kb + 4 + tig < ke) b1 = L[oj + kb + 4 + tig];
          const int c0 = (J << 3) + 2 * tig, c1 = c0 + 1;
          const bool v0 = rv && c0 < n && (c0 <= r || r >= n);
          const bool v1 = rv && c1 < n && (c1 <= r || r >= n);
          double d0 = v0 ? L[orow + c0] : 0.0;
          double d1 = v1 ? L[orow + c1] : 0.0;
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(d0), "+d"(d1) : "d"(a0), "d"(b0));
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(d0), "+d"(d1) : "d"(a1), "d"(b1));
          if (v0) L[orow + c0] = d0;
          if (v1) L[orow + c1] = d1;
        }
      }
      __syncwarp();
    }
    if (bad) {
      if (lane == 0) a.status[t] = KRIGE_RETRY;
      __syncwarp();
      continue;
    }
    // y1 = row n, y2 = row n+1;  mu = (1 - y2.y1) / (y2.y2)
    double s11 = 0.0, s12 = 0.0;
    for (int i = lane; i < n; i += 32) {
      const double y1 = L[rowc + i], y2 = L[rowo + i];
      s12 = fma(y2, y1, s12);
      s11 = fma(y2, y2, s11);
    }
    s12 = warp_sum(s12);
    s11 = warp_sum(s11);
    s12 = __shfl_sync(0xffffffffu, s12, 0);
    s11 = __shfl_sync(0xffffffffu, s11, 0);
    const double mu = (1.0 - s12) / s11;
    // back substitution L' w = y1 + mu y2 in axpy form: y lives in registers (lane owns
    // entries lane and lane+32), row i of L is read contiguously, no reductions
    if (NQ == 2) {
      double yA = lane < n ? fma(mu, L[rowo + lane], L[rowc + lane]) : 0.0;
      double yB = lane + 32 < n ? fma(mu, L[rowo + lane + 32], L[rowc + lane + 32]) : 0.0;
      for (int i = n - 1; i >= 0; --i) {
        const int oi = i * (i + 1) / 2;
        const double yi = __shfl_sync(0xffffffffu, i < 32 ? yA : yB, i & 31);
        const double wi = yi / L[oi + i];
        if (lane == (i & 31)) { if (i < 32) yA = wi; else yB = wi; }
        if (lane < i) yA = fma(-L[oi + lane], wi, yA);
        if (lane + 32 < i) yB = fma(-L[oi + lane + 32], wi, yB);
      }
      if (lane < n) xv[lane] = yA;
      if (lane + 32 < n) xv[lane + 32] = yB;
    } else {
      double y[NQ];
#pragma unroll
      for (int q = 0; q < NQ; ++q)
        y[q] = lane + 32 * q < n ? fma(mu, L[rowo + lane + 32 * q], L[rowc + lane + 32 * q]) : 0.0;
      for (int i = n - 1; i >= 0; --i) {
        const int oi = i * (i + 1) / 2;
        double ysrc = y[0];
#pragma unroll
        for (int q = 1; q < NQ; ++q) ysrc = (i >> 5) == q ? y[q] : ysrc;
        const double yi = __shfl_sync(0xffffffffu, ysrc, i & 31);
        const double wi = yi / L[oi + i];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          if (lane == (i & 31) && (i >> 5) == q) y[q] = wi;
          if (lane + 32 * q < i) y[q] = fma(-L[oi + lane + 32 * q], wi, y[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < NQ; ++q)
        if (lane + 32 * q < n) xv[lane + 32 * q] = y[q];
    }
    __syncwarp();
    double sz = 0.0, ss = 0.0;
    for (int i = lane; i < n; i += 32) {
      const double w = xv[i];
      sz = fma(w, nval[i], sz);   // Kriging.py:647
      ss = fma(w, gvec[i], ss);   // Kriging.py:644
    }
    sz = warp_sum(sz);
    ss = warp_sum(ss);
    if (lane == 0) {
      a.z[t] = sz;
      a.sigma[t] = ss + mu;
      a.status[t] = SKG_KRIGE_OK;
    }
    __syncwarp();
  }
}

// c_t - gamma_matrix(d_ij) for every sample pair, with the arithmetic of the per-target build
// (same device functions, same operand order; s is symmetric in (i, j) bit for bit)
template <int DIM>
__global__ void __launch_bounds__(256)
krige_sample_cov_kernel(const double* __restrict__ coords, int64_t n, const __grid_constant__ ModelParams model,
                        double* __restrict__ cov) {
  const double ct = model.b + model.c0;
  const double* __restrict__ X = coords;
  const double* __restrict__ Y = coords + n;
  const double* __restrict__ W = coords + 2 * n;
  const int64_t total = n * n;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total;
       k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = k / n, j = k - i * n;
    if (i == j) { cov[k] = ct; continue; }
    const double dx = X[i] - X[j];
    const double dy = Y[i] - Y[j];
    double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    if (DIM == 3) {
      const double dw = W[i] - W[j];
      s = __dadd_rn(s, __dmul_rn(dw, dw));
    }
    cov[k] = ct - gamma_matrix(sqrt(s), model);
  }
}

static int parse_grid(const double* grid, int dim, GridParams* g) {
  if (!grid) { set_error("grid NULL"); return SKG_E_BADARG; }
  g->ox = grid[0]; g->oy = grid[1]; g->oz = grid[2]; g->h = grid[3];
  g->nx = (int)grid[4]; g->ny = (int)grid[5]; g->nz = dim == 3 ? (int)grid[6] : 1;
  if (!(g->h > 0.0) || g->nx < 1 || g->ny < 1 || g->nz < 1 ||
      (double)g->nx * g->ny * g->nz > 2.0e9) {
    set_error("bad grid: h=%g nx=%d ny=%d nz=%d", g->h, g->nx, g->ny, g->nz);
    return SKG_E_BADARG;
  }
  return 0;
}

static void krige_layout(int max_points, int dim, int* cap, int* region, int* per_warp) {
  const int N = max_points + 1;
  int reg = N * ((N + 1) | 1) * 8;
  reg = std::max(reg, 1024 * 12);
  reg = (reg + 15) & ~15;
  *cap = reg / 12;
  *region = reg;
  *per_warp = reg + (dim + 1) * max_points * 8 + 2 * (max_points + 1) * 8;
}

static void krige_chol_layout(int max_points, int dim, bool cached, int* cap, int* region, int* per_warp) {
  const int n = max_points;
  int reg = (n * (n + 1) / 2 + 2 * n) * 8;
  reg = std::max(reg, 1024 * 12);
  reg = (reg + 15) & ~15;
  *cap = reg / 12;
  *region = reg;
  *per_warp = reg + 3 * n * 8 + (cached ? 0 : dim * n * 8);
}

// warps per CTA that maximise resident warps per SM for a given per-warp footprint
static int best_wpb(int per_warp, int smem_budget, int max_w) {
  // most resident warps per SM; among equals the SMALLEST CTA (measured: 4-warp CTAs run the small
  // systems 35 % faster than one 16-warp CTA per SM at the same occupancy)
  int best = 0, best_w = 0;
  for (int w = 1; w <= max_w; ++w) {
    if ((size_t)w * per_warp > (size_t)smem_budget) break;
    const int ctas = std::min(32, (smem_budget + 1024) / (w * per_warp + 1024));
    const int resident = std::min(64, ctas * w);
    if (resident > best && !(w == 1 && max_w > 1)) { best = resident; best_w = w; }
  }
  if (!best_w && (size_t)per_warp <= (size_t)smem_budget) best_w = 1;
  return best_w;
}

}  // namespace skg

using namespace skg;

extern "C" {

int skg_krige_build_grid(const double* coords, int64_t n, int dim, const double* grid,
                         int32_t* cell_start, int32_t* order, void* scratch,
                         int64_t scratch_bytes, void* stream) {
  if (!coords || n < 1 || !cell_start || !order) { set_error("NULL argument or n < 1"); return SKG_E_BADARG; }
  if (dim != 2 && dim != 3) { set_error("dim=%d unsupported", dim); return SKG_E_UNSUPPORTED; }
  if (n > 2000000000ll) { set_error("n too large"); return SKG_E_UNSUPPORTED; }
  GridParams g;
  int rc = parse_grid(grid, dim, &g);
  if (rc) return rc;
  const int64_t ncell = (int64_t)g.nx * g.ny * g.nz;
  if (!scratch || scratch_bytes < (ncell + n) * 4) {
    set_error("scratch too small: need %lld bytes", (long long)((ncell + n) * 4));
    return SKG_E_SCRATCH;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int* counts = reinterpret_cast<int*>(scratch);
  int* cellid = counts + ncell;
  SKG_CUDA(cudaMemsetAsync(counts, 0, ncell * 4, st));
  const int threads = 256;
  const unsigned nblk = (unsigned)((n + threads - 1) / threads);
  if (dim == 3) grid_count_kernel<3><<<nblk, threads, 0, st>>>(coords, n, g, cellid, counts);
  else grid_count_kernel<2><<<nblk, threads, 0, st>>>(coords, n, g, cellid, counts);
  grid_scan_kernel<<<1, 1024, 0, st>>>(counts, (int)ncell, cell_start);
  grid_scatter_kernel<<<nblk, threads, 0, st>>>(cellid, n, counts, order);
  grid_sort_cells_kernel<<<(unsigned)((ncell + threads - 1) / threads), threads, 0, st>>>(
      cell_start, (int)ncell, order);
  SKG_CUDA(cudaGetLastError());
  return 0;
}

static int fill_model(int model, const double* model_params, const double* matrix_lut, int lut_size,
                      ModelParams* mp) {
  if (model < SKG_MODEL_SPHERICAL || model > SKG_MODEL_MATERN) { set_error("unknown model id %d", model); return SKG_E_BADARG; }
  if ((matrix_lut != nullptr) != (lut_size > 0)) { set_error("matrix_lut and lut_size must come together"); return SKG_E_BADARG; }
  std::memset(mp, 0, sizeof(*mp));
  mp->model = model;
  mp->r = model_params[0]; mp->c0 = model_params[1];
  mp->s = model_params[2]; mp->b = model_params[3];
  switch (model) {
    case SKG_MODEL_SPHERICAL: mp->a = mp->r / 1.0; break;
    case SKG_MODEL_EXPONENTIAL: mp->a = mp->r / 3.0; break;
    case SKG_MODEL_GAUSSIAN: mp->a = mp->r / 2.0; break;
    case SKG_MODEL_CUBIC: mp->a = mp->r / 1.0; break;
    case SKG_MODEL_MATERN:
      mp->a = mp->r / 2.0;
      mp->gnorm = 2.0 / std::tgamma(mp->s);
      break;
    default: mp->a = mp->r / std::pow(3.0, 1.0 / mp->s); break;
  }
  mp->lut = matrix_lut;
  mp->lut_size = lut_size;
  mp->lut_range = model_params[4];
  mp->lut_out = model_params[5];
  return 0;
}

int skg_krige_sample_cov(const double* coords, int64_t n, int dim, int model, const double* model_params,
                         const double* matrix_lut, int lut_size, double* cov, void* stream) {
  if (!coords || !model_params || !cov || n < 1) { set_error("NULL argument or n < 1"); return SKG_E_BADARG; }
  if (dim != 2 && dim != 3) { set_error("dim=%d unsupported", dim); return SKG_E_UNSUPPORTED; }
  ModelParams mp;
  int rc = fill_model(model, model_params, matrix_lut, lut_size, &mp);
  if (rc) return rc;
  const int64_t blocks = std::min<int64_t>((n * n + 255) / 256, (int64_t)sm_count() * 32);
  if (dim == 3)
    krige_sample_cov_kernel<3><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(coords, n, mp, cov);
  else
    krige_sample_cov_kernel<2><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(coords, n, mp, cov);
  SKG_CUDA(cudaGetLastError());
  return 0;
}

int64_t skg_krige_scratch_bytes(int max_points) {
  (void)max_points;
  return 16;  // the warp-per-target kernel keeps everything in shared memory
}

int skg_krige(const double* coords, const double* values, int64_t n, int dim, const double* grid,
              const int32_t* cell_start, const int32_t* order, const double* targets, int64_t m,
              int model, const double* model_params, const double* matrix_lut, int lut_size,
              double radius, int min_points, int max_points, const int32_t* exclude,
              const double* sample_cov, double* z, double* sigma, int32_t* status, void* scratch,
              int64_t scratch_bytes, int64_t target_begin, int64_t target_end, void* stream) {
  (void)scratch; (void)scratch_bytes;
  if (!coords || !values || !cell_start || !order || !targets || !model_params || !z || !sigma || !status) {
    set_error("NULL argument");
    return SKG_E_BADARG;
  }
  if (dim != 2 && dim != 3) { set_error("dim=%d unsupported", dim); return SKG_E_UNSUPPORTED; }
  if (n < 1 || m < 0 || target_begin < 0 || target_end < target_begin || target_end > m) {
    set_error("bad sizes / target range");
    return SKG_E_BADARG;
  }
  if (max_points < 1 || max_points > SKG_KRIGE_MAX_POINTS) {
    set_error("max_points=%d outside [1, %d]", max_points, SKG_KRIGE_MAX_POINTS);
    return SKG_E_UNSUPPORTED;
  }
  if (min_points < 0 || min_points > max_points) { set_error("min_points outside [0, max_points]"); return SKG_E_BADARG; }
  if (!(radius >= 0.0)) { set_error("radius must be >= 0"); return SKG_E_BADARG; }
  KrigeArgs a;
  std::memset(&a, 0, sizeof(a));
  int rc = parse_grid(grid, dim, &a.grid);
  if (rc) return rc;
  rc = fill_model(model, model_params, matrix_lut, lut_size, &a.model);
  if (rc) return rc;
  a.coords = coords; a.values = values; a.n = n;
  a.cell_start = cell_start; a.order = order; a.targets = targets; a.m = m;
  a.sample_cov = sample_cov;
  a.radius = radius; a.min_points = min_points; a.max_points = max_points;
  a.z = z; a.sigma = sigma; a.status = status; a.t0 = target_begin; a.t1 = target_end;
  a.exclude = exclude;
  if (target_end == target_begin) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int smem_budget = 227 * 1024;  // the opt-in maximum of dynamic shared memory per CTA
  const int64_t ntarget = target_end - target_begin;
  // pass 1: covariance-form Cholesky (fast path)
  {
    KrigeArgs c = a;
    int per_warp = 0;
    krige_chol_layout(max_points, dim, sample_cov != nullptr, &c.cap, &c.region_bytes, &per_warp);
    const int wpb = best_wpb(per_warp, smem_budget, 16);
    if (wpb < 1) { set_error("max_points=%d does not fit shared memory", max_points); return SKG_E_UNSUPPORTED; }
    const size_t smem = (size_t)wpb * per_warp;
    auto kern = max_points <= 64 ? (dim == 3 ? krige_chol_kernel<3, 2> : krige_chol_kernel<2, 2>)
                                 : (dim == 3 ? krige_chol_kernel<3, 4> : krige_chol_kernel<2, 4>);
    SKG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    SKG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, wpb * 32, smem));
    occ = std::max(1, occ);
    const int64_t grid_blocks = std::min<int64_t>((ntarget + wpb - 1) / wpb, (int64_t)sm_count() * occ);
    kern<<<(unsigned)grid_blocks, wpb * 32, smem, st>>>(c);
    SKG_CUDA(cudaGetLastError());
  }
  // pass 2: partial-pivot LU of the original system for targets marked KRIGE_RETRY
  {
    int per_warp = 0;
    krige_layout(max_points, dim, &a.cap, &a.region_bytes, &per_warp);
    a.retry_only = 1;
    const int wpb = best_wpb(per_warp, smem_budget, 8);
    if (wpb < 1) { set_error("max_points=%d does not fit shared memory", max_points); return SKG_E_UNSUPPORTED; }
    const size_t smem = (size_t)wpb * per_warp;
    auto kern = max_points <= 64 ? (dim == 3 ? krige_kernel<3, 2> : krige_kernel<2, 2>)
                                 : (dim == 3 ? krige_kernel<3, 4> : krige_kernel<2, 4>);
    SKG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    SKG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, wpb * 32, smem));
    occ = std::max(1, occ);
    const int64_t grid_blocks = std::min<int64_t>((ntarget + wpb - 1) / wpb, (int64_t)sm_count() * occ);
    kern<<<(unsigned)grid_blocks, wpb * 32, smem, st>>>(a);
    SKG_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // extern "C"
