// Batched fused pass: ONE sweep over the point pairs accumulates the estimator sums of
// BV value vectors that share the coordinates (Monte-Carlo uncertainty propagation runs hundreds
// of variograms on one MetricSpace, util/uncertainty.py:119-171; the cross_variograms grid,
// util/cross_variogram.py:61-87).  Distance, lag lookup and the pair count are paid once per
// pair; each extra vector costs one subtraction, one FMA and its share of a 128-bit
// shared-memory read-modify-write.
//
// Same structure as pair_hist_kernel (pair_sweep.cuh): 64 threads x 4 "i" points, j chunk in
// static shared memory, private per-thread records, software-pipelined lag lookup, work items
// from the culled tile list, fixed-order reductions (deterministic).
#include <algorithm>
#include <cstring>

#include "pair_sweep.cuh"

namespace skg {

constexpr int BV = 4;  // value vectors per sweep

// per lag class: [BV/2][HB] 16-byte sum pairs, then [HB] u32 counts
constexpr int BATCH_LAG_BYTES = HB * (BV * 8 + 4);
constexpr int BATCH_CNT_OFF = HB * BV * 8;

struct BatchPartial {  // one per (CTA, lag class)
  unsigned long long count;
  double sum[BV];
};

struct BatchCols {
  long long off[BV];  // element offsets of the BV columns inside `values`
};

template <bool SQRT>
__device__ __forceinline__ void rmw_batch(unsigned addr, unsigned caddr, const double (&d)[BV], int pred) {
  double t[BV];
#pragma unroll
  for (int k = 0; k < BV; ++k) t[k] = SQRT ? sqrt(fabs(d[k])) : d[k];
  if (SQRT) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 a0, a1, b0, b1;\n\t.reg .f64 f0, f1, f2, f3;\n\t.reg .b32 c;\n\t"
        "setp.ne.s32 p, %6, 0;\n\t"
        "@p ld.shared.v2.b64 {a0, a1}, [%0];\n\t"
        "@p ld.shared.v2.b64 {b0, b1}, [%0+1024];\n\t"
        "@p ld.shared.u32 c, [%1];\n\t"
        "mov.b64 f0, a0;\n\tmov.b64 f1, a1;\n\tmov.b64 f2, b0;\n\tmov.b64 f3, b1;\n\t"
        "add.rn.f64 f0, f0, %2;\n\tadd.rn.f64 f1, f1, %3;\n\t"
        "add.rn.f64 f2, f2, %4;\n\tadd.rn.f64 f3, f3, %5;\n\t"
        "add.u32 c, c, 1;\n\t"
        "mov.b64 a0, f0;\n\tmov.b64 a1, f1;\n\tmov.b64 b0, f2;\n\tmov.b64 b1, f3;\n\t"
        "@p st.shared.v2.b64 [%0], {a0, a1};\n\t"
        "@p st.shared.v2.b64 [%0+1024], {b0, b1};\n\t"
        "@p st.shared.u32 [%1], c;\n\t}"
        :: "r"(addr), "r"(caddr), "d"(t[0]), "d"(t[1]), "d"(t[2]), "d"(t[3]), "r"(pred));
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 a0, a1, b0, b1;\n\t.reg .f64 f0, f1, f2, f3;\n\t.reg .b32 c;\n\t"
        "setp.ne.s32 p, %6, 0;\n\t"
        "@p ld.shared.v2.b64 {a0, a1}, [%0];\n\t"
        "@p ld.shared.v2.b64 {b0, b1}, [%0+1024];\n\t"
        "@p ld.shared.u32 c, [%1];\n\t"
        "mov.b64 f0, a0;\n\tmov.b64 f1, a1;\n\tmov.b64 f2, b0;\n\tmov.b64 f3, b1;\n\t"
        "fma.rn.f64 f0, %2, %2, f0;\n\tfma.rn.f64 f1, %3, %3, f1;\n\t"
        "fma.rn.f64 f2, %4, %4, f2;\n\tfma.rn.f64 f3, %5, %5, f3;\n\t"
        "add.u32 c, c, 1;\n\t"
        "mov.b64 a0, f0;\n\tmov.b64 a1, f1;\n\tmov.b64 b0, f2;\n\tmov.b64 b1, f3;\n\t"
        "@p st.shared.v2.b64 [%0], {a0, a1};\n\t"
        "@p st.shared.v2.b64 [%0+1024], {b0, b1};\n\t"
        "@p st.shared.u32 [%1], c;\n\t}"
        :: "r"(addr), "r"(caddr), "d"(t[0]), "d"(t[1]), "d"(t[2]), "d"(t[3]), "r"(pred));
  }
}
static_assert(BV == 4 && HB * 16 == 1024, "rmw_batch hard-codes the record layout");

template <int DIM, int DIR, bool SQRT, int EMAX>
__global__ void __launch_bounds__(HB)
pair_hist_batch_kernel(const double* __restrict__ coords, const double* __restrict__ values,
                       const __grid_constant__ BatchCols cols, int64_t n,
                       const __grid_constant__ BinTable tab, const __grid_constant__ DirParams dp,
                       BatchPartial* __restrict__ partial, const unsigned int* __restrict__ tile_list,
                       const unsigned long long* __restrict__ tile_count,
                       const double* __restrict__ bbox32, double wlo, double whi, int js) {
  __shared__ double2 sXY[TILE + 2];
  __shared__ double sW[DIM == 3 ? TILE + 2 : 1];
  __shared__ __align__(16) double sZ[(TILE + 2) * BV];  // j values, BV per point
  __shared__ uint64_t sThr[THR_MAX];
  __shared__ uint8_t sLut[LUT_MAX];
  extern __shared__ __align__(16) unsigned char acc_raw[];

  const int nb = tab.nb;
  const int tid = threadIdx.x;
  const double qnan = __longlong_as_double(NAN_BITS);
  for (int k = tid; k < THR_MAX; k += HB) sThr[k] = tab.thr[k];
  for (int k = tid; k < tab.ncell; k += HB) sLut[k] = tab.lut[k];
  {
    unsigned int* z = reinterpret_cast<unsigned int*>(acc_raw);
    for (int k = tid; k < nb * (BATCH_LAG_BYTES / 4); k += HB) z[k] = 0u;
  }
  __syncthreads();

  const int shift = tab.shift, klo = tab.key0, khi = tab.key0 + tab.ncell - 1, emax = tab.emax;
  const uint8_t* lutb = sLut - klo;
  const unsigned aBase = (unsigned)__cvta_generic_to_shared(acc_raw);
  const unsigned aRec = aBase + tid * 16;
  const unsigned aCnt = aBase + BATCH_CNT_OFF + tid * 4;
  const int64_t nt = (n + TILE - 1) / TILE;
  const double* __restrict__ X = coords;
  const double* __restrict__ Y = coords + n;
  const double* __restrict__ W = coords + 2 * n;
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
    double xi[HR], yi[HR], wi[HR], zi[HR][BV];
#pragma unroll
    for (int r = 0; r < HR; ++r) {
      const int64_t i = (int64_t)I * TILE + r * HB + tid;
      const bool v = i < n;
      xi[r] = v ? X[i] : qnan;
      yi[r] = v ? Y[i] : qnan;
      wi[r] = (DIM == 3) ? (v ? W[i] : qnan) : 0.0;
#pragma unroll
      for (int k = 0; k < BV; ++k) zi[r][k] = v ? values[cols.off[k] + i] : 0.0;
    }
    __syncthreads();
    for (int k = tid; k < jcount + 2; k += HB) {
      const int64_t j = (int64_t)J * TILE + jbeg + k;
      const bool v = k < jcount;
      sXY[k] = make_double2(v ? X[j] : qnan, v ? Y[j] : qnan);
      if (DIM == 3) sW[k] = v ? W[j] : qnan;
#pragma unroll
      for (int q = 0; q < BV; ++q) sZ[k * BV + q] = v ? values[cols.off[q] + j] : 0.0;
    }
    __syncthreads();
    if (jcount == 0 || !wact) continue;

    auto body = [&](auto diag_tag) {
      constexpr bool DIAG = decltype(diag_tag)::value;
      auto front = [&](int jj, int (&g)[HR]) {
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
          const uint64_t sb = (uint64_t)__double_as_longlong(s);
          int gg = lag_of<EMAX>(sb, sThr, lutb, shift, klo, khi, emax);
          if (DIAG) gg = (jbeg + jj <= r * HB + tid) ? nb : gg;
          if (DIR) {
            if (gg < nb && !dir_mask_t<DIR == 2>(dx, dy, s, dp)) gg = nb;
          }
          g[r] = gg;
        }
      };
      auto accumulate = [&](int jj, const int (&g)[HR]) {
        const double2 za = *reinterpret_cast<const double2*>(&sZ[jj * BV]);
        const double2 zb = *reinterpret_cast<const double2*>(&sZ[jj * BV + 2]);
#pragma unroll
        for (int r = 0; r < HR; ++r) {
          const int in = g[r] < nb ? 1 : 0;
          const double d[BV] = {zi[r][0] - za.x, zi[r][1] - za.y, zi[r][2] - zb.x, zi[r][3] - zb.y};
          const unsigned off = (unsigned)g[r] * (unsigned)BATCH_LAG_BYTES;
          rmw_batch<SQRT>(aRec + off, aCnt + off, d, in);
        }
      };
      int gA[HR], gB[HR];
      front(0, gA);
      int jj = 1;
      for (; jj + 1 <= jcount; jj += 2) {
        front(jj, gB);
        accumulate(jj - 1, gA);
        front(jj + 1, gA);
        accumulate(jj, gB);
      }
      if (jj <= jcount) {
        front(jj, gB);
        accumulate(jj - 1, gA);
      }
    };
    if (I == J) body(std::true_type{}); else body(std::false_type{});
  }
  __syncthreads();

  BatchPartial* out = partial + (int64_t)blockIdx.x * nb;
  const int warp = tid >> 5, lane = tid & 31;
  for (int gI = warp; gI < nb; gI += HB / 32) {
    const unsigned char* lag = acc_raw + (size_t)gI * BATCH_LAG_BYTES;
    unsigned long long c = 0;
    double a[BV] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < HB / 32; ++k) {
      const int th = lane + 32 * k;
      c += *reinterpret_cast<const unsigned int*>(lag + BATCH_CNT_OFF + th * 4);
#pragma unroll
      for (int q = 0; q < BV; ++q)
        a[q] += *reinterpret_cast<const double*>(lag + (q >> 1) * (HB * 16) + th * 16 + (q & 1) * 8);
    }
    c = warp_sum(c);
#pragma unroll
    for (int q = 0; q < BV; ++q) a[q] = warp_sum(a[q]);
    if (lane == 0) {
      out[gI].count = c;
#pragma unroll
      for (int q = 0; q < BV; ++q) out[gI].sum[q] = a[q];
    }
  }
}

// one warp per (lag class, column): fixed-order fold of the per-CTA partials
__global__ void __launch_bounds__(32)
pair_batch_reduce_kernel(const BatchPartial* __restrict__ partial, int nblocks, int nb, int ncols,
                         unsigned long long* __restrict__ count /* or NULL */,
                         double* __restrict__ sums /* [ncols][nb] */) {
  const int g = blockIdx.x, q = blockIdx.y, lane = threadIdx.x;
  unsigned long long c = 0;
  double a = 0.0;
  for (int k = lane; k < nblocks; k += 32) {
    const BatchPartial& p = partial[(int64_t)k * nb + g];
    if (q == 0) c += p.count;
    a += p.sum[q];
  }
  c = warp_sum(c);
  a = warp_sum(a);
  if (lane == 0) {
    if (q == 0 && count) count[g] += c;
    if (q < ncols) sums[(int64_t)q * nb + g] += a;
  }
}

struct BatchLaunch {
  const double* coords; const double* values; BatchCols cols; int64_t n;
  const BinTable* tab; const DirParams* dp; BatchPartial* partial;
  const unsigned int* list; const unsigned long long* list_count; const double* bbox32;
  double wlo, whi; int max_blocks; int64_t ntiles; int js; size_t smem; cudaStream_t st;
  int* nblocks_out;
};

template <int DIM, int DIR, bool SQRT, int EMAX>
static int launch_batch(const BatchLaunch& L) {
  auto kern = pair_hist_batch_kernel<DIM, DIR, SQRT, EMAX>;
  static thread_local size_t cached_smem = (size_t)-1;
  static thread_local int cached_occ = 0;
  if (cached_smem != L.smem) {
    SKG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
    SKG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cached_occ, kern, HB, L.smem));
    cached_smem = L.smem;
  }
  int occ = cached_occ;
  if (occ < 1) { set_error("pair_hist_batch kernel does not fit (smem=%zu)", L.smem); return SKG_E_UNSUPPORTED; }
  occ = std::min(occ, MAX_GRID_PER_SM);
  int64_t grid = std::min<int64_t>((int64_t)sm_count() * occ, L.ntiles * L.js);
  grid = std::min<int64_t>(grid, L.max_blocks);
  *L.nblocks_out = (int)grid;
  kern<<<(unsigned)grid, HB, L.smem, L.st>>>(L.coords, L.values, L.cols, L.n, *L.tab, *L.dp, L.partial,
                                             L.list, L.list_count, L.bbox32, L.wlo, L.whi, L.js);
  SKG_CUDA(cudaGetLastError());
  return 0;
}

template <int DIM, int DIR>
static int batch_variant(bool sqrt_stat, int emax, const BatchLaunch& L) {
  if (sqrt_stat) {
    if (emax <= 1) return launch_batch<DIM, DIR, true, 1>(L);
    return launch_batch<DIM, DIR, true, 0>(L);
  }
  if (emax <= 1) return launch_batch<DIM, DIR, false, 1>(L);
  return launch_batch<DIM, DIR, false, 0>(L);
}

static int dispatch_batch(int dim, int dir, bool sqrt_stat, int emax, const BatchLaunch& L) {
  if (dim == 3) return batch_variant<3, 0>(sqrt_stat, emax, L);
  if (dir == 1) return batch_variant<2, 1>(sqrt_stat, emax, L);
  if (dir == 2) return batch_variant<2, 2>(sqrt_stat, emax, L);
  return batch_variant<2, 0>(sqrt_stat, emax, L);
}

}  // namespace skg

using namespace skg;

extern "C" {

int skg_pair_batch_max_lags(void) { return (180 * 1024) / BATCH_LAG_BYTES; }

int64_t skg_pair_batch_scratch_bytes(int64_t n, int n_lags) {
  if (n_lags < 1) n_lags = 1;
  if (n < 0) n = 0;
  return sweep_head_bytes(n) +
         (int64_t)sm_count() * MAX_GRID_PER_SM * n_lags * (int64_t)sizeof(BatchPartial);
}

int skg_pair_hist_batch(const double* coords, const double* values, int n_vectors, int64_t n, int dim,
                        const double* dir, const uint64_t* thr_bits, int n_lags, int stat,
                        uint64_t* count, double* sums, void* scratch, int64_t scratch_bytes,
                        int64_t tile_begin, int64_t tile_end, int64_t tile_stride, void* stream) {
  int rc = check_pair_args(coords, n, dim, tile_begin, tile_end, tile_stride);
  if (rc) return rc;
  if (!values || !thr_bits || !count || !sums || n_lags < 1 || n_vectors < 1) {
    set_error("values/thr_bits/count/sums NULL, n_lags < 1 or n_vectors < 1");
    return SKG_E_BADARG;
  }
  if (stat != SKG_STAT_SUM_SQ && stat != SKG_STAT_SUM_SQRT) {
    set_error("batched pass: stat must be SKG_STAT_SUM_SQ or SKG_STAT_SUM_SQRT");
    return SKG_E_BADARG;
  }
  if (n_lags > skg_pair_batch_max_lags()) {
    set_error("batched pass: n_lags=%d exceeds %d (private records); use skg_pair_hist per vector",
              n_lags, skg_pair_batch_max_lags());
    return SKG_E_UNSUPPORTED;
  }
  const int max_blocks = sm_count() * MAX_GRID_PER_SM;
  SweepScratch S;
  rc = carve_scratch(scratch, scratch_bytes, n, (int64_t)max_blocks * n_lags * (int64_t)sizeof(BatchPartial), &S);
  if (rc) return rc;
  BinTable tab;
  rc = build_bin_table(thr_bits, n_lags, &tab);
  if (rc) return rc;
  DirParams dp;
  rc = build_dir_params(dir, dim, &dp);
  if (rc) return rc;
  if (tile_end == tile_begin) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  CullWindows win;
  std::memset(&win, 0, sizeof(win));
  win.count = 1;
  win.lo[0] = 0.0;
  std::memcpy(&win.hi[0], &thr_bits[n_lags - 1], 8);
  const int64_t ntl = (tile_end - tile_begin + tile_stride - 1) / tile_stride;
  const int js = choose_js(ntl);
  cull_set_dir(&win, dp);
  rc = launch_cull(coords, n, dim, win, tile_begin, tile_end, tile_stride, js, S, st);  // once per call
  if (rc) return rc;
  const int use_dir = dp.model == SKG_DIR_NONE ? 0 : (dp.amb_exclude ? 1 : 2);
  BatchPartial* partial = reinterpret_cast<BatchPartial*>(S.tail);
  for (int v0 = 0; v0 < n_vectors; v0 += BV) {
    const int ncols = std::min(BV, n_vectors - v0);
    int nblocks = 0;
    BatchLaunch L{coords, values, {}, n, &tab, &dp, partial, S.list, S.counters, S.bbox32,
                  win.lo[0], win.hi[0], max_blocks, ntl, js, (size_t)n_lags * BATCH_LAG_BYTES, st, &nblocks};
    for (int q = 0; q < BV; ++q) L.cols.off[q] = (long long)(v0 + std::min(q, ncols - 1)) * n;
    rc = dispatch_batch(dim, use_dir, stat == SKG_STAT_SUM_SQRT, tab.emax, L);
    if (rc) return rc;
    pair_batch_reduce_kernel<<<dim3((unsigned)n_lags, BV), 32, 0, st>>>(
        partial, nblocks, n_lags, ncols, v0 == 0 ? reinterpret_cast<unsigned long long*>(count) : nullptr,
        sums + (int64_t)v0 * n_lags);
    SKG_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // extern "C"
