// Shared device/host helpers for the sm_100a pair-sweep and kriging kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "skg_b200.h"

namespace skg {

constexpr int TILE = SKG_TILE;  // points per tile edge
constexpr int BLOCK = 128;      // threads per CTA in the pair sweeps
constexpr int RI = TILE / BLOCK;  // i-points held in registers per thread
constexpr int LUT_MAX = 1024;
constexpr int THR_MAX = SKG_MAX_LAGS + 6;
constexpr uint64_t INF_BITS = 0x7FF0000000000000ull;
constexpr uint64_t NAN_BITS = 0x7FF8000000000000ull;

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int sm_count();

#define SKG_CUDA(call)                                  \
  do {                                                  \
    cudaError_t _e = (call);                            \
    if (_e != cudaSuccess) return skg::cuda_fail(_e, #call); \
  } while (0)

// Lag-class lookup table.  bin(s) = #{k : thr[k] <= bits(s)} is found as
// lut[cell(s)] plus at most `emax` integer compares; cell(s) is a monotone
// function of the high word of s (log-spaced cells), see build_bin_table().
struct BinTable {
  uint64_t thr[THR_MAX];  // thresholds (bit patterns), padded with UINT64_MAX
  uint8_t lut[LUT_MAX];
  int nb;     // number of lag classes
  int key0;   // (hi >> shift) of cell 0
  int shift;
  int ncell;
  int emax;
};

// Directional mask (DirectionalVariogram._triangle / _compass).
struct DirParams {
  int model;          // SKG_DIR_*
  int tol_all;        // radians(tol/2) >= pi/2: tolerance test always true
  int amb_exclude;    // 1: guard-band pairs fail the mask here (the caller resolves them, see
                      //    skg_pair_dir_ambiguous); 0: they take the literal formula on the device
  double cos_az, sin_az, tan_ht;
  double az_rad, half_tol_rad, half_bw;
};

// Per-segment key windows for the select passes.
struct SegTable {
  uint64_t lo_bits[SKG_MAX_LAGS + 1];
  uint64_t hi_bits[SKG_MAX_LAGS + 1];
  double scale[SKG_MAX_LAGS + 1];
};

int build_bin_table(const uint64_t* thr_bits, int n_lags, BinTable* out);
int build_dir_params(const double* dir, int dim, DirParams* out);

// ---------------------------------------------------------------- device ---

// DirectionalVariogram.py:348-413 + :659-714 / :749-782.
// Fast path: acute angle test through dot/cross products against the azimuth
// direction; pairs within a relative 2^-40 guard band of either decision
// boundary (and coincident points) take the literal acos/sin formula.
__device__ __forceinline__ bool dir_literal(double dx, double dy, double s, const DirParams& p) {
  const double PI = 3.141592653589793;
  double d = sqrt(s);
  double pos = acos(dx / d);  // NaN for coincident points -> every test false
  double ang = (dy >= 0.0) ? pos : -pos;
  double t = fabs(ang + p.az_rad);
  double a = t;
  a = (a > PI) ? a - PI : a;
  a = (a > PI / 2) ? PI - a : a;
  bool in_tol = a <= p.half_tol_rad;
  if (p.model == SKG_DIR_COMPASS) return in_tol;
  bool in_band = p.half_bw >= fabs(d * sin(t));
  return in_tol && in_band;
}

// returns the fast-path decision; *amb is set when the pair lies in the guard band
__device__ __forceinline__ bool dir_fast(double dx, double dy, const DirParams& p, bool* amb) {
  const double G = 9.094947017729282e-13;  // 2^-40
  double dot = fabs(dx * p.cos_az - dy * p.sin_az);
  double crs = fabs(dx * p.sin_az + dy * p.cos_az);
  double guard = G * (fabs(dx) + fabs(dy));
  bool ok = true, a = false;
  if (!p.tol_all) {
    double lim = p.tan_ht * dot;
    ok = crs <= lim;
    a |= fabs(crs - lim) <= guard * (1.0 + p.tan_ht);
  }
  if (p.model == SKG_DIR_TRIANGLE) {
    ok &= crs <= p.half_bw;
    a |= fabs(crs - p.half_bw) <= guard;
  }
  *amb = a;
  return ok;
}

// LITERAL = false: guard-band pairs fail (policy 1, resolved by the caller) - no
// transcendental code is even compiled into the sweep; LITERAL = true: policy 0.
template <bool LITERAL>
__device__ __forceinline__ bool dir_mask_t(double dx, double dy, double s, const DirParams& p) {
  if (!(s > 0.0)) return false;  // coincident points: 0/0 -> NaN angle -> every test false (:405)
  bool amb;
  bool ok = dir_fast(dx, dy, p, &amb);
  if (amb) ok = LITERAL ? dir_literal(dx, dy, s, p) : false;
  return ok;
}

__device__ __forceinline__ bool dir_mask(double dx, double dy, double s, const DirParams& p) {
  return p.amb_exclude ? dir_mask_t<false>(dx, dy, s, p) : dir_mask_t<true>(dx, dy, s, p);
}

// tile index -> (I, J), I <= J, row-major over the upper triangle of nt x nt.
__device__ __forceinline__ void tile_decode(int64_t t, int64_t nt, int& I, int& J) {
  double b = 2.0 * (double)nt + 1.0;
  int64_t i = (int64_t)((b - sqrt(b * b - 8.0 * (double)t)) * 0.5);
  if (i < 0) i = 0;
  if (i >= nt) i = nt - 1;
  // rowstart(i) = i*nt - i(i-1)/2
  while (i > 0 && i * nt - i * (i - 1) / 2 > t) --i;
  while ((i + 1) * nt - (i + 1) * i / 2 <= t) ++i;
  I = (int)i;
  J = (int)(t - (i * nt - i * (i - 1) / 2) + i);
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace skg
