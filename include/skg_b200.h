/*
 * skg_b200.h - C ABI of the B200-native (sm_100a) experimental-variogram and
 * ordinary-kriging hot path.  This is the drop-in boundary: everything the
 * reference (scikit-gstat 1.0.23, pure Python) computes on this path through
 * scipy/numpy/numba is replaced by the entry points below.  Host code (Python,
 * scikit-gstat_b200/skgstat_b200/_lib.py) binds them with ctypes; a reference
 * maintainer would add the same ctypes stub.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / C++ types in any signature.
 *   - pointers marked [dev] are device pointers on the current CUDA device
 *     (e.g. torch.Tensor.data_ptr()); [host] pointers are ordinary host memory
 *     and are consumed before the call returns.
 *   - coordinates are fp64 struct-of-arrays: coords[k*n + i] = k-th coordinate
 *     of point i, dim in {2, 3} (1-D input gets a zero column appended by the
 *     caller, as Variogram.py:291-297 does).
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as
 *     void*; NULL = legacy default stream); no implicit synchronisation, no
 *     persistent allocation, no global state besides the last-error string.
 *   - the caller owns all buffers, including scratch.  Accumulating outputs
 *     must be zero-initialised by the caller unless stated otherwise.
 *   - return value: 0 ok, <0 invalid argument (SKG_E_*), >0 a cudaError_t.
 *     Never throws, never aborts.  skg_last_error() describes the last failure
 *     of the calling thread.
 *   - the pair triangle is cut into SKG_TILE x SKG_TILE point tiles (upper
 *     triangle incl. diagonal, row-major linearised).  [tile_begin, tile_end)
 *     selects a sub-range, which is how the sweep is sharded across GPUs.
 *
 * Squared-distance space.  Per pair the kernels compute
 *     s = fl(fl(dx*dx) + fl(dy*dy)) [ + fl(dz*dz) ]     (no FMA)
 * which is the operand scipy's euclidean kernel hands to sqrt
 * (MetricSpace.py:151-153 -> scipy pdist).  IEEE sqrt is monotone and
 * correctly rounded, so `d >= e` <=> `s >= T(e)` with
 * T(e) = min{ s : fl(sqrt(s)) >= e }.  Lag-class edges are therefore passed
 * as the BIT PATTERNS of T(e_k) (uint64, monotone for non-negative doubles) and
 * compared as integers: np.digitize(d, edges) (Variogram.py:2005) is reproduced
 * exactly with no sqrt in the hot loop.
 */
#ifndef SKG_B200_H
#define SKG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKG_ABI_VERSION 6
#define SKG_TILE 256            /* points per tile edge                          */
#define SKG_MAX_LAGS 250        /* upper bound for n_lags                        */
#define SKG_KRIGE_MAX_POINTS 128 /* max_points supported by the warp-per-target kernels (<= 64: the fast layout) */

/* error codes (<0) */
#define SKG_E_BADARG (-1)
#define SKG_E_UNSUPPORTED (-2)
#define SKG_E_NODEVICE (-3)
#define SKG_E_SCRATCH (-4)

/* directional model ids (DirectionalVariogram.py:522-575) */
#define SKG_DIR_NONE 0
#define SKG_DIR_TRIANGLE 1 /* DirectionalVariogram._triangle :659-714 */
#define SKG_DIR_COMPASS 2  /* DirectionalVariogram._compass  :749-782 */

/* estimator statistic flags for skg_pair_hist */
#define SKG_STAT_SUM_SQ 1   /* sum (dz)^2     - estimators.matheron :66   */
#define SKG_STAT_SUM_SQRT 2 /* sum |dz|^(1/2) - estimators.cressie  :123  */
#define SKG_STAT_COUNT_BELOW 4 /* (alone) count pairs with |dz| < dz_lo[lag]: exact rank offsets
                                  for the windowed Dowd median select; counts land in sum_sq */

/* value columns: `values` is [n_values][n] (row per column), value_mode says how the columns of
 * a pair combine into the "difference" every statistic / |dz| key / diffs output is built from */
#define SKG_VALUE_SCALAR 0    /* n_values = 1: |z_i - z_j|             (Variogram.py:1938-1943)   */
#define SKG_VALUE_CROSS 1     /* n_values = 2: |dz_0| * |dz_1|         (:1979-1984 cross-variogram) */
#define SKG_VALUE_AGGREGATE 2 /* n_values <= 8: sqrt(sum_k dz_k^2)     (:1974-1977 'aggregate')   */
#define SKG_MAX_VALUE_COLUMNS 8
/* skg_pair_hist only, n_values = 0, values ignored: the "difference" is a power of the pair's
 * distance, so sum_sq / sum_sqrt become moments of the distance distribution per lag class
 * (np.std / skewness inside np.histogram_bin_edges, binning.py:114-122) */
#define SKG_VALUE_DIST_S 3    /* s = d^2:  sum_sq = sum d^4, sum_sqrt = sum d   */
#define SKG_VALUE_DIST_D 4    /* d:        sum_sq = sum d^2                     */
#define SKG_VALUE_DIST_D15 5  /* d^1.5:    sum_sq = sum d^3                     */
/* skg_pair_hist only, OR-ed onto SCALAR / CROSS / AGGREGATE: the difference is squared first, so
 * sum_sqrt = sum |difference| (np.nanmean in estimators.minmax, estimators.py:257-296) */
#define SKG_VALUE_SQUARE_FLAG 0x100

/* key kinds for the select passes */
#define SKG_KEY_SQDIST 0 /* s = squared distance (order statistics of d)       */
#define SKG_KEY_ABSDZ 1  /* |z_i - z_j|        (Dowd per-lag median)            */

/* theoretical models for kriging (models.py) */
#define SKG_MODEL_SPHERICAL 0   /* models.py:21-82   */
#define SKG_MODEL_EXPONENTIAL 1 /* models.py:85-144  */
#define SKG_MODEL_GAUSSIAN 2    /* models.py:147-209 */
#define SKG_MODEL_CUBIC 3       /* models.py:212-272 */
#define SKG_MODEL_STABLE 4      /* models.py:275-344 */
#define SKG_MODEL_MATERN 5      /* models.py:347-422 (K_nu evaluated on the device) */

/* kriging per-target status (Kriging.py:25-38, 492-513) */
#define SKG_KRIGE_OK 0
#define SKG_KRIGE_LESS_POINTS 1 /* LessPointsError  -> z = sigma = NaN */
#define SKG_KRIGE_SINGULAR 2    /* SingularMatrixError -> NaN          */

int skg_abi_version(void);
const char* skg_last_error(void);

/* Number of SMs of the current device (grid sizing is internal; exposed for
 * reporting) and a DFMA-chain microbenchmark used as the FP64 roofline
 * denominator (MEASURED_PEAKS.json has no FP64 figure).  skg_fp64_peak runs
 * `iters` dependent-chain DFMA batches on every SM and returns the elapsed
 * kernel time in ms through *ms and the number of FP64 FMA executed through
 * *fma_count (flops = 2 * fma_count).  Synchronous. */
int skg_device_sm_count(int* sm_count);
int skg_fp64_peak(int iters, double* ms, double* fma_count);

/* Z-order (Morton) key of every point over the box [lo, lo+ext] ([host, dim] each): the pair
 * sweeps are fastest over spatially sorted points (lanes of a warp then fall into the same
 * lag classes); results do not depend on the order.  keys [dev, n] uint64 (out). */
int skg_morton_keys(const double* coords, int64_t n, int dim, const double* lo, const double* ext,
                    uint64_t* keys, void* stream);

/* Number of SKG_TILE tiles in the pair triangle of n points. */
int64_t skg_pair_num_tiles(int64_t n);

/* Directional mask parameters [host, 5 doubles] or NULL for the isotropic case:
 *   dir[0] = model id (SKG_DIR_*), dir[1] = np.radians(azimuth),
 *   dir[2] = np.radians(tolerance / 2), dir[3] = bandwidth / 2,
 *   dir[4] = guard-band policy (below)
 * exactly the quantities DirectionalVariogram._triangle/_compass compare
 * against.  Pair vector is P_i - P_j for i < j (pdist order,
 * DirectionalVariogram.py:404-413).  2-D coordinates only (as the reference).
 *
 * The kernels decide the mask without transcendentals (dot/cross products
 * against the azimuth direction).  The reference decides it through
 * arccos/sin of its host libm (DirectionalVariogram.py:405, 712), so pairs that
 * lie within a relative 2^-40 of a decision boundary (exact geometric ties on
 * lattice data; none on generic float data) are libm-dependent.  Policy
 *   dir[4] = 0: such pairs are decided by the literal formula with the CUDA libm;
 *   dir[4] = 1: such pairs FAIL the mask in every kernel; the caller lists them
 *               once with skg_pair_dir_ambiguous, decides them with the literal
 *               numpy formula on the host (the reference's own arithmetic) and
 *               adds their contributions.  This is what skgstat_b200 does.
 *   scratch [dev] or NULL: >= skg_pair_scratch_bytes(n, 1) bytes lets the call sweep only the
 *   tiles whose difference vectors can come near the mask boundary (Z-order sorted points). */
int skg_pair_dir_ambiguous(const double* coords, int64_t n, int dim, const double* dir,
                           int32_t* pair_i, int32_t* pair_j, uint64_t* count, int64_t capacity,
                           void* scratch, int64_t scratch_bytes,
                           int64_t tile_begin, int64_t tile_end, int64_t tile_stride, void* stream);

/* K2 - fused distance -> lag class -> estimator statistics.
 * Replaces MetricSpace.dists + Variogram.distance + _calc_diff + _calc_groups
 * + lag_classes + estimators.matheron/cressie (MetricSpace.py:135-156,
 * Variogram.py:1202-1208, 1938-2006, 1535-1557, 2092; estimators.py:14-128).
 *   values   [dev, n_values][n]  value columns, combined per pair as value_mode says (SKG_VALUE_*)
 *   thr_bits [host, n_lags]  bit patterns of T(edge_k), non-decreasing
 *   count    [dev, n_lags]   += pairs per lag class (exact)
 *   sum_sq   [dev, n_lags]   += sum dz^2          (if stats & SKG_STAT_SUM_SQ)
 *   sum_sqrt [dev, n_lags]   += sum sqrt|dz|      (if stats & SKG_STAT_SUM_SQRT)
 *   scratch  [dev]           >= skg_pair_scratch_bytes(n, n_lags) bytes
 * Per-CTA partials are reduced in a fixed order: results are run-to-run
 * reproducible for a given (n, tile selection).
 *
 * Tile selection (all pair sweeps): tiles tile_begin, tile_begin + tile_stride, ... below
 * tile_end.  G GPUs take (rank, ntiles, G): interleaved, hence balanced.
 *
 * Tile culling (all pair sweeps with scratch): the coordinates are expected in a
 * spatially coherent order (Z-order, skg_morton_keys).  Each sweep first computes the
 * bounding box of every 256-point block and keeps, in ascending order, only the tiles
 * whose range of squared distances can matter: below the last lag edge (skg_pair_hist),
 * inside a segment window / possible maximum (select passes) - and, with a directional
 * mask, whose box of difference vectors can meet the azimuth cone / bandwidth strip.
 * Exact for ANY point order - an incoherent order merely culls nothing.  After a sweep the first four uint64 of
 * `scratch` hold {surviving work items, -, j-chunks per tile (js), candidate work items};
 * a work item is 256 x (256/js) pair slots. */
int64_t skg_pair_scratch_bytes(int64_t n, int n_lags);
int skg_pair_hist(const double* coords, const double* values, int n_values, int value_mode,
                  int64_t n, int dim,
                  const double* dir, const uint64_t* thr_bits, int n_lags, int stats,
                  const double* dz_lo /* [host, n_lags] or NULL */,
                  uint64_t* count, double* sum_sq, double* sum_sqrt,
                  void* scratch, int64_t scratch_bytes,
                  int64_t tile_begin, int64_t tile_end, int64_t tile_stride, void* stream);

/* K3a - one sweep for the windowed Dowd median (estimators.py:131-178, np.nanmedian per lag class):
 * per lag class the exact pair count, the count of pairs with |dz| below the class's window
 * [dz_lo, dz_hi) (private counters, no atomics) AND the fine histogram of the |dz| inside the
 * window (one L2 atomic for the few in-window pairs; bucket function as skg_pair_select_hist).
 * Replaces a skg_pair_hist(SKG_STAT_COUNT_BELOW) sweep followed by a skg_pair_select_hist sweep.
 *   dz_lo_bits/dz_hi_bits [host, n_lags] bit patterns of the window ends, dz_scale [host, n_lags]
 *   count [dev, n_lags] += pairs;  below [dev, n_lags] (double) += pairs with |dz| < dz_lo
 *   hist  [dev, n_lags*n_buckets] += in-window pairs per bucket
 * Everything else (values, dir, thr_bits, scratch, tiles, culling, determinism) as skg_pair_hist. */
int skg_pair_dz_windows(const double* coords, const double* values, int n_values, int value_mode,
                        int64_t n, int dim, const double* dir, const uint64_t* thr_bits, int n_lags,
                        const uint64_t* dz_lo_bits, const uint64_t* dz_hi_bits, const double* dz_scale,
                        int64_t n_buckets, uint64_t* count, double* below, uint64_t* hist,
                        void* scratch, int64_t scratch_bytes,
                        int64_t tile_begin, int64_t tile_end, int64_t tile_stride, void* stream);

/* K2b - batched fused pass: the estimator sums of n_vectors value vectors that share the
 * coordinates, in ceil(n_vectors/4) sweeps (4 vectors per sweep; distance, lag class and pair
 * count are computed once per pair and sweep).  Replaces the loop of Variogram constructions in
 * util/uncertainty.py:152-171 (Monte-Carlo propagation) and util/cross_variogram.py:70-85
 * (diagonal of the grid): same bins, different values.
 *   values [dev, n_vectors][n]   one row per vector (same point order as coords)
 *   stat                          SKG_STAT_SUM_SQ or SKG_STAT_SUM_SQRT (one of them)
 *   count  [dev, n_lags]         += pairs per lag class (identical for every vector)
 *   sums   [dev, n_vectors][n_lags] += sum dz^2 or sum sqrt|dz| per vector and lag class
 *   scratch >= skg_pair_batch_scratch_bytes(n, n_lags); n_lags <= skg_pair_batch_max_lags()
 * Everything else (dir, thr_bits, tiles, culling, determinism) as skg_pair_hist. */
int skg_pair_batch_max_lags(void);
int64_t skg_pair_batch_scratch_bytes(int64_t n, int n_lags);
int skg_pair_hist_batch(const double* coords, const double* values, int n_vectors, int64_t n, int dim,
                        const double* dir, const uint64_t* thr_bits, int n_lags, int stat,
                        uint64_t* count, double* sums, void* scratch, int64_t scratch_bytes,
                        int64_t tile_begin, int64_t tile_end, int64_t tile_stride, void* stream);

/* K1/K3 - histogram pass of the exact order-statistic search.
 * Replaces np.nanmax / np.median / np.percentile / np.nanpercentile over the
 * condensed distance vector (binning.py:37-40, 70-79; Variogram.py:1288-1293;
 * DirectionalVariogram.py:509) and np.nanmedian per lag class
 * (estimators.py:178) without materialising anything.
 * Segments: n_lags == 0 -> one segment holding every (mask-passing) pair;
 *           n_lags  > 0 -> segment = lag class of the pair (thr_bits as in K2).
 * Key: SKG_KEY_SQDIST or SKG_KEY_ABSDZ.  A key v of segment g is counted iff
 *   lo_bits[g] <= bits(v) < hi_bits[g]; its bucket is
 *   min(n_buckets-1, trunc(fl(fl(v - lo[g]) * scale[g])))   (no FMA; monotone in v)
 *   seg_lo_bits/seg_hi_bits [host, nseg] uint64, seg_scale [host, nseg] double
 *   hist     [dev, nseg*n_buckets] uint64 += counts
 *   max_bits [dev, 1] uint64 atomicMax of the key bits over ALL mask-passing
 *            pairs (window ignored); only written when n_lags == 0; may be NULL.
 *   scratch  [dev] >= skg_pair_scratch_bytes(n, 1) bytes (block boxes + tile list)
 */
int skg_pair_select_hist(const double* coords, const double* values, int n_values, int value_mode,
                         int64_t n, int dim,
                         const double* dir, const uint64_t* thr_bits, int n_lags,
                         int key_kind, const uint64_t* seg_lo_bits,
                         const uint64_t* seg_hi_bits, const double* seg_scale,
                         int64_t n_buckets, uint64_t* hist, uint64_t* max_bits,
                         void* scratch, int64_t scratch_bytes,
                         int64_t tile_begin, int64_t tile_end, int64_t tile_stride, void* stream);

/* K1a - exact pair counts on a log-spaced grid of the squared distance s:
 *   cell(s) = clamp((hi32(bits(s)) >> shift) - key0, 0, SKG_CELL_COUNT - 1)
 * over the (mask-passing) pairs with lo_bits <= bits(s) < hi_bits.  2^(20-shift) cells per octave
 * of s and no threshold compares: the first sweep of every distance order statistic
 * (np.median / np.percentile / np.nanpercentile over the distance vector, Variogram.py:1288-1293,
 * binning.py:70-79, DirectionalVariogram.py:509); the cell holding a wanted rank is narrow enough
 * to be gathered (skg_pair_select_gather) or refined with nearly every tile culled.
 *   count [dev, SKG_CELL_COUNT] += pairs per cell;  scratch >= skg_pair_cell_scratch_bytes(n).
 * With a directional mask under policy dir[4] = 1 the sweep also counts the guard-band pairs it
 * meets (any distance inside the swept tiles) in the 7th uint64 of `scratch`: zero after a sweep
 * over the full distance range means skg_pair_dir_ambiguous has nothing to list. */
#define SKG_CELL_COUNT 4096
int64_t skg_pair_cell_scratch_bytes(int64_t n);
int skg_pair_cell_counts(const double* coords, int64_t n, int dim, const double* dir,
                         uint64_t lo_bits, uint64_t hi_bits, int shift, int key0, uint64_t* count,
                         void* scratch, int64_t scratch_bytes, int64_t tile_begin, int64_t tile_end,
                         int64_t tile_stride, void* stream);

/* Gather pass: every key whose (segment, bucket) bit is set in `bitmap`
 * ([dev] uint32 words, bit index = g*n_buckets + bucket) is appended to
 * cand_key / cand_slot (slot = g*n_buckets + bucket).  cand_count [dev,1]
 * counts all hits, also those beyond `capacity` (which are dropped; the caller
 * then refines the window and repeats).  Order of candidates is unspecified. */
int skg_pair_select_gather(const double* coords, const double* values, int n_values,
                           int value_mode, int64_t n, int dim,
                           const double* dir, const uint64_t* thr_bits, int n_lags,
                           int key_kind, const uint64_t* seg_lo_bits,
                           const uint64_t* seg_hi_bits, const double* seg_scale,
                           int64_t n_buckets, const uint32_t* bitmap,
                           double* cand_key, uint32_t* cand_slot, uint64_t* cand_count,
                           int64_t capacity, void* scratch, int64_t scratch_bytes,
                           int64_t tile_begin, int64_t tile_end, int64_t tile_stride, void* stream);

/* Order statistics of the gathered candidates, picked where they lie (no sort, no host round
 * trip for the candidate count): for each wanted slot want_slot[k] the keys among the first
 * min(*cand_count, capacity) candidates carrying that slot are ranked, and
 *   out[k]                      = number of candidates with that slot          (k < n_want)
 *   out[n_want + r]             = the key of rank want_rank[r] inside its slot
 *                                 (want_off[k] <= r < want_off[k+1]; 0 if the rank is absent)
 * The two values np.median / np.percentile interpolate between (Variogram.py:1288-1293,
 * binning.py:70-79, estimators.py:178) come back in one small copy.  want_* are HOST arrays
 * (n_want <= 32, at most 64 ranks in total); keys/slots/count/out/scratch are device pointers.
 * capacity <= SKG_PICK_DIRECT_MAX: one launch, ranks by counting (O(m^2) compares), the key
 * windows and scratch may be NULL.  Larger: want_lo_bits/want_hi_bits[k] bound the bit patterns
 * of the slot's (non-negative) keys; a histogram linear in the key bits narrows each rank to a
 * bucket whose keys are then ranked by counting (4 launches, scratch >= skg_pick_scratch_bytes()).
 * If more than 4096 keys share a wanted bucket (heavy ties) out[k] = -1 for that slot and the
 * caller has to sort. */
#define SKG_PICK_DIRECT_MAX 16384
int64_t skg_pick_scratch_bytes(void);
int skg_pick_order_stats(const double* cand_key, const uint32_t* cand_slot, const uint64_t* cand_count,
                         int64_t capacity, const uint32_t* want_slot, const int32_t* want_off,
                         const int64_t* want_rank, const uint64_t* want_lo_bits,
                         const uint64_t* want_hi_bits, int n_want, double* out, void* scratch,
                         int64_t scratch_bytes, void* stream);

/* Materialising accessors (Variogram.distance :1202, pairwise_diffs :592-606,
 * lag_groups :1514-1533, DirectionalVariogram._direction_mask :605-624) for the
 * condensed pair range [pair_begin, pair_end), k = n*i - i(i+1)/2 + (j-i-1).
 * Any output may be NULL.  dist = fl(sqrt(s)) (bit-identical to scipy pdist),
 * diffs = |z_i - z_j|, groups = lag class or -1, mask = 1/0. */
int skg_pair_materialize(const double* coords, const double* values, int n_values, int value_mode,
                         int64_t n, int dim,
                         const double* dir, const uint64_t* thr_bits, int n_lags,
                         double* dist, double* diffs, int64_t* groups, uint8_t* mask,
                         int64_t pair_begin, int64_t pair_end, void* stream);

/* K5 - batched ordinary kriging (OrdinaryKriging.transform, Kriging.py:405-650;
 * find_closest MetricSpace.py:26-71; models.py).
 *
 * Step 1, once per sample set: bucket the samples into a uniform cell grid.
 *   grid [host, 8 doubles]: origin x,y,z ; cell edge h ; nx, ny, nz ; unused
 *   cell_start [dev, ncell+1] int32 (out), order [dev, n] int32 (out):
 *   order[cell_start[c] .. cell_start[c+1]) = sample ids of cell c, ascending.
 *   scratch [dev] >= (ncell + n) * 4 bytes.
 */
int skg_krige_build_grid(const double* coords, int64_t n, int dim, const double* grid,
                         int32_t* cell_start, int32_t* order,
                         void* scratch, int64_t scratch_bytes, void* stream);

/* Step 2: one warp per target.  For target t: neighbours = samples with
 * d <= radius; if more than max_points, the max_points smallest by
 * (d, sample id) (stable argsort, MetricSpace.py:64-70); fewer than min_points
 * -> status LESS_POINTS, NaN.  System: A_ij = gamma(d_ij) (i != j), A_ii = 0,
 * border 1, corner 0; b_i = gamma(d_0i), b_n = 1 (Kriging.py:594-614); solved
 * by partial-pivot LU in fp64; sigma = sum(w_i b_i) + lambda, z = w . values
 * (Kriging.py:644-647).
 *   model_params [host, 6 doubles]: range r, sill c0, shape/smoothness s (stable, matern),
 *                nugget b (0 = none) - the cof of Variogram.fitted_model_function -,
 *                then the table range and the beyond-table value of matrix_lut
 *   matrix_lut [dev, lut_size] or NULL: OrdinaryKriging(mode='estimate') (Kriging.py:656-679) -
 *                the kriging MATRIX takes gamma from this table, index
 *                int((d / model_params[4]) * lut_size), model_params[5] beyond it; the
 *                right-hand side always uses the exact model (Kriging.py:611-613)
 *   targets [dev, dim*m] SoA; z, sigma [dev, m] (out); status [dev, m] int32 (out)
 *   scratch [dev] >= skg_krige_scratch_bytes(max_points) bytes
 *   [target_begin, target_end) sub-range (multi-GPU sharding; outputs indexed
 *   by absolute target id).
 *   exclude [dev, m] int32 or NULL: sample id that must not be used as a neighbour
 *   of target t (-1 = none).  Leave-one-out cross validation
 *   (util/cross_validation.py:9-21 rebuilds an OrdinaryKriging without sample i and
 *   krigs at its location) is ONE call with targets = samples, exclude[t] = t.
 *   The solve is a Cholesky factorisation of the covariance form of the system on
 *   the FP64 tensor cores; targets whose covariance matrix is not positive
 *   definite in fp64 are re-solved by partial-pivot LU of the original system.
 *   sample_cov [dev, n*n] or NULL: c_t - gamma(d_ij) of ALL sample pairs, filled once per sample
 *   set and model by skg_krige_sample_cov.  The kriging matrix of a target only holds
 *   sample-to-sample entries, so with the cache its build is a gather (rows of neighbouring
 *   targets stay L2-resident) instead of n(n-1)/2 distance + model evaluations - the same idea as
 *   the reference's cached N x N sample distances (MetricSpace.diagonal, MetricSpace.py:158-187).
 *   Bit-identical results with and without the cache.
 */
int skg_krige_sample_cov(const double* coords, int64_t n, int dim, int model, const double* model_params,
                         const double* matrix_lut, int lut_size, double* cov, void* stream);
int64_t skg_krige_scratch_bytes(int max_points);
int skg_krige(const double* coords, const double* values, int64_t n, int dim,
              const double* grid, const int32_t* cell_start, const int32_t* order,
              const double* targets, int64_t m, int model, const double* model_params,
              const double* matrix_lut, int lut_size,
              double radius, int min_points, int max_points, const int32_t* exclude,
              const double* sample_cov, double* z, double* sigma, int32_t* status,
              void* scratch, int64_t scratch_bytes,
              int64_t target_begin, int64_t target_end, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SKG_B200_H */
