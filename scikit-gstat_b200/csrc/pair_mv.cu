// Instantiations of the sweep kernels for multi-column values (cross-variogram |dz_0|*|dz_1|,
// aggregate sqrt(sum dz_k^2); Variogram.py:1965-1984): private records, generic lag lookup.
#include "pair_sweep.cuh"

namespace skg {

template <int DIM, int DIR>
static int hist_mv_stats(int stats, const HistLaunch& L) {
  switch (stats) {
    case 1: return launch_hist<DIM, DIR, 1, 0, true, true>(L);
    case 2: return launch_hist<DIM, DIR, 2, 0, true, true>(L);
    case 4: return launch_hist<DIM, DIR, 4, 0, true, true>(L);
    default: return launch_hist<DIM, DIR, 3, 0, true, true>(L);
  }
}

int dispatch_hist_mv(int dim, int dir, int stats, const HistLaunch& L) {
  if (dim == 3) return hist_mv_stats<3, 0>(stats, L);
  if (dir == 1) return hist_mv_stats<2, 1>(stats, L);
  if (dir == 2) return hist_mv_stats<2, 2>(stats, L);
  return hist_mv_stats<2, 0>(stats, L);
}

template <bool GATHER>
static int select_mv_dim(int dim, int dir, const SelectLaunch& L) {
  if (dim == 3) return launch_select<3, 0, true, SKG_KEY_ABSDZ, GATHER, 0, true>(L);
  if (dir == 1) return launch_select<2, 1, true, SKG_KEY_ABSDZ, GATHER, 0, true>(L);
  if (dir == 2) return launch_select<2, 2, true, SKG_KEY_ABSDZ, GATHER, 0, true>(L);
  return launch_select<2, 0, true, SKG_KEY_ABSDZ, GATHER, 0, true>(L);
}

int dispatch_select_mv(bool gather, int dim, int dir, const SelectLaunch& L) {
  return gather ? select_mv_dim<true>(dim, dir, L) : select_mv_dim<false>(dim, dir, L);
}

}  // namespace skg
