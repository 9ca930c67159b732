// Instantiations of the select-pass kernels (K1/K3) for scalar values.
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
