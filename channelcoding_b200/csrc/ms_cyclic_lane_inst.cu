// ms_cyclic_lane_inst.cu -- instantiates ms_cyclic_lane_kernel (one lane per frame) for the shapes of CCGPU_MS_LANE_LIST
// (ms_shapes_generated.h): plain (MS / NMS / OMS) and two-dimensional flavours, sum-product for CCGPU_MS_LANE_SPA_LIST.
#include "ms_cyclic_lane.cuh"
#include "ms_shapes_generated.h"

namespace ccgpu {

template <class S> struct LaneTapTable {
  int v[S::W];
  constexpr LaneTapTable() : v{} {
    for (int j = 0; j < S::W; ++j) v[j] = S::taps::get(j);
  }
};
template <class S> static const LaneTapTable<S> kLaneTapTable{};

// `cta` = 2 marks the lane-per-frame mapping: a CTA of kMsThreads threads works on kMsThreads frames at once
template <class S, int VN> MsCyclicEntry make_lane_entry(const char *name) {
  return MsCyclicEntry{ name, S::N, S::K, S::W, 1, 32, 1, 0, VN, kMsThreads, 2, 1, ms_lane_smem_bytes<S>(), kLaneTapTable<S>.v,
                        reinterpret_cast<ms_kernel_fn>(&ms_cyclic_lane_kernel<S, VN>) };
}

#define X(NAME) make_lane_entry<shapes::NAME, VN_PLAIN>(#NAME), make_lane_entry<shapes::NAME, VN_2D>(#NAME),
#define Y(NAME) make_lane_entry<shapes::NAME, VN_SPA>(#NAME),
static const MsCyclicEntry kLaneEntries[] = { CCGPU_MS_LANE_LIST(X) CCGPU_MS_LANE_SPA_LIST(Y) };
#undef X
#undef Y

const MsCyclicEntry *ms_cyclic_group_lane(int *count) {
  *count = static_cast<int>(sizeof(kLaneEntries) / sizeof(kLaneEntries[0]));
  return kLaneEntries;
}

}  // namespace ccgpu
