// ms_cyclic_inst.cu -- compiled once per -DCCGPU_GROUP=g: instantiates ms_cyclic_kernel for the
// shapes of CCGPU_MS_LIST_g (ms_shapes_generated.h), four message-update flavours each plus the three QUICK ones and the fixed-point kernel (ms_cyclic_q.cuh), and
// exposes them to the registry in ms_registry.cu.
#include "ms_cyclic.cuh"
#include "ms_cyclic_q.cuh"
#include "ms_shapes_generated.h"

#ifndef CCGPU_GROUP
#error "compile with -DCCGPU_GROUP=<0..CCGPU_MS_GROUPS-1>"
#endif
#define CCGPU_CAT_(a, b) a##b
#define CCGPU_CAT(a, b) CCGPU_CAT_(a, b)
#define CCGPU_LIST CCGPU_CAT(CCGPU_MS_LIST_, CCGPU_GROUP)
#define CCGPU_GROUP_FN CCGPU_CAT(ms_cyclic_group_, CCGPU_GROUP)

namespace ccgpu {

template <class S> struct TapTable {
  int v[S::W];
  constexpr TapTable() : v{} {
    for (int j = 0; j < S::W; ++j) v[j] = S::taps::get(j);
  }
};
template <class S> static const TapTable<S> kTapTable{};

template <class S, int VN> MsCyclicEntry make_entry(const char *name) {
  return MsCyclicEntry{ name, S::N, S::K, S::W, S::RPL, S::FPW, S::NP, S::WRAP ? 1 : 0, VN, kMsThreads, 0, 1, 0, kTapTable<S>.v,
                        reinterpret_cast<ms_kernel_fn>(&ms_cyclic_kernel<S, VN>) };
}
template <class S> MsCyclicEntry make_q_entry(const char *name) {
  return MsCyclicEntry{ name, S::N, S::K, S::W, S::RPL, S::FPW, S::NP, S::WRAP ? 1 : 0, VN_FIX, kMsThreads, 0, 2, 0, kTapTable<S>.v,
                        reinterpret_cast<ms_kernel_fn>(&ms_cyclic_q_kernel<S>) };
}

#define X(NAME) make_entry<shapes::NAME, VN_PLAIN>(#NAME), make_entry<shapes::NAME, VN_SC>(#NAME), \
                make_entry<shapes::NAME, VN_2D>(#NAME), make_entry<shapes::NAME, VN_SPA>(#NAME), \
                make_entry<shapes::NAME, VN_QUICK + VN_PLAIN>(#NAME), make_entry<shapes::NAME, VN_QUICK + VN_SC>(#NAME), \
                make_entry<shapes::NAME, VN_QUICK + VN_2D>(#NAME), make_q_entry<shapes::NAME>(#NAME),
static const MsCyclicEntry kEntries[] = { CCGPU_LIST(X) };
#undef X

const MsCyclicEntry *CCGPU_GROUP_FN(int *count) {
  *count = static_cast<int>(sizeof(kEntries) / sizeof(kEntries[0]));
  return kEntries;
}

}  // namespace ccgpu
