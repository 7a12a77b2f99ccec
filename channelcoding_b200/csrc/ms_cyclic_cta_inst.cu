// ms_cyclic_cta_inst.cu -- instantiates ms_cyclic_cta_kernel for the shapes of CCGPU_MS_CTA_LIST
// (ms_shapes_generated.h), three vertical-node flavours each, plus the fixed-point kernel (ms_cyclic_cta_q.cuh).
#include "ms_cyclic_cta.cuh"
#include "ms_cyclic_cta_q.cuh"
#include "ms_shapes_generated.h"

namespace ccgpu {

template <class S> struct CtaTapTable {
  int v[S::W];
  constexpr CtaTapTable() : v{} {
    for (int j = 0; j < S::W; ++j) v[j] = S::taps::get(j);
  }
};
template <class S> static const CtaTapTable<S> kCtaTapTable{};

template <class S, int VN> MsCyclicEntry make_cta_entry(const char *name) {
  return MsCyclicEntry{ name, S::N, S::K, S::W, S::RPL, 1, S::NPW, S::WRAP ? 1 : 0, VN, S::THREADS, 1, 1, S::DYN_SMEM, kCtaTapTable<S>.v,
                        reinterpret_cast<ms_kernel_fn>(&ms_cyclic_cta_kernel<S, VN>) };
}

template <class S> MsCyclicEntry make_cta_q_entry(const char *name) {
  return MsCyclicEntry{ name, S::N, S::K, S::W, S::RPL, 1, S::NPW, S::WRAP ? 1 : 0, VN_FIX, S::THREADS, 1, 2, 0, kCtaTapTable<S>.v,
                        reinterpret_cast<ms_kernel_fn>(&ms_cyclic_cta_q_kernel<S>) };
}

#define X(NAME) make_cta_entry<shapes::NAME, VN_PLAIN>(#NAME), make_cta_entry<shapes::NAME, VN_SC>(#NAME), \
                make_cta_entry<shapes::NAME, VN_2D>(#NAME), make_cta_q_entry<shapes::NAME>(#NAME),
static const MsCyclicEntry kCtaEntries[] = { CCGPU_MS_CTA_LIST(X) };
#undef X

const MsCyclicEntry *ms_cyclic_group_cta(int *count) {
  *count = static_cast<int>(sizeof(kCtaEntries) / sizeof(kCtaEntries[0]));
  return kCtaEntries;
}

}  // namespace ccgpu
