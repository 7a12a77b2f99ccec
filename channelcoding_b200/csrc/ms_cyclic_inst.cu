// ms_cyclic_inst.cu -- compiled once per -DCCGPU_GROUP=g: instantiates the kernels of
// CCGPU_MS_LIST_g (ms_cyclic_list.h) and exposes them to the registry in ms_registry.cu.
#include "ms_cyclic.cuh"
#include "ms_cyclic_list.h"

#ifndef CCGPU_GROUP
#error "compile with -DCCGPU_GROUP=<0..CCGPU_MS_GROUPS-1>"
#endif
#define CCGPU_CAT_(a, b) a##b
#define CCGPU_CAT(a, b) CCGPU_CAT_(a, b)
#define CCGPU_LIST CCGPU_CAT(CCGPU_MS_LIST_, CCGPU_GROUP)
#define CCGPU_GROUP_FN CCGPU_CAT(ms_cyclic_group_, CCGPU_GROUP)

namespace ccgpu {
#define X(W, RPL, NP, WRAP) CCGPU_MS_CYCLIC_ENTRY(W, RPL, NP, 0, WRAP), CCGPU_MS_CYCLIC_ENTRY(W, RPL, NP, 1, WRAP),
static const MsCyclicEntry kEntries[] = { CCGPU_LIST(X) };
#undef X
const MsCyclicEntry *CCGPU_GROUP_FN(int *count) {
  *count = static_cast<int>(sizeof(kEntries) / sizeof(kEntries[0]));
  return kEntries;
}
}  // namespace ccgpu
