// ms_registry.cu -- lookup over the kernel instantiation groups (ms_cyclic_inst.cu x CCGPU_GROUP).
#include "ms_params.h"
#include "ms_shapes_generated.h"

namespace ccgpu {

#define CCGPU_DECL(g) const MsCyclicEntry *ms_cyclic_group_##g(int *count);
CCGPU_DECL(0) CCGPU_DECL(1) CCGPU_DECL(2) CCGPU_DECL(3) CCGPU_DECL(4) CCGPU_DECL(5) CCGPU_DECL(6) CCGPU_DECL(7)
#undef CCGPU_DECL
const MsCyclicEntry *ms_cyclic_group_cta(int *count);
const MsCyclicEntry *ms_cyclic_group_lane(int *count);
static_assert(CCGPU_MS_GROUPS == 8, "update the declarations above");

using group_fn = const MsCyclicEntry *(*)(int *);
static const group_fn kGroups[CCGPU_MS_GROUPS + 2] = { ms_cyclic_group_0, ms_cyclic_group_1, ms_cyclic_group_2,
                                                       ms_cyclic_group_3, ms_cyclic_group_4, ms_cyclic_group_5,
                                                       ms_cyclic_group_6, ms_cyclic_group_7, ms_cyclic_group_cta,
                                                       ms_cyclic_group_lane };

int ms_cyclic_count() {
  int total = 0;
  for (group_fn g : kGroups) {
    int c = 0;
    g(&c);
    total += c;
  }
  return total;
}

const MsCyclicEntry *ms_cyclic_at(int i) {
  for (group_fn g : kGroups) {
    int c = 0;
    const MsCyclicEntry *e = g(&c);
    if (i < c) return e + i;
    i -= c;
  }
  return nullptr;
}

}  // namespace ccgpu
