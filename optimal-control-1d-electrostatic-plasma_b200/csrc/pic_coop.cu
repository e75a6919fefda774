#include "pic_kernels.cuh"
#include "pic_variants.h"
// Cooperative single-launch env step (step_coop_kernel): CIC, split32 deposit, the default 1024 x 2 launch shape.
namespace pic {
const void* coop_kernel(bool f32, int threads, int unroll, int dep, bool exact_w, int ip) {
    if (threads != 1024 || unroll != 2 || dep != DEP_SPLIT32 || exact_w || ip != IP_CIC) return nullptr;
    return f32 ? (const void*)&step_coop_kernel<float, 1024, 2, DEP_SPLIT32, false>
               : (const void*)&step_coop_kernel<double, 1024, 2, DEP_SPLIT32, false>;
}
}  // namespace pic
