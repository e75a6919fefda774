#include "pic_variant_macros.cuh"
namespace pic {
const void* resident_kernel_f64(int threads, int dep, bool exact_w) {
    PIC_R_DEPS(double, 256, false) PIC_R_DEPS(double, 512, false) PIC_R_DEPS(double, 1024, false)
    PIC_R_DEPS(double, 512, true) PIC_R_DEPS(double, 1024, true)
    return nullptr;
}
const void* resident_kernel_f32(int threads, int dep, bool exact_w) {
    PIC_R_CASE(float, 256, pic::DEP_SPLIT32, false) PIC_R_CASE(float, 512, pic::DEP_SPLIT32, false)
    PIC_R_CASE(float, 1024, pic::DEP_SPLIT32, false)
    return nullptr;
}
}  // namespace pic
