#include "pic_variant_macros.cuh"
namespace pic { const void* resident_kernel_f32(int threads, int ppt, int dep, bool exact_w) {
    PIC_R_DEPS(float, 256, 4, false) PIC_R_DEPS(float, 256, 8, false) PIC_R_DEPS(float, 256, 12, false)
    PIC_R_DEPS(float, 256, 16, false) PIC_R_DEPS(float, 256, 20, false) PIC_R_DEPS(float, 256, 24, false)
    PIC_R_DEPS(float, 256, 32, false) PIC_R_DEPS(float, 256, 40, false)
    return nullptr; } }
