#include "pic_variant_macros.cuh"
namespace pic { const void* resident_kernel_f64_c(int threads, int ppt, int dep, bool exact_w) {
    PIC_R_DEPS(double, 256, 20, false) PIC_R_DEPS(double, 256, 24, false) PIC_R_DEPS(double, 256, 32, false)
    PIC_R_DEPS(double, 256, 40, false)
    return nullptr; } }
