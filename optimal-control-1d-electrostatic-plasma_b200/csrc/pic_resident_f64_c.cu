#include "pic_variant_macros.cuh"
namespace pic { const void* resident_kernel_f64_c(int threads, int ppt, int dep, bool exact_w) {
    PIC_R_DEPS(double, 128, 40, false) PIC_R_DEPS(double, 512, 10, false) PIC_R_DEPS(double, 512, 20, false)
    PIC_R_DEPS(double, 1024, 5, false) PIC_R_DEPS(double, 256, 20, true)
    return nullptr; } }
