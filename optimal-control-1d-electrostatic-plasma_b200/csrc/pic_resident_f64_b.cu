#include "pic_variant_macros.cuh"
namespace pic { const void* resident_kernel_f64_b(int threads, int ppt, int dep, bool exact_w) {
    PIC_R_DEPS(double, 512, 8, false) PIC_R_DEPS(double, 512, 10, false) PIC_R_DEPS(double, 512, 12, false)
    PIC_R_DEPS(double, 512, 16, false) PIC_R_DEPS(double, 512, 20, false)
    return nullptr; } }
