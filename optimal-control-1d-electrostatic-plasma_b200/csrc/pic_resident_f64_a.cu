#include "pic_variant_macros.cuh"
namespace pic { const void* resident_kernel_f64_a(int threads, int ppt, int dep, bool exact_w) {
    PIC_R_DEPS(double, 256, 4, false) PIC_R_DEPS(double, 256, 8, false) PIC_R_DEPS(double, 256, 12, false)
    PIC_R_DEPS(double, 256, 16, false) PIC_R_DEPS(double, 256, 20, false) PIC_R_DEPS(double, 256, 4, true)
    return nullptr; } }
