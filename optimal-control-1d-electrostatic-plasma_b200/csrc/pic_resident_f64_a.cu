#include "pic_variant_macros.cuh"
namespace pic { const void* resident_kernel_f64_a(int threads, int ppt, int dep, bool exact_w) {
    PIC_R_DEPS(double, 1024, 1, false) PIC_R_DEPS(double, 1024, 2, false) PIC_R_DEPS(double, 1024, 3, false)
    PIC_R_DEPS(double, 1024, 4, false) PIC_R_DEPS(double, 1024, 5, false) PIC_R_DEPS(double, 1024, 6, false)
    PIC_R_DEPS(double, 1024, 5, true)
    return nullptr; } }
