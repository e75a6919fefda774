#include "pic_variant_macros.cuh"
namespace pic { const void* stream_kernel_f64_c(int threads, int unroll, int mode, int dep, bool exact_w) {
    PIC_S_DEPS(double, 1024, 1, false) PIC_S_DEPS(double, 1024, 2, false) PIC_S_DEPS(double, 1024, 2, true)
    return nullptr; } }
