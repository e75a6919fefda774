#include "pic_variant_macros.cuh"
namespace pic { const void* stream_kernel_f64_a(int threads, int unroll, int mode, int dep, bool exact_w) {
    PIC_S_DEPS(double, 256, 1, false) PIC_S_DEPS(double, 256, 2, false) PIC_S_DEPS(double, 256, 4, false)
    return nullptr; } }
