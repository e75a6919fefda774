#include "pic_variant_macros.cuh"
namespace pic { const void* stream_kernel_f32(int threads, int unroll, int mode, int dep, bool exact_w) {
    PIC_S_DEPS(float, 256, 1, false) PIC_S_DEPS(float, 256, 2, false) PIC_S_DEPS(float, 512, 2, false)
    return nullptr; } }
