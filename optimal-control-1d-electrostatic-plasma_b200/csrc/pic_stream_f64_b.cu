#include "pic_variant_macros.cuh"
namespace pic { const void* stream_kernel_f64_b(int threads, int unroll, int mode, int dep, bool exact_w) {
    PIC_S_DEPS(double, 512, 1, false) PIC_S_DEPS(double, 512, 2, false) PIC_S_DEPS(double, 512, 4, false)
    PIC_S_MODES(double, 1024, 4, pic::DEP_SPLIT32, false) PIC_S_MODES(double, 768, 2, pic::DEP_SPLIT32, false)
    return nullptr; } }
