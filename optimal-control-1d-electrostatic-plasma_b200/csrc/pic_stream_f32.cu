#include "pic_variant_macros.cuh"
namespace pic { const void* stream_kernel_f32(int threads, int unroll, int mode, int dep, bool exact_w) {
    PIC_S_MODES(float, 256, 2, pic::DEP_SPLIT32, false) PIC_S_MODES(float, 512, 2, pic::DEP_SPLIT32, false)
    PIC_S_MODES(float, 1024, 1, pic::DEP_SPLIT32, false) PIC_S_MODES(float, 1024, 2, pic::DEP_SPLIT32, false)
    return nullptr; } }
