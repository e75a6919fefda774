#include "pic_variant_macros.cuh"
// interpol="TSC" (src/env/interpolate.py:22-44): float64, native 32-bit atomics
namespace pic {
const void* stream_kernel_tsc(int threads, int unroll, int mode) {
    PIC_S_TSC_MODES(1024, 2) PIC_S_TSC_MODES(512, 2)
    return nullptr;
}
const void* resident_kernel_tsc(int threads) {
    PIC_R_TSC(256) PIC_R_TSC(512) PIC_R_TSC(1024)
    return nullptr;
}
}  // namespace pic
