#include "pic_variant_macros.cuh"
namespace pic { const void* resident_kernel_f32(int threads, int ppt, int dep, bool exact_w) {
    PIC_R_CASE(float, 1024, 1, pic::DEP_SPLIT32, false) PIC_R_CASE(float, 1024, 2, pic::DEP_SPLIT32, false)
    PIC_R_CASE(float, 1024, 3, pic::DEP_SPLIT32, false) PIC_R_CASE(float, 1024, 4, pic::DEP_SPLIT32, false)
    PIC_R_CASE(float, 1024, 5, pic::DEP_SPLIT32, false) PIC_R_CASE(float, 1024, 6, pic::DEP_SPLIT32, false)
    PIC_R_CASE(float, 1024, 8, pic::DEP_SPLIT32, false) PIC_R_CASE(float, 1024, 10, pic::DEP_SPLIT32, false)
    PIC_R_CASE(float, 512, 12, pic::DEP_SPLIT32, false) PIC_R_CASE(float, 512, 16, pic::DEP_SPLIT32, false)
    PIC_R_CASE(float, 512, 20, pic::DEP_SPLIT32, false)
    return nullptr; } }
