#include "pic_variant_macros.cuh"
// resident kernel with one env spread over a CTA cluster (distributed shared memory): float64 / float32, CIC and TSC
namespace pic {
#define PIC_C_CASE(R, T, C, IP) \
    if (f32 == (sizeof(R) == 4) && threads == T && cl == C && ip == IP) \
        return (const void*)&pic::env_step_cluster_kernel<R, T, C, false, IP>;
const void* cluster_kernel(bool f32, int threads, int cl, int ip) {
    PIC_C_CASE(double, 256, 2, pic::IP_CIC) PIC_C_CASE(double, 512, 2, pic::IP_CIC) PIC_C_CASE(double, 1024, 2, pic::IP_CIC)
    PIC_C_CASE(double, 256, 4, pic::IP_CIC) PIC_C_CASE(double, 512, 4, pic::IP_CIC) PIC_C_CASE(double, 1024, 4, pic::IP_CIC)
    PIC_C_CASE(double, 1024, 8, pic::IP_CIC)
    PIC_C_CASE(float, 256, 2, pic::IP_CIC) PIC_C_CASE(float, 512, 2, pic::IP_CIC) PIC_C_CASE(float, 1024, 2, pic::IP_CIC)
    PIC_C_CASE(double, 256, 2, pic::IP_TSC) PIC_C_CASE(double, 512, 2, pic::IP_TSC) PIC_C_CASE(double, 1024, 2, pic::IP_TSC)
    return nullptr;
}
}  // namespace pic
