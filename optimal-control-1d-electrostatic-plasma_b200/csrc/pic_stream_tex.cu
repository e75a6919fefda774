#include "pic_variant_macros.cuh"
// gather = texture: the streaming passes with the field table read through the texture pipe (split32 deposit, 1024 x 2)
#define PIC_S_TEX(R, MD, IPV) \
    if (mode == MD && ip == IPV) \
        return (const void*)&pic::push_stream_kernel<R, 1024, 2, MD, pic::DEP_SPLIT32, false, IPV, true>;
#define PIC_S_TEX_MODES(R, IPV) PIC_S_TEX(R, pic::MODE_KICK, IPV) PIC_S_TEX(R, pic::MODE_KICK0, IPV) PIC_S_TEX(R, pic::MODE_FINAL, IPV)
namespace pic {
const void* stream_kernel_tex(bool f32, int threads, int unroll, int mode, int ip) {
    if (threads != 1024 || unroll != 2) return nullptr;
    if (!f32) { PIC_S_TEX_MODES(double, pic::IP_CIC) PIC_S_TEX_MODES(double, pic::IP_TSC)
                PIC_S_TEX(double, pic::MODE_INIT, pic::IP_TSC) }    // table-less init deposit (large TSC meshes)
    else { PIC_S_TEX_MODES(float, pic::IP_CIC) }
    return nullptr;
}
const void* field_table_kernel_for(bool f32) {
    return f32 ? (const void*)&pic::field_table_kernel<float, 1024> : (const void*)&pic::field_table_kernel<double, 1024>;
}
}  // namespace pic
