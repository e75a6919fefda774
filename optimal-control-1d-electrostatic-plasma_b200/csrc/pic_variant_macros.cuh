// Case macros for the per-group variant translation units (split so that nvcc runs on all cores).
#pragma once
#include "pic_kernels.cuh"
#include "pic_variants.h"

#define PIC_S_CASE(R, T, U, MD, DP, EX) \
    if (threads == T && unroll == U && mode == MD && dep == DP && exact_w == EX) \
        return (const void*)&pic::push_stream_kernel<R, T, U, MD, DP, EX>;
#define PIC_S_MODES(R, T, U, DP, EX) \
    PIC_S_CASE(R, T, U, pic::MODE_KICK, DP, EX) PIC_S_CASE(R, T, U, pic::MODE_KICK0, DP, EX) \
    PIC_S_CASE(R, T, U, pic::MODE_FINAL, DP, EX) PIC_S_CASE(R, T, U, pic::MODE_INIT, DP, EX)
#define PIC_S_DEPS(R, T, U, EX) PIC_S_MODES(R, T, U, pic::DEP_CAS64, EX) PIC_S_MODES(R, T, U, pic::DEP_SPLIT32, EX)

#define PIC_R_CASE(R, T, DP, EX) \
    if (threads == T && dep == DP && exact_w == EX) \
        return (const void*)&pic::env_step_resident_kernel<R, T, DP, EX>;
#define PIC_R_DEPS(R, T, EX) PIC_R_CASE(R, T, pic::DEP_CAS64, EX) PIC_R_CASE(R, T, pic::DEP_SPLIT32, EX)

// TSC interpolation: split32 deposit, float64, the default launch shapes only
#define PIC_S_TSC(T, U, MD) \
    if (threads == T && unroll == U && mode == MD) \
        return (const void*)&pic::push_stream_kernel<double, T, U, MD, pic::DEP_SPLIT32, false, pic::IP_TSC>;
#define PIC_S_TSC_MODES(T, U) PIC_S_TSC(T, U, pic::MODE_KICK) PIC_S_TSC(T, U, pic::MODE_KICK0) \
    PIC_S_TSC(T, U, pic::MODE_FINAL) PIC_S_TSC(T, U, pic::MODE_INIT)
#define PIC_R_TSC(T) \
    if (threads == T) return (const void*)&pic::env_step_resident_kernel<double, T, pic::DEP_SPLIT32, false, pic::IP_TSC>;
