// Kernel variant lookup shared between the per-precision translation units and the C ABI.
#pragma once
#include <cuda_runtime.h>

namespace pic {

// returns nullptr when the combination was not compiled
const void* stream_kernel_f64(int threads, int unroll, int mode, int dep, bool exact_w);
const void* stream_kernel_f32(int threads, int unroll, int mode, int dep, bool exact_w);
const void* resident_kernel_f64(int threads, int ppt, int dep, bool exact_w);
const void* resident_kernel_f32(int threads, int ppt, int dep, bool exact_w);

// default resident launch shape for n particles per env; false when n exceeds what one CTA keeps in registers
bool resident_pick_shape(long long n, bool f32, int* threads, int* ppt);
long long resident_capacity(bool f32);




}  // namespace pic
