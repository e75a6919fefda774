// Kernel variant lookup shared between the per-precision translation units and the C ABI.
#pragma once
#include <cuda_runtime.h>

namespace pic {

// returns nullptr when the combination was not compiled
const void* stream_kernel_f64(int threads, int unroll, int mode, int dep, bool exact_w);
const void* stream_kernel_f32(int threads, int unroll, int mode, int dep, bool exact_w);
const void* resident_kernel_f64(int threads, int ppt, int dep, bool exact_w);
const void* resident_kernel_f32(int threads, int ppt, int dep, bool exact_w);

// smallest compiled PPT with threads*ppt >= n for the given thread count; 0 when none fits
int resident_pick_ppt(int threads, long long n);




}  // namespace pic
