// Kernel variant lookup shared between the per-group translation units and the C ABI.
#pragma once
#include <cuda_runtime.h>

namespace pic {

// return nullptr when the combination was not compiled
const void* stream_kernel_f64(int threads, int unroll, int mode, int dep, bool exact_w);
const void* stream_kernel_f32(int threads, int unroll, int mode, int dep, bool exact_w);
const void* resident_kernel_f64(int threads, int dep, bool exact_w);
const void* resident_kernel_f32(int threads, int dep, bool exact_w);
const void* stream_kernel_tsc(int threads, int unroll, int mode);     // TSC: float64, split32, default shapes
const void* resident_kernel_tsc(int threads);
const void* cluster_kernel(bool f32, int threads, int cluster, int ip);   // env over a CTA cluster (split32 deposit)
const void* stream_kernel_tex(bool f32, int threads, int unroll, int mode, int ip);   // gather through the texture pipe
const void* field_table_kernel_for(bool f32);                             // writes the table those kernels read
const void* coop_kernel(bool f32, int threads, int unroll, int dep, bool exact_w, int ip);   // whole steps in one cooperative launch

}  // namespace pic
