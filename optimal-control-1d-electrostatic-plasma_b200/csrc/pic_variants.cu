// Variant lookup: chains the per-group translation units of the streaming kernel.
#include "pic_variants.h"
namespace pic {
const void* stream_kernel_f64_a(int, int, int, int, bool);
const void* stream_kernel_f64_b(int, int, int, int, bool);
const void* stream_kernel_f64_c(int, int, int, int, bool);

const void* stream_kernel_f64(int t, int u, int m, int d, bool e) {
    const void* k = stream_kernel_f64_a(t, u, m, d, e);
    if (!k) k = stream_kernel_f64_b(t, u, m, d, e);
    if (!k) k = stream_kernel_f64_c(t, u, m, d, e);
    return k;
}
}  // namespace pic
