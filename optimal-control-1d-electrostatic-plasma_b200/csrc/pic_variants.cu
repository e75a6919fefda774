// Variant lookup: chains the per-group translation units.
#include "pic_variants.h"
namespace pic {
const void* stream_kernel_f64_a(int, int, int, int, bool);
const void* stream_kernel_f64_b(int, int, int, int, bool);
const void* stream_kernel_f64_c(int, int, int, int, bool);
const void* resident_kernel_f64_a(int, int, int, bool);
const void* resident_kernel_f64_b(int, int, int, bool);
const void* resident_kernel_f64_c(int, int, int, bool);

const void* stream_kernel_f64(int t, int u, int m, int d, bool e) {
    const void* k = stream_kernel_f64_a(t, u, m, d, e);
    if (!k) k = stream_kernel_f64_b(t, u, m, d, e);
    if (!k) k = stream_kernel_f64_c(t, u, m, d, e);
    return k;
}
const void* resident_kernel_f64(int t, int p, int d, bool e) {
    const void* k = resident_kernel_f64_a(t, p, d, e);
    if (!k) k = resident_kernel_f64_b(t, p, d, e);
    if (!k) k = resident_kernel_f64_c(t, p, d, e);
    return k;
}

// Compiled (threads, particles-per-thread) shapes of the resident kernel.  More threads per CTA win as long as the
// per-thread particle state fits the register budget (1024 threads -> 64 registers -> at most 6 fp64 particles).
struct Shape { int threads, ppt; };
static const Shape kShapesF64[] = {{1024, 1}, {1024, 2}, {1024, 3}, {1024, 4}, {1024, 5}, {1024, 6}, {512, 8}, {512, 10},
                                   {512, 12}, {512, 16}, {512, 20}, {256, 24}, {256, 32}, {256, 40}, {0, 0}};
static const Shape kShapesF32[] = {{1024, 1}, {1024, 2}, {1024, 3}, {1024, 4}, {1024, 5}, {1024, 6}, {1024, 8}, {1024, 10},
                                   {512, 12}, {512, 16}, {512, 20}, {0, 0}};
bool resident_pick_shape(long long n, bool f32, int* threads, int* ppt) {
    for (const Shape* s = f32 ? kShapesF32 : kShapesF64; s->threads; ++s)
        if ((long long)s->threads * s->ppt >= n) { *threads = s->threads; *ppt = s->ppt; return true; }
    return false;
}
long long resident_capacity(bool f32) {
    long long best = 0;
    for (const Shape* s = f32 ? kShapesF32 : kShapesF64; s->threads; ++s)
        if ((long long)s->threads * s->ppt > best) best = (long long)s->threads * s->ppt;
    return best;
}
}  // namespace pic
