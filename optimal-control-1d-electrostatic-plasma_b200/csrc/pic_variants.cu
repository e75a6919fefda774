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

static const int kPpt128[] = {40, 0};
static const int kPpt256[] = {4, 8, 12, 16, 20, 24, 32, 40, 0};
static const int kPpt512[] = {10, 20, 0};
static const int kPpt1024[] = {5, 0};
int resident_pick_ppt(int threads, long long n) {
    const int* t = threads == 128 ? kPpt128 : threads == 256 ? kPpt256 : threads == 512 ? kPpt512
                 : threads == 1024 ? kPpt1024 : nullptr;
    if (!t) return 0;
    for (; *t; ++t) if ((long long)threads * *t >= n) return *t;
    return 0;
}
}  // namespace pic
