// LSU data-pipe cost of streaming global loads / stores by access width (run under ncu; see profiles/README.md).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/lsu_probe tools/lsu_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void ld64(const double* __restrict__ q, double* __restrict__ sink, long long n) {
    double acc = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) acc += __ldcs(q + i);
    if (acc == 12345.678) sink[0] = acc;
}
__global__ void ld128(const double2* __restrict__ q, double* __restrict__ sink, long long n) {
    double acc = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n / 2; i += (long long)gridDim.x * blockDim.x) { double2 t = __ldcs(q + i); acc += t.x + t.y; }
    if (acc == 12345.678) sink[0] = acc;
}
__global__ void ld256(const double* __restrict__ q, double* __restrict__ sink, long long n) {
    double acc = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n / 4; i += (long long)gridDim.x * blockDim.x) {
        double a, b, c, d;
        asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(q + 4 * i));
        acc += a + b + c + d;
    }
    if (acc == 12345.678) sink[0] = acc;
}
__global__ void st64(double* __restrict__ p, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) __stcs(p + i, (double)i);
}
__global__ void st128(double2* __restrict__ p, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n / 2; i += (long long)gridDim.x * blockDim.x) __stcs(p + i, make_double2((double)i, 1.0));
}
__global__ void st256(double* __restrict__ p, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n / 4; i += (long long)gridDim.x * blockDim.x) {
        const double a = (double)i;
        asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p + 4 * i), "d"(a), "d"(1.0), "d"(2.0), "d"(3.0) : "memory");
    }
}
// staged store: registers -> shared (STS.128) -> one bulk copy per warp-row (TMA engine, no LSU store wavefronts)
__global__ void st_bulk(double* __restrict__ p, long long n) {
    extern __shared__ __align__(128) unsigned char sm[];
    double2* buf = (double2*)sm;                       // blockDim.x double2
    const int tid = threadIdx.x;
    const long long per = blockDim.x;                  // double2 per CTA-iteration
    for (long long base = blockIdx.x * per; base + per <= n / 2; base += (long long)gridDim.x * per) {
        buf[tid] = make_double2((double)(base + tid), 1.0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            unsigned s = (unsigned)__cvta_generic_to_shared(buf);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + 2 * base), "r"(s), "r"((unsigned)(per * 16)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
    }
}

int main() {
    const long long n = 1ll << 28;                     // 2 GiB of doubles
    double *p, *q, *sink;
    cudaMalloc(&p, n * 8); cudaMalloc(&q, n * 8); cudaMalloc(&sink, 8);
    cudaMemset(q, 0, n * 8);
    const int g = 148 * 2, b = 1024;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto t = [&](const char* name, auto f) {
        f(); cudaDeviceSynchronize();
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-8s %.3f ms  %.0f GB/s  %s\n", name, ms, n * 8 / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    };
    t("ld64", [&] { ld64<<<g, b>>>(q, sink, n); });
    t("ld128", [&] { ld128<<<g, b>>>((const double2*)q, sink, n); });
    t("ld256", [&] { ld256<<<g, b>>>(q, sink, n); });
    t("st64", [&] { st64<<<g, b>>>(p, n); });
    t("st128", [&] { st128<<<g, b>>>((double2*)p, n); });
    t("st256", [&] { st256<<<g, b>>>(p, n); });
    t("st_bulk", [&] { st_bulk<<<g, b, b * 16>>>(p, n); });
    return 0;
}
