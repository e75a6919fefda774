// Packed float32 pair path (f32x2 in pic_device.cuh) against the scalar path, lane by lane, on random particles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I optimal-control-1d-electrostatic-plasma_b200/csrc -o build/f32x2_check tools/f32x2_check.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "pic_device.cuh"
using namespace pic;

__global__ void check(const float* x, const float* v, const float2* E, int n, MeshConst mc, PartConst<float> pc, float cc, float dd,
                      unsigned long long* stats) {
    extern __shared__ float2 E_s[];
    for (int j = threadIdx.x; j < mc.M; j += blockDim.x) E_s[j] = E[j];
    __syncthreads();
    for (int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x); i + 1 < n; i += 2 * gridDim.x * blockDim.x) {
        float xs[2], vs[2]; int il_s[2]; long long Wa_s[2], Wb; bool slow_s[2];
        for (int k = 0; k < 2; ++k)
            slow_s[k] = particle_fast<float, IP_CIC, true, true, false>(x[i + k], v[i + k], xs[k], vs[k], il_s[k], Wa_s[k], Wb, E_s, cc, dd, pc, mc);
        float2 xn, vn, wr; int i0, i1;
        bool p0, p1;
        f32x2::particle_fast(make_float2(x[i], x[i + 1]), make_float2(v[i], v[i + 1]), xn, vn, i0, i1, wr, E_s, cc, dd, pc, mc.M, p0, p1);
        const bool slow_p = p0 | p1;
        if (p0 != slow_s[0] || p1 != slow_s[1]) atomicAdd(&stats[7], 1ull);
        long long W0 = fix_weight((double)wr.x, mc.fix_scale), W1 = fix_weight((double)wr.y, mc.fix_scale);
        atomicAdd(&stats[0], 1ull);
        if (slow_p != (slow_s[0] | slow_s[1])) atomicAdd(&stats[1], 1ull);
        if (slow_p) atomicAdd(&stats[2], 1ull);
        if (!slow_p) {
            if (__float_as_int(xn.x) != __float_as_int(xs[0]) || __float_as_int(xn.y) != __float_as_int(xs[1])) atomicAdd(&stats[3], 1ull);
            if (__float_as_int(vn.x) != __float_as_int(vs[0]) || __float_as_int(vn.y) != __float_as_int(vs[1])) atomicAdd(&stats[4], 1ull);
            if (i0 != il_s[0] || i1 != il_s[1]) atomicAdd(&stats[5], 1ull);
            if (W0 != Wa_s[0] || W1 != Wa_s[1]) atomicAdd(&stats[6], 1ull);
        }
    }
}

int main() {
    const int n = 1 << 22, M = 4096;
    MeshConst mc{};
    mc.M = M; mc.L = 50.0; mc.dx = mc.L / M; mc.inv_dx = 1.0 / mc.dx; mc.dt = 4.4e-4; mc.fix_scale = ldexp(1.0, 41); mc.fix_one = 1ll << 41;
    mc.idx_thr = M * ldexp(1.0, -22);
    PartConst<float> pc = make_part_const<float>(mc);
    std::vector<float> x(n), v(n); std::vector<float2> E(M);
    srand(1);
    for (int i = 0; i < n; ++i) { x[i] = 50.0f * (rand() / (RAND_MAX + 1.0f)); v[i] = 6.0f * (rand() / (float)RAND_MAX) - 3.0f; }
    for (int j = 0; j < M; ++j) { E[j].x = sinf(0.01f * j); E[j].y = sinf(0.01f * ((j + 1) % M)); }
    float *dx, *dv; float2* dE; unsigned long long* ds;
    cudaMalloc(&dx, n * 4); cudaMalloc(&dv, n * 4); cudaMalloc(&dE, M * 8); cudaMalloc(&ds, 64);
    cudaMemcpy(dx, x.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(dv, v.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dE, E.data(), M * 8, cudaMemcpyHostToDevice); cudaMemset(ds, 0, 64);
    check<<<148, 256, M * 8>>>(dx, dv, dE, n, mc, pc, 0.6756f, 1.3512f, ds);
    unsigned long long s[8];
    cudaMemcpy(s, ds, 64, cudaMemcpyDeviceToHost);
    printf("%s\npairs %llu | slow-flag mismatch %llu | pair slow %llu | x mismatch %llu | v mismatch %llu | cell mismatch %llu | weight mismatch %llu | per-lane flag mismatch %llu\n",
           cudaGetErrorString(cudaGetLastError()), s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7]);
    return 0;
}
