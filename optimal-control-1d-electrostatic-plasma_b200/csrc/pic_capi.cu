// C ABI (include/pic_b200.h) over the kernels: handle management, buffer rotation, sub-stage sequencing,
// NCCL all-reduce of the fixed-point density for the particle-sharded mode.
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/pic_b200.h"
#include "pic_kernels.cuh"
#include "pic_variants.h"

using namespace pic;

namespace {

thread_local std::string g_create_error;

// ---- NCCL, resolved at run time from the libnccl already in the process (torch's) ---------------------------
struct nccl_uid { char internal[128]; };               // ncclUniqueId is a 128-byte struct passed by value
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_uid_fn)(nccl_uid*);
typedef int (*nccl_init_rank_fn2)(void**, int, nccl_uid, int);
typedef int (*nccl_destroy_fn)(void*);
typedef const char* (*nccl_errstr_fn)(int);
constexpr int kNcclUint64 = 5, kNcclFloat64 = 8, kNcclSum = 0;

struct NcclApi {
    nccl_allreduce_fn allreduce = nullptr;
    nccl_uid_fn get_uid = nullptr;
    nccl_init_rank_fn2 init_rank = nullptr;
    nccl_destroy_fn destroy = nullptr;
    nccl_errstr_fn errstr = nullptr;
    bool ok = false;
};

NcclApi& nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    void* lib = RTLD_DEFAULT;
    if (!dlsym(RTLD_DEFAULT, "ncclAllReduce")) {
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return api;
    }
    api.allreduce = (nccl_allreduce_fn)dlsym(lib, "ncclAllReduce");
    api.get_uid = (nccl_uid_fn)dlsym(lib, "ncclGetUniqueId");
    api.init_rank = (nccl_init_rank_fn2)dlsym(lib, "ncclCommInitRank");
    api.destroy = (nccl_destroy_fn)dlsym(lib, "ncclCommDestroy");
    api.errstr = (nccl_errstr_fn)dlsym(lib, "ncclGetErrorString");
    api.ok = api.allreduce && api.get_uid && api.init_rank;
    return api;
}

}  // namespace

struct pic_handle {
    pic_config cfg{};
    int device = 0, sm_count = 0, max_smem = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;                        // cfg.stream == PIC_STREAM_OWN: created (non-blocking) and destroyed here
    MeshConst mc{};
    PartConsts pcs{};
    long long N = 0, ld = 0, Ntotal = 0;
    int M = 0, n_envs = 1, esize = 8, fixed_bits = 0, dep = DEP_CAS64;
    bool f32 = false, exact_w = false, resident = false;
    int ip = IP_CIC;                                // interpolation: CIC | TSC
    double cs[4]{}, ds[4]{};

    // tuning
    int threads = 256, per_thread = 2, ctas_per_sm = 0, grid_x = 1, occ = 1;
    int cluster = 1;                                // resident mode: CTAs per env (thread-block cluster over DSMEM)
    size_t smem = 0;

    // device buffers
    void *x = nullptr, *v = nullptr;
    unsigned long long* rho_block = nullptr;        // one allocation [S][W0][W1][W2], each n_envs * M
    unsigned long long* rho[4] = {nullptr, nullptr, nullptr, nullptr};   // W0, W1, W2, S (S and W0 are adjacent)
    double *n = nullptr, *E = nullptr, *diag = nullptr, *vsum = nullptr, *partial = nullptr;
    double *ext = nullptr, *coeffs = nullptr, *bcos = nullptr, *bsin = nullptr, *trace = nullptr;
    double *tw_cos = nullptr, *tw_sin = nullptr, *modes = nullptr, *mode_trace = nullptr;   // spectral read-out
    int n_modes = 0;
    long long mode_trace_cap = 0;
    RewardConst rw{1.0, 1.0, 1.0, 1.0, 1.0};
    unsigned* ph_counts = nullptr; double *ph_feq = nullptr, *ph_kl = nullptr;             // phase-space histogram
    int ph_nb = 0; double ph_vmin = 0, ph_vmax = 0;
    double* stage64 = nullptr;                      // staging for f32 <-> f64 state conversion / cells
    size_t stage64_elems = 0;
    long long coeffs_cap = 0, trace_cap = 0;
    int trace_steps = 0;
    unsigned* err = nullptr;
    int m = 0;
    bool have_state = false, have_basis = false;

    // staged single-stage driving (pic_run_stage)
    const double *stage_ext = nullptr, *stage_coeffs = nullptr;

    // sharding
    void* comm = nullptr;
    bool own_comm = false;
    int rank = 0, world = 1;
    // fused exchange over peer memory (pic_comm_init_peer): no collective library in the step loop
    bool fused = false;
    unsigned long long* exch[COMM_MAX_WORLD] = {};
    unsigned long long* cflags[COMM_MAX_WORLD] = {};
    unsigned long long* exch_mc = nullptr;           // NVLS multicast mapping of the exchange buffers (pic_comm_set_multicast)
    unsigned long long seq = 0, seq_state = 0;       // last exchange issued; the exchange holding the state density
    unsigned* ticket = nullptr;
    // finalize of step t overlapped with the first pass of step t+1 inside a multi-step call (single GPU)
    cudaStream_t fin_stream = nullptr;
    cudaEvent_t ev_pass_done = nullptr, ev_fin_done = nullptr;

    // gather route of the streaming kernels: shared-memory table rebuilt in every CTA's prologue, or one global table per
    // sub-stage (written by field_table_kernel) read through the texture pipe
    int gather_req = PIC_GATHER_AUTO;               // what the caller asked for
    bool texg = false;                              // what is in effect: any stage on the texture route ...
    bool tex_stage[3] = {false, false, false};      // ... and which (Yoshida stages 1, 2, 3)
    void* table[3] = {nullptr, nullptr, nullptr};   // [n_envs][M] pairs, one per kick stage (1, 2, 3)
    cudaTextureObject_t table_tex[3] = {0, 0, 0};
    size_t smem_tex[3] = {0, 0, 0};                 // dynamic shared memory of the TEXG kernel of stage 1, 2, 3
    // Large meshes (TSC at 4096 cells): the two-histogram kernels (stage 3, init) do not fit next to a shared-memory
    // gather table.  They then run their table-less instances -- stage 3 on the texture route whatever was asked for,
    // the init deposit (which gathers nothing) with the table-less layout -- and h->smem is the one-histogram plan.
    bool tableless = false;
    size_t smem_init = 0;

    // whole steps in one cooperative launch (step_coop_kernel): mid-size envs, single GPU
    int coop_req = PIC_COOP_AUTO;                   // what the caller asked for
    bool coop_ok = false;                           // a kernel variant exists and its grid is co-resident on this device
    int coop_workers = 0;                           // pass CTAs per env (one more CTA per env does the finalize)
    size_t coop_smem = 0;
    unsigned long long* coop_bar = nullptr;         // grid-barrier arrival counter (monotonic) ...
    unsigned long long coop_bar_count = 0;          // ... and its value after everything launched so far

    long long launches = 0;
    std::string last_error;
};

namespace {

void drop_handle(pic_handle* h) {                    // pic_create bailing out before any device buffer exists
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int fail(pic_handle* h, int code, const std::string& msg) {
    if (h) h->last_error = msg; else g_create_error = msg;
    return code;
}
#define CK(h, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(h, PIC_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

// Function attributes belong to a kernel on a device, not to a handle, and handles with different meshes share the
// kernels: the opt-in dynamic shared-memory limit (and the carve-out preference) therefore only ever GROW -- a small env
// created after a large one must not pull the limit back under what the large one launches with.
cudaError_t raise_func_attr(const void* k, cudaFuncAttribute attr, int value) {
    static std::mutex mu;
    static std::map<std::tuple<int, const void*, int>, int> current;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(mu);
    int& c = current[std::make_tuple(dev, k, (int)attr)];
    if (value <= c) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(k, attr, value);
    if (e == cudaSuccess) c = value;
    return e;
}

void yoshida(double cs[4], double ds[4]) {          // src/env/integration.py:62-69, same evaluation order
    double phi = pow(2.0, 1.0 / 3.0);
    double w0 = (-1) * phi / (2 - phi);
    double w1 = 1 / (2 - phi);
    cs[0] = cs[3] = 0.5 * w1;
    cs[1] = cs[2] = 0.5 * (w0 + w1);
    ds[0] = 0.0; ds[1] = w1; ds[2] = w0; ds[3] = w1;
}

const void* stream_kernel(const pic_handle* h, int mode) {
    if (h->ip == IP_TSC) return stream_kernel_tsc(h->threads, h->per_thread, mode);
    return h->f32 ? stream_kernel_f32(h->threads, h->per_thread, mode, h->dep, h->exact_w)
                  : stream_kernel_f64(h->threads, h->per_thread, mode, h->dep, h->exact_w);
}
const void* resident_kernel(const pic_handle* h) {
    if (h->cluster > 1)
        return (h->dep == DEP_SPLIT32 && !h->exact_w) ? cluster_kernel(h->f32, h->threads, h->cluster, h->ip) : nullptr;
    if (h->ip == IP_TSC) return resident_kernel_tsc(h->threads);
    return h->f32 ? resident_kernel_f32(h->threads, h->dep, h->exact_w) : resident_kernel_f64(h->threads, h->dep, h->exact_w);
}
long long cluster_slice(const pic_handle* h) { return (h->N + h->cluster - 1) / h->cluster; }
size_t smem_for(const pic_handle* h) {
    if (h->resident && h->cluster > 1)
        return h->f32 ? cluster_smem_bytes<float>(h->M, h->threads, cluster_slice(h), h->cluster, h->ip)
                      : cluster_smem_bytes<double>(h->M, h->threads, cluster_slice(h), h->cluster, h->ip);
    if (h->resident)
        return h->f32 ? resident_smem_bytes<float>(h->M, h->threads, h->N, h->ip)
                      : resident_smem_bytes<double>(h->M, h->threads, h->N, h->ip);
    return h->f32 ? smem_plan_bytes<float>(h->M, h->threads, false, h->ip, true)      // two histograms in the
                  : smem_plan_bytes<double>(h->M, h->threads, false, h->ip, true);    // stage-3 / init kernels
}

const int kStageMode[3] = {MODE_KICK0, MODE_KICK, MODE_FINAL};     // Yoshida stages 1, 2, 3

// Gather route of the streaming kernels (see TexTable in pic_device.cuh), chosen per Yoshida stage.  AUTO takes the
// texture route where it was measured faster (N = 1e9, 4096 cells): the stage-3 pass, whose two deposits + gather
// saturate the LSU data pipe (6.03 -> 5.52 ms under the power cap); not the stage-1 pass, which is bound by HBM alone and
// loses (5.31 -> 6.35 ms: the texture fetches queue behind the pass's own outstanding global loads), nor stage 2 (equal
// under the cap, slower at boost clocks).  Conditions: split32 deposit, default launch shape, and enough particles per
// launch that the extra one-CTA field launch per pass does not show (kTexMinParticles).  With the fused peer exchange
// that one-CTA kernel is the consumer of the ranks' slots.
constexpr long long kTexMinParticles = 1ll << 22;
constexpr int kTexAutoStages = 0x4;                 // bit s-1: stage s
int configure_gather(pic_handle* h) {
    CK(h, cudaSetDevice(h->device));                // allocates tables / texture objects and sets kernel attributes
    h->texg = false;
    for (int i = 0; i < 3; ++i) h->tex_stage[i] = false;
    if (h->resident) return PIC_OK;
    if (h->tableless) {
        const void* ki = stream_kernel_tex(h->f32, h->threads, h->per_thread, MODE_INIT, h->ip);
        CK(h, raise_func_attr(ki, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_init));
    }
    if (h->gather_req == PIC_GATHER_SHARED && !h->tableless) return PIC_OK;
    int mask = h->gather_req == PIC_GATHER_AUTO ? kTexAutoStages
             : h->gather_req == PIC_GATHER_TEXTURE ? 0x7
             : h->gather_req == PIC_GATHER_SHARED ? 0 : (h->gather_req & 0x7);
    if (h->tableless) mask |= 0x4;
    bool have = h->dep == DEP_SPLIT32 && !h->exact_w;
    {                                               // a linear texture addresses at most maxTexture1DLinear texels
        size_t max_texels = 0;
        const cudaChannelFormatDesc fd = h->f32 ? cudaCreateChannelDesc<float2>() : cudaCreateChannelDesc<int4>();
        if (cudaDeviceGetTexture1DLinearMaxWidth(&max_texels, &fd, h->device) != cudaSuccess ||
            (size_t)h->M * (size_t)h->n_envs > max_texels) have = false;
    }
    for (int i = 0; i < 3 && have; ++i) have = stream_kernel_tex(h->f32, h->threads, h->per_thread, kStageMode[i], h->ip) != nullptr;
    if (!have) {
        if (h->tableless)
            return fail(h, PIC_EUNSUPPORTED, "n_mesh too large for the shared-memory mesh tables of this configuration "
                                             "(the table-less texture route needs the 1024 x 2 shape)");
        if (h->gather_req != PIC_GATHER_AUTO && h->gather_req != PIC_GATHER_SHARED)
            return fail(h, PIC_EUNSUPPORTED, "gather = texture needs streaming mode, the split32 deposit, exact_weights = 0, "
                                             "the 1024 x 2 launch shape and n_envs * n_mesh within the 1D texture limit");
        return PIC_OK;
    }
    if (h->gather_req == PIC_GATHER_AUTO && !h->tableless) {
        // measured: float64 CIC gains from ~1e6 particles up; float32 (an 8-byte gather is cheap on the LSU pipe) loses
        if (h->f32 || h->N * (long long)h->n_envs < kTexMinParticles) return PIC_OK;
    }
    const size_t pair = h->f32 ? sizeof(float2) : sizeof(double2);
    cudaDeviceProp prop;
    CK(h, cudaGetDeviceProperties(&prop, h->device));
    for (int i = 0; i < 3; ++i) {
        if (!h->table[i]) {
            CK(h, cudaMalloc(&h->table[i], pair * (size_t)h->M * h->n_envs));
            cudaResourceDesc rd{};
            rd.resType = cudaResourceTypeLinear;
            rd.res.linear.devPtr = h->table[i];
            rd.res.linear.desc = h->f32 ? cudaCreateChannelDesc<float2>() : cudaCreateChannelDesc<int4>();
            rd.res.linear.sizeInBytes = pair * (size_t)h->M * h->n_envs;
            cudaTextureDesc td{};
            td.readMode = cudaReadModeElementType;
            CK(h, cudaCreateTextureObject(&h->table_tex[i], &rd, &td, nullptr));
        }
        const bool second = kStageMode[i] == MODE_FINAL;
        h->smem_tex[i] = h->f32 ? smem_plan_bytes<float>(h->M, h->threads, false, h->ip, second, false)
                                : smem_plan_bytes<double>(h->M, h->threads, false, h->ip, second, false);
        const void* k = stream_kernel_tex(h->f32, h->threads, h->per_thread, kStageMode[i], h->ip);
        CK(h, raise_func_attr(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_tex[i]));
        // the smallest shared-memory carve-out that holds the CTA: everything else of the SM's 256 KB is L1 for the table
        int pct = (int)(((h->smem_tex[i] + 1024) * 100 + prop.sharedMemPerMultiprocessor - 1) / prop.sharedMemPerMultiprocessor);
        if (pct > 100) pct = 100;
        CK(h, raise_func_attr(k, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    }
    const void* kt = field_table_kernel_for(h->f32);
    CK(h, raise_func_attr(kt, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)(h->f32 ? smem_plan_bytes<float>(h->M, 1024, false) : smem_plan_bytes<double>(h->M, 1024, false))));
    for (int i = 0; i < 3; ++i) h->tex_stage[i] = (mask >> i) & 1;
    h->texg = mask != 0;
    return PIC_OK;
}

// Whole steps in one cooperative launch (step_coop_kernel).  Available when a kernel variant exists for the handle's
// flavour (CIC, split32, 1024 x 2, shared-memory gather) and workers + one finalize CTA per env are co-resident.
// AUTO takes it for calls of two or more steps on envs up to kCoopMaxParticles (overridable: PIC_COOP_MAX_PARTICLES).
// Measured (one B200, float64, us per step inside a multi-step call, kernel-per-pass -> cooperative): 2e4 particles /
// 250 cells 26.7 -> 23.4, 1e6 / 1024 42.7 -> 36.6, 1e6 / 4096 62.6 -> 51.8, 3e6 / 4096 86.3 -> 75.7, 1e7 / 4096
// 194.7 -> 189.7, 3e7 equal, 1e8 0.5 % slower (the kernel-per-pass path has the texture-pipe stage 3).  One-step calls
// are no faster (2e4: 32.8 = 32.8) or slower (1e6: 43.5 -> 47.2 us: a cooperative launch costs more than a plain one and
// a lone step has no next pass to hide its finalize behind), so AUTO leaves them on the kernel-per-pass path.
constexpr long long kCoopMaxParticles = 1ll << 24;
int configure_coop(pic_handle* h) {
    h->coop_ok = false;
    if (h->resident || h->tableless) return PIC_OK;
    const void* k = coop_kernel(h->f32, h->threads, h->per_thread, h->dep, h->exact_w, h->ip);
    int can = 0;
    if (!k || cudaDeviceGetAttribute(&can, cudaDevAttrCooperativeLaunch, h->device) != cudaSuccess || !can) return PIC_OK;
    size_t smem = smem_plan_bytes<double>(h->M, 1024, false);          // the finalize CTA's plan
    if (h->smem > smem) smem = h->smem;
    if ((int)smem > h->max_smem) return PIC_OK;
    CK(h, raise_func_attr(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, h->threads, smem));
    const long long per_env = (long long)h->sm_count * occ / h->n_envs;  // co-resident CTAs an env can have
    if (per_env < 2) return PIC_OK;
    h->coop_workers = (int)(per_env - 1 < h->grid_x ? per_env - 1 : h->grid_x);
    h->coop_smem = smem;
    if (!h->coop_bar) {
        CK(h, cudaMalloc(&h->coop_bar, sizeof(unsigned long long)));
        CK(h, cudaMemsetAsync(h->coop_bar, 0, sizeof(unsigned long long), h->stream));
        h->coop_bar_count = 0;
    }
    h->coop_ok = true;
    return PIC_OK;
}
long long coop_max_particles() {
    static const long long v = [] {
        const char* e = getenv("PIC_COOP_MAX_PARTICLES");
        return e && *e ? atoll(e) : kCoopMaxParticles;
    }();
    return v;
}
// does a step_device call of n_steps steps go through the cooperative kernel?
bool coop_in_effect(const pic_handle* h, int n_steps) {
    if (!h->coop_ok || h->coop_req == PIC_COOP_OFF || h->world > 1 || h->fused || h->n_modes > 0) return false;
    if (h->gather_req != PIC_GATHER_AUTO && h->gather_req != PIC_GATHER_SHARED) return false;   // an explicit texture route
    if (h->coop_req == PIC_COOP_ON) return true;
    return n_steps >= 2 && h->N * (long long)h->n_envs <= coop_max_particles();
}

int configure_launch(pic_handle* h) {
    h->smem = smem_for(h);
    h->tableless = false;
    if (!h->resident && (int)h->smem > h->max_smem) {
        const size_t one = h->f32 ? smem_plan_bytes<float>(h->M, h->threads, false, h->ip, false) : smem_plan_bytes<double>(h->M, h->threads, false, h->ip, false);
        const size_t two_tl = h->f32 ? smem_plan_bytes<float>(h->M, h->threads, false, h->ip, true, false) : smem_plan_bytes<double>(h->M, h->threads, false, h->ip, true, false);
        if ((int)one <= h->max_smem && (int)two_tl <= h->max_smem && h->dep == DEP_SPLIT32 && !h->exact_w &&
            stream_kernel_tex(h->f32, h->threads, h->per_thread, MODE_FINAL, h->ip) &&
            stream_kernel_tex(h->f32, h->threads, h->per_thread, MODE_INIT, h->ip)) {
            h->tableless = true; h->smem = one; h->smem_init = two_tl;
        }
    }
    if ((int)h->smem > h->max_smem)
        return fail(h, PIC_EUNSUPPORTED, h->resident ? "env does not fit the shared memory of its CTA(s) (" + std::to_string(h->smem) +
                    " B needed per CTA, " + std::to_string(h->max_smem) + " B available)"
                    : "n_mesh too large for the shared-memory mesh tables (" +
                    std::to_string(h->smem) + " B needed, " + std::to_string(h->max_smem) + " B available)");
    if (h->resident) {
        const void* k = resident_kernel(h);
        if (!k) return fail(h, PIC_EUNSUPPORTED, "no resident kernel variant for threads=" + std::to_string(h->threads) +
                            " cluster=" + std::to_string(h->cluster) + " (clusters: split32 deposit, exact_weights=0)");
        CK(h, raise_func_attr(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
        h->grid_x = h->n_envs * h->cluster;
        return PIC_OK;
    }
    int occ_min = 1 << 30;
    for (int mode = MODE_KICK; mode <= MODE_KICK0; ++mode) {
        if (h->tableless && (mode == MODE_FINAL || mode == MODE_INIT)) continue;
        const void* k = stream_kernel(h, mode);
        if (!k) return fail(h, PIC_EUNSUPPORTED, "no streaming kernel variant for threads=" + std::to_string(h->threads) +
                            " unroll=" + std::to_string(h->per_thread));
        CK(h, raise_func_attr(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
        int occ = 0;
        CK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, h->threads, h->smem));
        if (occ < occ_min) occ_min = occ;
    }
    if (occ_min < 1) return fail(h, PIC_EUNSUPPORTED, "streaming kernel does not fit on an SM");
    h->occ = h->ctas_per_sm > 0 ? (h->ctas_per_sm < occ_min ? h->ctas_per_sm : occ_min) : occ_min;
    long long want = (long long)h->sm_count * h->occ / h->n_envs;
    const int VEC = h->f32 ? 4 : 2;
    long long tiles = (h->N / VEC + (long long)h->threads * h->per_thread - 1) / ((long long)h->threads * h->per_thread);
    if (want > tiles) want = tiles;
    if (want < 1) want = 1;
    h->grid_x = (int)want;
    if (h->partial) { cudaFree(h->partial); h->partial = nullptr; }
    CK(h, cudaMalloc(&h->partial, sizeof(double) * 2 * (size_t)h->grid_x * h->n_envs));
    const void* kf = (const void*)&field_finalize_kernel<1024>;
    CK(h, raise_func_attr(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_plan_bytes<double>(h->M, 1024, false)));
    int rc = configure_gather(h);
    return rc ? rc : configure_coop(h);
}

int comm_slot_len(const pic_handle* h) { return 2 * h->M * h->n_envs + 2 * h->n_envs; }

// launch with programmatic stream serialization (see griddep_wait in pic_device.cuh)
cudaError_t launch_pdl(const void* kernel, dim3 grid, dim3 block, void** args, size_t smem, cudaStream_t stream) {
#ifdef PIC_NO_PDL
    return cudaLaunchKernel(kernel, grid, block, args, smem, stream);
#else
    cudaLaunchConfig_t lc{};
    lc.gridDim = grid; lc.blockDim = block; lc.dynamicSmemBytes = smem; lc.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = 1;
    return cudaLaunchKernelExC(&lc, kernel, args);
#endif
}

void fill_comm(const pic_handle* h, CommArgs& c, unsigned long long seq_in, int in_offset, unsigned long long seq_out,
               const unsigned long long* out_src, int out_words) {
    c.world = h->fused ? h->world : 1; c.rank = h->rank; c.slot_len = comm_slot_len(h);
    for (int r = 0; r < COMM_MAX_WORLD; ++r) { c.exch[r] = h->exch[r]; c.flags[r] = h->cflags[r]; }
    c.exch_mc = h->fused ? h->exch_mc : nullptr;
    c.seq_in = seq_in; c.in_offset = in_offset; c.seq_out = seq_out; c.out_src = out_src; c.out_words = out_words;
    c.ticket = h->ticket;
}

int allreduce_u64(pic_handle* h, unsigned long long* buf, size_t count) {
    if (h->world <= 1 || h->fused) return PIC_OK;
    int r = nccl_api().allreduce(buf, buf, count, kNcclUint64, kNcclSum, h->comm, h->stream);
    if (r != 0) return fail(h, PIC_ENCCL, std::string("ncclAllReduce(uint64): ") +
                            (nccl_api().errstr ? nccl_api().errstr(r) : "error"));
    return PIC_OK;
}

template <typename R, bool EXACT_W, int IP>
__global__ void cells_kernel(const R* __restrict__ x, long long N, long long ld, MeshConst mc, int* il,
                             double* wl, double* wr, double* wm, const double* __restrict__ Emesh, double* Ep) {
    const PartConst<R> pc = make_part_const<R>(mc);
    const int env = blockIdx.y, M = mc.M;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        unsigned err = 0;
        R xw = wrap_pos<R>(x[(size_t)env * ld + i], pc, err);
        size_t o = (size_t)env * N + i;
        const double* E = Emesh + (size_t)env * M;
        if (IP == IP_TSC) {                            // interpolate.py:22-36, pic.py:123: il holds indx_m
            R f;
            int im = cell_index<R>(xw, pc, M, f, err);
            R nr = RT<R>::sub(xw, RT<R>::mul(f, pc.dx));
            R d = EXACT_W ? RT<R>::div(nr, pc.dx) : RT<R>::mul(nr, pc.inv_dx);
            R a, b, c;
            tsc_weights<R>(d, a, b, c);
            if (il) il[o] = im;
            if (wl) wl[o] = (double)a;
            if (wm) wm[o] = (double)b;
            if (wr) wr[o] = (double)c;
            if (Ep) Ep[o] = __dadd_rn(__dadd_rn(__dmul_rn((double)a, E[im == 0 ? M - 1 : im - 1]), __dmul_rn((double)b, E[im])),
                                      __dmul_rn((double)c, E[im == M - 1 ? 0 : im + 1]));
        } else {
            R a, b;
            Cell c = cell_weights<R, EXACT_W>(xw, pc, M, a, b, err);
            if (il) il[o] = c.il;
            if (wl) wl[o] = (double)a;
            if (wr) wr[o] = (double)b;
            if (wm) wm[o] = 0.0;
            if (Ep) Ep[o] = __dadd_rn(__dmul_rn((double)a, E[c.il]), __dmul_rn((double)b, E[c.ir == M ? 0 : c.ir]));   // pic.py:120
        }
    }
}

// Device-side initial sampler: the distributions of src/env/dist.py:70-102 (two-stream) and :151-189
// (bump-on-tail) drawn with Philox instead of host rejection sampling -- same densities (Gaussians truncated to the
// reference's proposal range |v| <= 10, x uniform on [0, L)), same deterministic split of the particle index range
// between the populations, followed by the velocity perturbation of src/env/pic.py:68.  Statistical, not bitwise,
// parity with the host sampler.  Particle i of env e uses Philox subsequence (e * n_global + global index), so the
// result does not depend on the launch shape or on how particles (or envs: env_offset) are sharded over ranks.
template <typename R>
__global__ void sample_kernel(R* __restrict__ x, R* __restrict__ v, long long N, long long ld, long long offset,
                              long long n_global, int kind, double a, double v0, double sigma, double A, int n_mode,
                              double L, unsigned long long seed, long long env_offset) {
    const int env = blockIdx.y;
    const long long n_first = kind == 0 ? (long long)((double)n_global * (1.0 / (1.0 + a))) : n_global / 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const long long g = offset + i;
        curandStatePhilox4_32_10_t st;
        curand_init(seed, (unsigned long long)(env_offset + env) * (unsigned long long)n_global + (unsigned long long)g, 0, &st);
        double mean, sd;
        if (kind == 0) { mean = g < n_first ? 0.0 : v0; sd = g < n_first ? 1.0 : sigma; }   // bump-on-tail
        else           { mean = g < n_first ? v0 : -v0; sd = sigma; }                      // two-stream
        double xx = curand_uniform_double(&st) * L;             // (0, 1] * L
        if (xx >= L) xx = 0.0;
        double vv;
        do { vv = mean + sd * curand_normal_double(&st); } while (fabs(vv) > 10.0);
        vv *= (1 + A * sin(2 * 3.141592653589793 * n_mode * xx / L));
        x[(size_t)env * ld + i] = (R)xx;
        v[(size_t)env * ld + i] = (R)vv;
    }
}

// collective-library sharding: the kinetic sums are all-reduced as two doubles after the finalize kernel (folding them
// into the 64 KB state all-reduce pushes it over NCCL's low-latency protocol limit: measured +100 us per step)
__global__ void apply_vsum_kernel(const double* vsum, double* diag, double* trace, int n_envs) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n_envs) {
        diag[e * DIAG_N + DIAG_KE] = 0.5 * vsum[2 * e];
        diag[e * DIAG_N + DIAG_SUM_V] = vsum[2 * e + 1];
        if (trace) for (int k = 0; k < DIAG_N; ++k) trace[e * DIAG_N + k] = diag[e * DIAG_N + k];
    }
}

// Phase-space histogram of src/control/objective.py:8-14: np.histogram2d(x, v, bins=[nb, nb], range=[[0, L],
// [vmin, vmax]]).  Bin i of a coordinate is the largest i with edge(i) <= value, edges as np.linspace builds them
// (i * step + start, last edge = stop exactly), the right-most edge belongs to the last bin, outliers are dropped.
__device__ __forceinline__ double hist_edge(int i, int nb, double lo, double hi, double step) {
    return i == nb ? hi : __dadd_rn(__dmul_rn((double)i, step), lo);
}
__device__ __forceinline__ int hist_bin(double val, int nb, double lo, double hi, double step, double inv_step) {
    if (!(val >= lo && val <= hi)) return -1;
    int i = (int)floor((val - lo) * inv_step);
    i = i < 0 ? 0 : (i > nb - 1 ? nb - 1 : i);
    while (i > 0 && val < hist_edge(i, nb, lo, hi, step)) --i;
    while (i < nb - 1 && val >= hist_edge(i + 1, nb, lo, hi, step)) ++i;
    return i;
}
template <typename R>
__global__ void phase_hist_kernel(const R* __restrict__ x, const R* __restrict__ v, long long N, long long ld, int nb,
                                  double L, double vmin, double vmax, unsigned* __restrict__ counts) {
    const int env = blockIdx.y;
    const double sx = L / nb, sv = (vmax - vmin) / nb;       // np.linspace: step = (stop - start) / div
    const double isx = 1.0 / sx, isv = 1.0 / sv;
    unsigned* c = counts + (size_t)env * nb * nb;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const int bx = hist_bin((double)x[(size_t)env * ld + i], nb, 0.0, L, sx, isx);
        const int bv = hist_bin((double)v[(size_t)env * ld + i], nb, vmin, vmax, sv, isv);
        if (bx >= 0 && bv >= 0) atomicAdd(&c[(size_t)bx * nb + bv], 1u);
    }
}
// KL divergence of objective.py:16-18: sum(rel_entr(f, feq + 1e-12)) * dx * dv with f = counts * n0 / dx / dv / N
__global__ void kl_kernel(const unsigned* __restrict__ counts, const double* __restrict__ feq, int nb, double scale,
                          double dxdv, double* __restrict__ kl) {
    __shared__ double red[33];
    const int env = blockIdx.x;
    const unsigned* c = counts + (size_t)env * nb * nb;
    double acc = 0.0;
    for (int i = threadIdx.x; i < nb * nb; i += blockDim.x) {
        const double a = (double)c[i] * scale, b = feq[i] + 1e-12;
        if (a > 0.0) acc += a * log(a / b);
    }
    acc = block_sum<256>(acc, red);
    if (threadIdx.x == 0) kl[env] = acc * dxdv;
}

// arguments of one streaming pass: stage 1..3 = Yoshida stages, -1 = init deposit (comm is left to the caller).
// overlap: the flavour of the buffer clearing that lets the finalize of a step run beside the first pass of the next
// (run_stage below, and always inside the cooperative step kernel).
StreamArgs stream_args(const pic_handle* h, int stage, const double* ext, const double* coeffs, bool overlap, int* mode,
                       unsigned long long** reduce, size_t* reduce_count) {
    StreamArgs a{};
    a.mc = h->mc; a.pcs = h->pcs; a.x = h->x; a.v = h->v; a.N = h->N; a.ld = h->ld;
    a.act.ext = ext; a.act.coeffs = coeffs; a.act.bcos = h->bcos; a.act.bsin = h->bsin; a.act.m = h->m;
    a.partial = h->partial; a.err = h->err; a.c_next = h->cs[0];
    const size_t sz = (size_t)h->M * h->n_envs;
    *reduce_count = sz;
    switch (stage) {
        case -1: *mode = MODE_INIT; a.rho_out = h->rho[3]; a.rho_next = h->rho[0]; a.c = 0; a.d = 0;
                 *reduce = h->rho[3]; *reduce_count = 2 * sz; break;
        case 1: *mode = MODE_KICK0; a.c_pre = h->cs[0]; a.rho_in = h->rho[0]; a.rho_out = h->rho[1];
                a.rho_zero = overlap ? nullptr : h->rho[3]; *reduce = h->rho[1]; break;
        case 2: *mode = MODE_KICK; a.rho_in = h->rho[1]; a.rho_out = h->rho[2]; a.rho_zero = h->rho[0];
                a.rho_zero2 = overlap ? h->rho[3] : nullptr; *reduce = h->rho[2]; break;
        default: *mode = MODE_FINAL; a.c_pre = h->cs[2]; a.rho_in = h->rho[2]; a.rho_out = h->rho[3]; a.rho_next = h->rho[0];
                a.rho_zero = h->rho[1]; *reduce = h->rho[3]; *reduce_count = 2 * sz; break;
    }
    if (stage >= 1) { a.c = h->cs[stage]; a.d = h->ds[stage]; }
    return a;
}

FinalizeArgs finalize_args(const pic_handle* h, const double* coeffs, double* trace_row) {
    FinalizeArgs f{};
    f.mc = h->mc; f.rho = h->rho[3]; f.rho_zero = h->rho[2]; f.n_out = h->n; f.E_out = h->E; f.diag = h->diag;
    f.partial = h->partial; f.vsum = h->vsum; f.n_partial = h->grid_x;
    f.rw = h->rw; f.coeffs = coeffs; f.two_m = 2 * h->m; f.step_done = trace_row != nullptr;
    f.tw_cos = h->tw_cos; f.tw_sin = h->tw_sin; f.n_modes = h->n_modes; f.modes = h->n_modes > 0 ? h->modes : nullptr;
    fill_comm(h, f.comm, h->seq_state, 0, 0, nullptr, 0);
    f.rho_reduced = h->rho[3]; f.err = h->err; f.trace_row = trace_row;
    return f;
}

// one streaming sub-stage: stage 0..3 = Yoshida stages, 4 = finalize, -1 = init deposit.
// overlap (single GPU, inside a multi-step call): the finalize of a step runs on `fin_on` beside the first pass of the
// next step; that pass then must not clear the state density the finalize is reading -- the stage-2 pass, which is
// launched after the finalize has been waited for, clears it instead.
int run_stage(pic_handle* h, int stage, const double* ext, const double* coeffs, double* trace_row, bool overlap = false,
              cudaStream_t fin_on = nullptr) {
    if (stage == 4) {
        cudaStream_t fs = fin_on ? fin_on : h->stream;
        const bool nccl_sharded = h->world > 1 && !h->fused;
        FinalizeArgs f = finalize_args(h, coeffs, nccl_sharded ? nullptr : trace_row);
        f.step_done = trace_row != nullptr;
        void* args[] = {&f};
        CK(h, launch_pdl((const void*)&field_finalize_kernel<1024>, dim3(h->n_envs), dim3(1024), args,
                         smem_plan_bytes<double>(h->M, 1024, false), fs));
        h->launches++;
        if (nccl_sharded) {
            int r = nccl_api().allreduce(h->vsum, h->vsum, 2 * (size_t)h->n_envs, kNcclFloat64, kNcclSum, h->comm, h->stream);
            if (r != 0) return fail(h, PIC_ENCCL, "ncclAllReduce(float64) failed");
            apply_vsum_kernel<<<(h->n_envs + 127) / 128, 128, 0, h->stream>>>(h->vsum, h->diag, trace_row, h->n_envs);
            h->launches++;
        }
        return PIC_OK;
    }
    if (stage == 0) return PIC_OK;      // stage 0 (pure drift): deposited ahead of time by stage 3 / init, redone by stage 1
    if (stage != -1 && (stage < 1 || stage > 3)) return fail(h, PIC_EINVAL, "stage must be -1..4");
    int mode;
    unsigned long long* reduce = nullptr; size_t reduce_count = 0;
    StreamArgs a = stream_args(h, stage, ext, coeffs, overlap, &mode, &reduce, &reduce_count);
    const size_t sz = (size_t)h->M * h->n_envs;
    if (stage == -1) CK(h, cudaMemsetAsync(h->rho_block, 0, sizeof(unsigned long long) * 4 * sz, h->stream));
    if (h->fused) {
        // consume: stage 1 reads the next-stage-0 half of the exchange that carried the state (offset sz), stages 2, 3
        // read the previous sub-stage's exchange; produce: a new exchange from the buffer this kernel reduces into
        const unsigned long long seq_in = stage == 1 ? h->seq_state : (stage == -1 ? 0 : h->seq);
        const unsigned long long seq_out = ++h->seq;
        fill_comm(h, a.comm, seq_in, stage == 1 ? (int)sz : 0, seq_out, reduce, (int)reduce_count);
        if (stage == 3 || stage == -1) h->seq_state = seq_out;
    } else {
        fill_comm(h, a.comm, 0, 0, 0, nullptr, 0);
    }
    void* args[] = {&a};
    if (stage >= 1 && h->tex_stage[stage - 1]) {
        // the field of this sub-stage, solved once and written as the gather table the pass reads through the texture pipe
        FieldTableArgs t{};
        t.mc = h->mc; t.rho_in = a.rho_in; t.act = a.act; t.table = h->table[stage - 1];
        t.comm = a.comm; t.err = h->err;                // fused exchange: the table kernel consumes the ranks' slots
        void* targs[] = {&t};
        CK(h, launch_pdl(field_table_kernel_for(h->f32), dim3(h->n_envs), dim3(1024), targs,
                         h->f32 ? smem_plan_bytes<float>(h->M, 1024, false) : smem_plan_bytes<double>(h->M, 1024, false), h->stream));
        a.table_tex = h->table_tex[stage - 1];
        CK(h, launch_pdl(stream_kernel_tex(h->f32, h->threads, h->per_thread, mode, h->ip), dim3(h->grid_x, h->n_envs),
                         dim3(h->threads), args, h->smem_tex[stage - 1], h->stream));
        h->launches += 2;
    } else if (stage == -1 && h->tableless) {
        CK(h, launch_pdl(stream_kernel_tex(h->f32, h->threads, h->per_thread, MODE_INIT, h->ip), dim3(h->grid_x, h->n_envs),
                         dim3(h->threads), args, h->smem_init, h->stream));
        h->launches++;
    } else {
        CK(h, launch_pdl(stream_kernel(h, mode), dim3(h->grid_x, h->n_envs), dim3(h->threads), args, h->smem, h->stream));
        h->launches++;
    }
    return allreduce_u64(h, reduce, reduce_count);     // S and W0 are adjacent: one all-reduce covers both
}

int ensure_trace(pic_handle* h, int n_steps) {
    long long need = (long long)n_steps * h->n_envs * DIAG_N;
    if (need > h->trace_cap) {
        if (h->trace) cudaFree(h->trace);
        h->trace = nullptr; h->trace_cap = 0;
        CK(h, cudaMalloc(&h->trace, sizeof(double) * need));
        h->trace_cap = need;
    }
    h->trace_steps = n_steps;
    if (h->n_modes > 0) {
        long long need_m = (long long)n_steps * h->n_envs * 2 * h->n_modes;
        if (need_m > h->mode_trace_cap) {
            if (h->mode_trace) cudaFree(h->mode_trace);
            h->mode_trace = nullptr; h->mode_trace_cap = 0;
            CK(h, cudaMalloc(&h->mode_trace, sizeof(double) * need_m));
            h->mode_trace_cap = need_m;
        }
    }
    return PIC_OK;
}

int launch_resident(pic_handle* h, int n_steps, const double* ext, const double* coeffs) {
    ResidentArgs a{};
    a.mc = h->mc; a.pcs = h->pcs; a.x = h->x; a.v = h->v; a.N = h->N; a.ld = h->ld; a.n_steps = n_steps;
    a.act.ext = ext; a.act.coeffs = coeffs; a.act.bcos = h->bcos; a.act.bsin = h->bsin; a.act.m = h->m;
    a.coeff_step_stride = (long long)h->n_envs * 2 * h->m; a.ext_step_stride = 0;
    for (int i = 0; i < 4; ++i) { a.c[i] = h->cs[i]; a.d[i] = h->ds[i]; }
    a.n_out = h->n; a.E_out = h->E; a.diag = h->diag; a.trace = n_steps > 0 ? h->trace : nullptr;
    a.rho_out = h->rho[3]; a.err = h->err;
    const bool cl = h->cluster > 1;                  // cluster mode: two histograms, the CTA's particle slice
    const long long n_res = cl ? cluster_slice(h) : h->N;
    a.lay = h->f32 ? smem_offsets<float>(h->M, h->threads, true, h->ip, cl, n_res)
                   : smem_offsets<double>(h->M, h->threads, true, h->ip, cl, n_res);
    a.rw = h->rw; a.tw_cos = h->tw_cos; a.tw_sin = h->tw_sin; a.n_modes = h->n_modes;
    a.modes = h->n_modes > 0 ? h->modes : nullptr;
    a.mode_trace = (h->n_modes > 0 && n_steps > 0) ? h->mode_trace : nullptr;
    void* args[] = {&a};
    if (cl) {
        cudaLaunchConfig_t lc{};
        lc.gridDim = dim3(h->n_envs * h->cluster); lc.blockDim = dim3(h->threads);
        lc.dynamicSmemBytes = h->smem; lc.stream = h->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = h->cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        CK(h, cudaLaunchKernelExC(&lc, resident_kernel(h), args));
    } else {
        CK(h, cudaLaunchKernel(resident_kernel(h), dim3(h->n_envs), dim3(h->threads), args, h->smem, h->stream));
    }
    h->launches++;
    if (a.mode_trace)            // the state's modes = the last row of the per-step record
        CK(h, cudaMemcpyAsync(h->modes, h->mode_trace + (size_t)(n_steps - 1) * h->n_envs * 2 * h->n_modes,
                              sizeof(double) * (size_t)h->n_envs * 2 * h->n_modes, cudaMemcpyDeviceToDevice, h->stream));
    return PIC_OK;
}

// n_steps env steps in one cooperative launch (see step_coop_kernel)
int launch_coop(pic_handle* h, int n_steps, const double* ext, const double* coeffs) {
    CoopArgs c{};
    int mode; unsigned long long* reduce; size_t reduce_count;
    for (int st = 1; st <= 3; ++st) {
        c.s[st - 1] = stream_args(h, st, ext, coeffs, true, &mode, &reduce, &reduce_count);
        fill_comm(h, c.s[st - 1].comm, 0, 0, 0, nullptr, 0);
    }
    c.f = finalize_args(h, coeffs, h->trace);
    c.f.n_partial = h->coop_workers;
    c.n_steps = n_steps; c.coeff_step_stride = (long long)h->n_envs * 2 * h->m; c.trace = h->trace;
    c.barrier = h->coop_bar; c.barrier_base = h->coop_bar_count; c.n_workers = (unsigned)h->coop_workers;
    const dim3 grid(h->coop_workers + 1, h->n_envs);
    void* args[] = {&c};
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = dim3(h->threads); cfg.dynamicSmemBytes = h->coop_smem; cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(h, cudaLaunchKernelExC(&cfg, coop_kernel(h->f32, h->threads, h->per_thread, h->dep, h->exact_w, h->ip), args));
    h->coop_bar_count += 3ull * (unsigned long long)n_steps * grid.x * grid.y;
    h->launches++;
    return PIC_OK;
}

// advance n_steps; ext / coeffs are DEVICE pointers (coeffs: [n_steps][n_envs][2m])
int step_device(pic_handle* h, const double* ext, const double* coeffs, int n_steps) {
    if (!h->have_state) return fail(h, PIC_ESTATE, "pic_set_state has not been called");
    if (n_steps < 1) return fail(h, PIC_EINVAL, "n_steps must be >= 1");
    if (coeffs && (!h->have_basis || h->m < 1)) return fail(h, PIC_ESTATE, "pic_set_actuator_basis has not been called");
    int rc = ensure_trace(h, n_steps);
    if (rc) return rc;
    if (h->resident) return launch_resident(h, n_steps, ext, coeffs);
    if (coop_in_effect(h, n_steps)) return launch_coop(h, n_steps, ext, coeffs);
    // Single GPU, several steps in one call: the finalize of step s (one CTA: state density -> field, energies,
    // reward) runs on a side stream beside the first pass of step s + 1, which needs nothing it produces.  The side
    // stream is joined before the stage-2 pass (which clears the buffers the finalize read / cleared) and at the end of
    // the call, so nothing outside this function ever sees it.
    const bool overlap = h->world <= 1 && !h->fused && n_steps > 1 && h->fin_stream != nullptr;
    bool pending = false;
    for (int s = 0; s < n_steps; ++s) {
        const double* cf = coeffs ? coeffs + (size_t)s * h->n_envs * 2 * h->m : nullptr;
        double* row = h->trace + (size_t)s * h->n_envs * DIAG_N;
        if ((rc = run_stage(h, 1, ext, cf, nullptr, overlap))) return rc;
        if (pending) { CK(h, cudaStreamWaitEvent(h->stream, h->ev_fin_done, 0)); pending = false; }
        if ((rc = run_stage(h, 2, ext, cf, nullptr, overlap))) return rc;
        if ((rc = run_stage(h, 3, ext, cf, nullptr, overlap))) return rc;
        const bool side = overlap && s < n_steps - 1;
        cudaStream_t fs = side ? h->fin_stream : h->stream;
        if (side) {
            CK(h, cudaEventRecord(h->ev_pass_done, h->stream));
            CK(h, cudaStreamWaitEvent(h->fin_stream, h->ev_pass_done, 0));
        }
        if ((rc = run_stage(h, 4, nullptr, cf, row, overlap, fs))) return rc;
        if (h->n_modes > 0)
            CK(h, cudaMemcpyAsync(h->mode_trace + (size_t)s * h->n_envs * 2 * h->n_modes, h->modes,
                                  sizeof(double) * (size_t)h->n_envs * 2 * h->n_modes, cudaMemcpyDeviceToDevice, fs));
        if (side) { CK(h, cudaEventRecord(h->ev_fin_done, h->fin_stream)); pending = true; }
    }
    return PIC_OK;
}

int init_fields(pic_handle* h) {
    int rc;
    if (h->resident) {
        if ((rc = launch_resident(h, 0, nullptr, nullptr))) return rc;
    } else {
        if ((rc = run_stage(h, -1, nullptr, nullptr, nullptr))) return rc;
        if ((rc = run_stage(h, 4, nullptr, nullptr, nullptr))) return rc;
    }
    h->have_state = true;
    return PIC_OK;
}

int ensure_stage64(pic_handle* h, size_t elems) {
    if (elems > h->stage64_elems) {
        if (h->stage64) cudaFree(h->stage64);
        h->stage64 = nullptr; h->stage64_elems = 0;
        CK(h, cudaMalloc(&h->stage64, sizeof(double) * elems));
        h->stage64_elems = elems;
    }
    return PIC_OK;
}

}  // namespace

// =============================================================================================== C ABI
extern "C" {

int pic_abi_version(void) { return PIC_B200_ABI_VERSION; }

const char* pic_build_info(void) {
    return "pic_b200 sm_100a; deposits: cas64,split32; precisions: f64,f32; modes: resident,streaming; nccl: runtime-resolved";
}

double pic_clip_dt(double dt, int64_t n_total, double L) {
    double lim = 2 / sqrt((double)n_total / L);          // src/env/pic.py:71-72
    return dt > lim ? lim : dt;
}

const char* pic_last_error(const pic_handle* h) { return h ? h->last_error.c_str() : g_create_error.c_str(); }

int pic_create(const pic_config* cfg, pic_handle** out) {
    if (!cfg || !out) return fail(nullptr, PIC_EINVAL, "null argument");
    *out = nullptr;
    if (cfg->n_particles < 1 || cfg->n_mesh < 2 || cfg->n_envs < 1 || !(cfg->L > 0) || !(cfg->dt > 0) || !(cfg->n0 > 0))
        return fail(nullptr, PIC_EINVAL, "invalid config (n_particles, n_mesh, n_envs, L, dt, n0 must be positive)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail(nullptr, PIC_ENODEVICE, "no CUDA device: this library has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, PIC_EINVAL, "device ordinal out of range");
    pic_handle* h = new pic_handle();
    h->cfg = *cfg;
    h->device = cfg->device;
    if (cudaSetDevice(h->device) != cudaSuccess) { delete h; return fail(nullptr, PIC_ECUDA, "cudaSetDevice failed"); }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, h->device);
    if (prop.major < 10) { delete h; return fail(nullptr, PIC_ENODEVICE, "device is not sm_100 class (built for sm_100a only)"); }
    h->sm_count = prop.multiProcessorCount;
    h->max_smem = (int)prop.sharedMemPerBlockOptin;
    if (cfg->stream == PIC_STREAM_OWN) {
        if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return fail(nullptr, PIC_ECUDA, "cudaStreamCreate failed"); }
        h->own_stream = true;
    } else {
        h->stream = (cudaStream_t)cfg->stream;
    }
    h->N = cfg->n_particles;
    h->Ntotal = cfg->n_particles_total > 0 ? cfg->n_particles_total : cfg->n_particles;
    h->M = cfg->n_mesh; h->n_envs = cfg->n_envs;
    h->f32 = cfg->precision == PIC_F32; h->esize = h->f32 ? 4 : 8;
    h->exact_w = cfg->exact_weights != 0;
    h->ip = cfg->interpolation == PIC_INTERP_TSC ? IP_TSC : IP_CIC;
    if (h->ip == IP_TSC && (cfg->precision == PIC_F32 || cfg->deposit == PIC_DEPOSIT_CAS64 || h->exact_w)) {
        drop_handle(h);
        return fail(nullptr, PIC_EUNSUPPORTED, "TSC interpolation is built for float64 with the split32 deposit only");
    }
    h->m = cfg->max_mode > 0 ? cfg->max_mode : 0;
    h->ld = (h->N + 15) / 16 * 16;
    yoshida(h->cs, h->ds);
    h->rw = RewardConst{1.0, 1.0, 1.0, 10.0 * cfg->L * 0.25, cfg->L};      // reward.py defaults: n_actions = 10

    MeshConst& mc = h->mc;
    mc.M = h->M; mc.L = cfg->L; mc.dx = cfg->L / cfg->n_mesh; mc.inv_dx = 1.0 / mc.dx; mc.n0 = cfg->n0; mc.dt = cfg->dt;
    mc.dx2 = mc.dx * mc.dx; mc.inv2dx = 1.0 / (2.0 * mc.dx);
    mc.scale = cfg->n0 * cfg->L / (double)h->Ntotal / mc.dx;
    h->dep = cfg->deposit == PIC_DEPOSIT_CAS64 ? DEP_CAS64 : DEP_SPLIT32;     // auto: native 32-bit atomics
    if (h->f32) h->dep = DEP_SPLIT32;
    int mode = cfg->mode;
    // resident = the whole env (particles + mesh tables) fits the shared memory of one CTA
    const size_t res512 = h->f32 ? resident_smem_bytes<float>(h->M, 512, h->N, h->ip) : resident_smem_bytes<double>(h->M, 512, h->N, h->ip);
    const size_t res1024 = h->f32 ? resident_smem_bytes<float>(h->M, 1024, h->N, h->ip) : resident_smem_bytes<double>(h->M, 1024, h->N, h->ip);
    // resident = the whole env (particles + mesh tables) fits the shared memory of one CTA; its field solve is the
    // small-mesh instance (block_field), so N_mesh <= 1024
    if (mode == PIC_MODE_AUTO) {
        mode = ((long long)res1024 <= (long long)h->max_smem && h->M <= FIELD_SMALL_MESH) ? PIC_MODE_RESIDENT : PIC_MODE_STREAMING;
        // an env that needs a CTA PAIR (up to ~26 000 particles at float64) still runs resident when the batch fills the
        // GPU: measured 49 vs 45 G particle-steps/s for 1024 envs x 20 000 particles; larger clusters lose to the streaming
        // kernels (22 vs 46 G for 256 x 40 000) and are only used on request
        if (mode == PIC_MODE_STREAMING && h->M <= FIELD_SMALL_MESH && h->n_envs >= h->sm_count &&
            cfg->deposit != PIC_DEPOSIT_CAS64 && !h->exact_w) {
            const size_t pair = h->f32 ? cluster_smem_bytes<float>(h->M, 1024, (h->N + 1) / 2, 2, h->ip)
                                       : cluster_smem_bytes<double>(h->M, 1024, (h->N + 1) / 2, 2, h->ip);
            if ((long long)pair <= (long long)h->max_smem) mode = PIC_MODE_RESIDENT;
        }
    }
    h->resident = mode == PIC_MODE_RESIDENT;
    if (h->resident && h->M > FIELD_SMALL_MESH) {
        drop_handle(h);
        return fail(nullptr, PIC_EUNSUPPORTED, "resident mode needs n_mesh <= " + std::to_string(FIELD_SMALL_MESH) + " (use streaming mode)");
    }

    int k = cfg->fixed_bits;
    if (k <= 0) {                                         // headroom: 8x the mean per-cell weight sum below 2^62
        double per_cell = (double)h->Ntotal / h->M;
        if (per_cell < 1) per_cell = 1;
        k = 62 - (int)ceil(log2(8.0 * per_cell));
        if (k > 50) k = 50;
        if (k < 20) k = 20;
        // float32 particles carry 24-bit weights: 24 fractional bits lose nothing, a weight then fits the low word of
        // the split deposit and the high word only sees (rare) carries -- see HistSplit32<RARE_HI>
        if (h->f32 && k > 24) k = 24;
    }
    if (k > 50) { drop_handle(h); return fail(nullptr, PIC_EINVAL, "fixed_bits must be <= 50"); }
    h->fixed_bits = k;
    mc.fix_scale = ldexp(1.0, k); mc.inv_fix = ldexp(1.0, -k); mc.fix_one = 1LL << k;
    // bracket half-width of the cell index (fast_cell): 16 ulp in float64 (never hit in practice); in float32 every spare
    // factor costs -- 2 N_mesh eps of all positions take the careful path, and with 32 lanes per warp that is most warps --
    // so eps = 2^-22 = 4 ulp, the smallest power of two that still brackets the rounded quotient (needs 3 x 2^-24:
    // rounding of 1/dx, of (1/dx)(1 -+ eps), and of the quotient itself)
    mc.idx_thr = (double)h->M * (h->f32 ? ldexp(1.0, -22) : ldexp(1.0, -49));
    mc.range_floor = h->ip == IP_TSC ? -(1LL << 61) : -(mc.fix_one << 2);
    h->pcs.d = make_part_const<double>(mc); h->pcs.f = make_part_const<float>(mc);
    mc.field_g = (h->M + 32 * FIELD_VWARPS - 1) / (32 * FIELD_VWARPS);
    mc.field_nvw = ((h->M + mc.field_g - 1) / mc.field_g + 31) / 32;

    if (h->resident) {
        // CTAs per env: one if the env fits one CTA's shared memory, else the smallest thread-block cluster that holds it
        // (each CTA keeps a slice of the particles; histograms meet through distributed shared memory)
        h->threads = 1024;
        for (h->cluster = 1; h->cluster <= 8 && (long long)smem_for(h) > (long long)h->max_smem; h->cluster *= 2) {}
        if (h->cluster > 8 || (h->cluster > 1 && (h->dep != DEP_SPLIT32 || h->exact_w))) {
            drop_handle(h);
            return fail(nullptr, PIC_EUNSUPPORTED, "n_particles too large for resident mode (env does not fit in shared memory)");
        }
        const size_t per_sm = (size_t)h->max_smem + 1024;          // every CTA reserves 1 KB on top of its request
        if (h->cluster == 1) {
            // as many envs per SM as fit side by side, so that one env's field solves overlap the others' particle loops
            // (measured best in every case tried): four 256-thread CTAs, else two of 512 threads, else one of 1024
            const size_t res256 = h->f32 ? resident_smem_bytes<float>(h->M, 256, h->N, h->ip) : resident_smem_bytes<double>(h->M, 256, h->N, h->ip);
            h->threads = 2 * (res512 + 1024) <= per_sm ? 512 : 1024;
            if (h->n_envs <= h->sm_count) {
                // every env has an SM to itself: nothing to overlap, the step latency decides (measured: 1024 threads
                // win from ~4000 particles per env, 512 below)
                h->threads = h->N >= 4096 ? 1024 : 512;
            } else if (4 * (res256 + 1024) <= per_sm) {
                h->threads = 256;
                if (!resident_kernel(h)) h->threads = 512;             // variant not compiled (exact_weights)
            }
        } else {
            for (int t : {256, 512}) {                                 // most CTAs per SM that fit
                h->threads = t;
                if ((1024 / t) * (smem_for(h) + 1024) <= per_sm && resident_kernel(h)) break;
                h->threads = 1024;
            }
        }
        h->per_thread = 0;
    } else {
        h->threads = 1024; h->per_thread = 2;       // measured best: 1024 threads x 2 vectors in flight, 1 CTA per SM
    }

    size_t pbytes = (size_t)h->ld * h->n_envs * h->esize, mbytes = sizeof(double) * (size_t)h->M * h->n_envs;
    bool okm = cudaMalloc(&h->x, pbytes) == cudaSuccess && cudaMalloc(&h->v, pbytes) == cudaSuccess;
    okm = okm && cudaMalloc(&h->rho_block, 4 * mbytes) == cudaSuccess;
    if (okm) {                                       // [S][W0][W1][W2]
        const size_t sz = (size_t)h->M * h->n_envs;
        h->rho[3] = h->rho_block; h->rho[0] = h->rho_block + sz; h->rho[1] = h->rho_block + 2 * sz; h->rho[2] = h->rho_block + 3 * sz;
    }
    okm = okm && cudaMalloc(&h->n, mbytes) == cudaSuccess && cudaMalloc(&h->E, mbytes) == cudaSuccess &&
          cudaMalloc(&h->ext, mbytes) == cudaSuccess &&
          cudaMalloc(&h->diag, sizeof(double) * DIAG_N * h->n_envs) == cudaSuccess &&
          cudaMalloc(&h->vsum, sizeof(double) * 2 * h->n_envs) == cudaSuccess &&
          cudaMalloc(&h->err, sizeof(unsigned)) == cudaSuccess;
    if (okm && h->m > 0)
        okm = cudaMalloc(&h->bcos, sizeof(double) * (size_t)h->M * h->m) == cudaSuccess &&
              cudaMalloc(&h->bsin, sizeof(double) * (size_t)h->M * h->m) == cudaSuccess;
    if (!okm) { std::string e = cudaGetErrorString(cudaGetLastError()); pic_destroy(h); return fail(nullptr, PIC_ENOMEM, "cudaMalloc: " + e); }
    cudaMemsetAsync(h->x, 0, pbytes, h->stream);
    cudaMemsetAsync(h->v, 0, pbytes, h->stream);
    cudaMemsetAsync(h->rho_block, 0, 4 * mbytes, h->stream);
    cudaMemsetAsync(h->diag, 0, sizeof(double) * DIAG_N * h->n_envs, h->stream);
    cudaMemsetAsync(h->vsum, 0, sizeof(double) * 2 * h->n_envs, h->stream);
    cudaMemsetAsync(h->err, 0, sizeof(unsigned), h->stream);
    if (cudaMalloc(&h->ticket, sizeof(unsigned)) != cudaSuccess) { pic_destroy(h); return fail(nullptr, PIC_ENOMEM, "cudaMalloc"); }
    cudaMemsetAsync(h->ticket, 0, sizeof(unsigned), h->stream);
    if (!h->resident) {                              // side stream of the overlapped finalize (step_device)
        if (cudaStreamCreateWithFlags(&h->fin_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_pass_done, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_fin_done, cudaEventDisableTiming) != cudaSuccess) {
            pic_destroy(h);
            return fail(nullptr, PIC_ECUDA, "cudaStreamCreate / cudaEventCreate failed");
        }
    }
    int rc = configure_launch(h);
    if (rc) { g_create_error = h->last_error; pic_destroy(h); return rc; }
    *out = h;
    return PIC_OK;
}

int pic_destroy(pic_handle* h) {
    if (!h) return PIC_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->own_comm && h->comm && nccl_api().destroy) nccl_api().destroy(h->comm);
    void* bufs[] = {h->x, h->v, h->rho_block, h->n, h->E, h->diag, h->vsum, h->partial,
                    h->ext, h->coeffs, h->bcos, h->bsin, h->trace, h->stage64, h->err, h->tw_cos, h->tw_sin, h->modes,
                    h->mode_trace, h->ph_counts, h->ph_feq, h->ph_kl, h->ticket, h->coop_bar};
    for (void* b : bufs) if (b) cudaFree(b);
    for (int i = 0; i < 3; ++i) {
        if (h->table_tex[i]) cudaDestroyTextureObject(h->table_tex[i]);
        if (h->table[i]) cudaFree(h->table[i]);
    }
    if (h->fin_stream) { cudaStreamSynchronize(h->fin_stream); cudaStreamDestroy(h->fin_stream); }
    if (h->ev_pass_done) cudaEventDestroy(h->ev_pass_done);
    if (h->ev_fin_done) cudaEventDestroy(h->ev_fin_done);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return PIC_OK;
}

int pic_set_stream(pic_handle* h, void* s) {
    if (!h) return PIC_EINVAL;
    cudaStreamSynchronize(h->stream);
    if (h->own_stream && h->stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
    h->stream = (cudaStream_t)s;
    return PIC_OK;
}

int pic_get_stream(pic_handle* h, void** out) {
    if (!h || !out) return PIC_EINVAL;
    *out = (void*)h->stream;
    return PIC_OK;
}

int pic_set_tuning(pic_handle* h, int32_t threads, int32_t per_thread, int32_t ctas_per_sm) {
    if (!h) return PIC_EINVAL;
    int t0 = h->threads, p0 = h->per_thread, c0 = h->ctas_per_sm;
    if (threads > 0) h->threads = threads;
    const int cl0 = h->cluster;
    if (per_thread > 0 && !h->resident) h->per_thread = per_thread;
    if (per_thread > 0 && h->resident) h->cluster = per_thread;          // resident mode: CTAs per env (1, 2, 4, 8)
    if (ctas_per_sm >= 0) h->ctas_per_sm = ctas_per_sm;
    cudaStreamSynchronize(h->stream);
    int rc = configure_launch(h);
    if (rc) { h->threads = t0; h->per_thread = p0; h->ctas_per_sm = c0; h->cluster = cl0; configure_launch(h); }
    return rc;
}

int pic_set_gather(pic_handle* h, int32_t route) {
    if (!h) return PIC_EINVAL;
    if (route != PIC_GATHER_AUTO && route != PIC_GATHER_SHARED && route != PIC_GATHER_TEXTURE &&
        (route < PIC_GATHER_TEXTURE_STAGES(1) || route > PIC_GATHER_TEXTURE_STAGES(7)))
        return fail(h, PIC_EINVAL, "route must be PIC_GATHER_AUTO, PIC_GATHER_SHARED, PIC_GATHER_TEXTURE or PIC_GATHER_TEXTURE_STAGES(mask)");
    const int r0 = h->gather_req;
    h->gather_req = route;
    cudaStreamSynchronize(h->stream);
    int rc = configure_gather(h);
    if (rc) { h->gather_req = r0; configure_gather(h); }
    return rc;
}

int pic_get_gather(pic_handle* h, int32_t* route) {
    if (!h || !route) return PIC_EINVAL;
    const int mask = (h->tex_stage[0] ? 1 : 0) | (h->tex_stage[1] ? 2 : 0) | (h->tex_stage[2] ? 4 : 0);
    *route = mask == 0 ? PIC_GATHER_SHARED : mask == 7 ? PIC_GATHER_TEXTURE : PIC_GATHER_TEXTURE_STAGES(mask);
    return PIC_OK;
}

int pic_set_coop(pic_handle* h, int32_t mode) {
    if (!h) return PIC_EINVAL;
    if (mode != PIC_COOP_AUTO && mode != PIC_COOP_OFF && mode != PIC_COOP_ON)
        return fail(h, PIC_EINVAL, "mode must be PIC_COOP_AUTO, PIC_COOP_OFF or PIC_COOP_ON");
    if (mode == PIC_COOP_ON) {
        const int r0 = h->coop_req;
        h->coop_req = mode;
        if (!coop_in_effect(h, 1)) {
            h->coop_req = r0;
            return fail(h, PIC_EUNSUPPORTED, "the cooperative step needs streaming mode on one GPU, CIC, the split32 deposit, "
                        "exact_weights = 0, the 1024 x 2 launch shape, the shared-memory gather, no spectral read-out, and "
                        "a grid (workers + 1 CTA per env) that is co-resident on the device");
        }
        return PIC_OK;
    }
    h->coop_req = mode;
    return PIC_OK;
}

int pic_get_coop(pic_handle* h, int32_t* in_effect, int32_t* workers) {
    if (!h) return PIC_EINVAL;
    const bool on = coop_in_effect(h, 2);
    if (in_effect) *in_effect = on ? 1 : 0;
    if (workers) *workers = on ? h->coop_workers : 0;
    return PIC_OK;
}

int pic_get_launch_info(pic_handle* h, int32_t* mode, int32_t* threads, int32_t* per_thread, int32_t* grid_x,
                        int32_t* smem_bytes, int32_t* fixed_bits, int32_t* deposit) {
    if (!h) return PIC_EINVAL;
    if (mode) *mode = h->resident ? PIC_MODE_RESIDENT : PIC_MODE_STREAMING;
    if (threads) *threads = h->threads;
    if (per_thread) *per_thread = h->resident ? h->cluster : h->per_thread;
    if (grid_x) *grid_x = h->grid_x;
    if (smem_bytes) *smem_bytes = (int32_t)h->smem;
    if (fixed_bits) *fixed_bits = h->fixed_bits;
    if (deposit) *deposit = h->dep;
    return PIC_OK;
}

int64_t pic_kernel_launch_count(const pic_handle* h) { return h ? h->launches : 0; }

int pic_set_state_device(pic_handle* h, const void* xd, const void* vd) {
    if (!h || !xd || !vd) return fail(h, PIC_EINVAL, "null argument");
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaMemcpy2DAsync(h->x, (size_t)h->ld * h->esize, xd, (size_t)h->N * h->esize, (size_t)h->N * h->esize,
                            h->n_envs, cudaMemcpyDeviceToDevice, h->stream));
    CK(h, cudaMemcpy2DAsync(h->v, (size_t)h->ld * h->esize, vd, (size_t)h->N * h->esize, (size_t)h->N * h->esize,
                            h->n_envs, cudaMemcpyDeviceToDevice, h->stream));
    return init_fields(h);
}

int pic_set_state(pic_handle* h, const double* x, const double* v) {
    if (!h || !x || !v) return fail(h, PIC_EINVAL, "null argument");
    CK(h, cudaSetDevice(h->device));
    const size_t row = (size_t)h->N;
    if (!h->f32) {
        CK(h, cudaMemcpy2DAsync(h->x, (size_t)h->ld * 8, x, row * 8, row * 8, h->n_envs, cudaMemcpyHostToDevice, h->stream));
        CK(h, cudaMemcpy2DAsync(h->v, (size_t)h->ld * 8, v, row * 8, row * 8, h->n_envs, cudaMemcpyHostToDevice, h->stream));
    } else {
        int rc = ensure_stage64(h, (size_t)h->ld * h->n_envs);
        if (rc) return rc;
        const double* src[2] = {x, v};
        void* dst[2] = {h->x, h->v};
        for (int a = 0; a < 2; ++a) {
            CK(h, cudaMemcpy2DAsync(h->stage64, (size_t)h->ld * 8, src[a], row * 8, row * 8, h->n_envs,
                                    cudaMemcpyHostToDevice, h->stream));
            convert_kernel<double, float><<<h->sm_count * 4, 256, 0, h->stream>>>(h->stage64, (float*)dst[a],
                                                                                   (long long)h->ld * h->n_envs);
            h->launches++;
        }
    }
    return init_fields(h);
}

int pic_sample_state(pic_handle* h, int32_t kind, double a, double v0, double sigma, double A, int32_t n_mode,
                     uint64_t seed, int64_t global_offset, int64_t n_global, int64_t env_offset) {
    if (!h) return PIC_EINVAL;
    if (kind != 0 && kind != 1) return fail(h, PIC_EINVAL, "kind: 0 = bump-on-tail, 1 = two-stream");
    if (n_global <= 0) n_global = h->Ntotal;
    CK(h, cudaSetDevice(h->device));
    long long gx = (h->N + 255) / 256;
    dim3 grid((unsigned)(gx < 65535 ? gx : 65535), h->n_envs);
    if (h->f32) sample_kernel<float><<<grid, 256, 0, h->stream>>>((float*)h->x, (float*)h->v, h->N, h->ld, global_offset,
                                                                  n_global, kind, a, v0, sigma, A, n_mode, h->mc.L, seed, env_offset);
    else sample_kernel<double><<<grid, 256, 0, h->stream>>>((double*)h->x, (double*)h->v, h->N, h->ld, global_offset,
                                                            n_global, kind, a, v0, sigma, A, n_mode, h->mc.L, seed, env_offset);
    h->launches++;
    CK(h, cudaGetLastError());
    return init_fields(h);
}

int pic_get_state(pic_handle* h, double* x, double* v) {
    if (!h) return PIC_EINVAL;
    if (!h->have_state) return fail(h, PIC_ESTATE, "no state");
    CK(h, cudaSetDevice(h->device));
    const size_t row = (size_t)h->N;
    double* dst[2] = {x, v};
    void* src[2] = {h->x, h->v};
    for (int a = 0; a < 2; ++a) {
        if (!dst[a]) continue;
        if (!h->f32) {
            CK(h, cudaMemcpy2DAsync(dst[a], row * 8, src[a], (size_t)h->ld * 8, row * 8, h->n_envs, cudaMemcpyDeviceToHost, h->stream));
        } else {
            int rc = ensure_stage64(h, (size_t)h->ld * h->n_envs);
            if (rc) return rc;
            convert_kernel<float, double><<<h->sm_count * 4, 256, 0, h->stream>>>((const float*)src[a], h->stage64,
                                                                                   (long long)h->ld * h->n_envs);
            h->launches++;
            CK(h, cudaMemcpy2DAsync(dst[a], row * 8, h->stage64, (size_t)h->ld * 8, row * 8, h->n_envs, cudaMemcpyDeviceToHost, h->stream));
            CK(h, cudaStreamSynchronize(h->stream));
        }
    }
    CK(h, cudaStreamSynchronize(h->stream));
    return PIC_OK;
}

int pic_get_fields(pic_handle* h, double* n, double* E) {
    if (!h) return PIC_EINVAL;
    if (!h->have_state) return fail(h, PIC_ESTATE, "no state");
    size_t b = sizeof(double) * (size_t)h->M * h->n_envs;
    if (n) CK(h, cudaMemcpyAsync(n, h->n, b, cudaMemcpyDeviceToHost, h->stream));
    if (E) CK(h, cudaMemcpyAsync(E, h->E, b, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return PIC_OK;
}

int pic_get_density_fixed(pic_handle* h, uint64_t* rho, int32_t* fixed_bits) {
    if (!h) return PIC_EINVAL;
    if (!h->have_state) return fail(h, PIC_ESTATE, "no state");
    if (rho) CK(h, cudaMemcpyAsync(rho, h->rho[3], sizeof(uint64_t) * (size_t)h->M * h->n_envs, cudaMemcpyDeviceToHost, h->stream));
    if (fixed_bits) *fixed_bits = h->fixed_bits;
    CK(h, cudaStreamSynchronize(h->stream));
    return PIC_OK;
}

int pic_get_diag(pic_handle* h, double* diag) {
    if (!h || !diag) return PIC_EINVAL;
    if (!h->have_state) return fail(h, PIC_ESTATE, "no state");
    CK(h, cudaMemcpyAsync(diag, h->diag, sizeof(double) * DIAG_N * h->n_envs, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return PIC_OK;
}

int pic_get_diag_flags(pic_handle* h, double* diag, uint32_t* flags) {
    if (!h || !diag || !flags) return PIC_EINVAL;
    if (!h->have_state) return fail(h, PIC_ESTATE, "no state");
    CK(h, cudaMemcpyAsync(diag, h->diag, sizeof(double) * DIAG_N * h->n_envs, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(flags, h->err, sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return PIC_OK;
}

int pic_get_trace(pic_handle* h, double* trace, int32_t n_steps) {
    if (!h || !trace) return PIC_EINVAL;
    if (n_steps < 1 || n_steps > h->trace_steps) return fail(h, PIC_EINVAL, "n_steps exceeds the last call's step count");
    CK(h, cudaMemcpyAsync(trace, h->trace, sizeof(double) * DIAG_N * h->n_envs * (size_t)n_steps, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return PIC_OK;
}

int pic_get_cells(pic_handle* h, int32_t* il, double* wl, double* wr, double* Ep, double* wm) {
    if (!h) return PIC_EINVAL;
    if (!h->have_state) return fail(h, PIC_ESTATE, "no state");
    size_t n = (size_t)h->N * h->n_envs;
    int rc = ensure_stage64(h, 5 * n);          // [wl | wr | Ep | wm | il(int32)]
    if (rc) return rc;
    double* dwl = h->stage64; double* dwr = h->stage64 + n; double* dE = h->stage64 + 2 * n; double* dwm = h->stage64 + 3 * n;
    int* dil = (int*)(h->stage64 + 4 * n);
    long long gx = (h->N + 255) / 256;
    dim3 grid((unsigned)(gx < 4096 ? gx : 4096), h->n_envs);
#define PIC_CELLS(R, EX, IP) cells_kernel<R, EX, IP><<<grid, 256, 0, h->stream>>>((const R*)h->x, h->N, h->ld, h->mc, dil, dwl, dwr, dwm, h->E, dE)
    if (h->ip == IP_TSC) PIC_CELLS(double, false, IP_TSC);
    else if (h->f32) { if (h->exact_w) PIC_CELLS(float, true, IP_CIC); else PIC_CELLS(float, false, IP_CIC); }
    else             { if (h->exact_w) PIC_CELLS(double, true, IP_CIC); else PIC_CELLS(double, false, IP_CIC); }
#undef PIC_CELLS
    h->launches++;
    if (il) CK(h, cudaMemcpyAsync(il, dil, sizeof(int) * n, cudaMemcpyDeviceToHost, h->stream));
    if (wl) CK(h, cudaMemcpyAsync(wl, dwl, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    if (wr) CK(h, cudaMemcpyAsync(wr, dwr, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    if (Ep) CK(h, cudaMemcpyAsync(Ep, dE, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    if (wm) CK(h, cudaMemcpyAsync(wm, dwm, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return PIC_OK;
}

int pic_set_actuator_basis(pic_handle* h, const double* bc, const double* bs, int32_t m) {
    if (!h || !bc || !bs) return fail(h, PIC_EINVAL, "null argument");
    if (m != h->m || m < 1) return fail(h, PIC_EINVAL, "m must equal cfg.max_mode (> 0)");
    size_t b = sizeof(double) * (size_t)h->M * m;
    CK(h, cudaMemcpyAsync(h->bcos, bc, b, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(h->bsin, bs, b, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    h->have_basis = true;
    return PIC_OK;
}

int pic_step_mesh_device(pic_handle* h, const double* ext_dev, int32_t n_steps) {
    if (!h) return PIC_EINVAL;
    CK(h, cudaSetDevice(h->device));
    return step_device(h, ext_dev, nullptr, n_steps);
}

int pic_step_coeffs_device(pic_handle* h, const double* coeffs_dev, int32_t n_steps) {
    if (!h || !coeffs_dev) return fail(h, PIC_EINVAL, "null argument");
    CK(h, cudaSetDevice(h->device));
    return step_device(h, nullptr, coeffs_dev, n_steps);
}

int pic_step_mesh(pic_handle* h, const double* ext, int32_t n_steps) {
    if (!h) return PIC_EINVAL;
    CK(h, cudaSetDevice(h->device));
    if (ext) CK(h, cudaMemcpyAsync(h->ext, ext, sizeof(double) * (size_t)h->M * h->n_envs, cudaMemcpyHostToDevice, h->stream));
    return step_device(h, ext ? h->ext : nullptr, nullptr, n_steps);
}

int pic_step_coeffs(pic_handle* h, const double* coeffs, int32_t n_steps) {
    if (!h || !coeffs) return fail(h, PIC_EINVAL, "null argument");
    if (h->m < 1) return fail(h, PIC_ESTATE, "handle was created with max_mode == 0");
    if (n_steps < 1) return fail(h, PIC_EINVAL, "n_steps must be >= 1");
    CK(h, cudaSetDevice(h->device));
    long long need = (long long)n_steps * h->n_envs * 2 * h->m;
    if (need > h->coeffs_cap) {
        CK(h, cudaStreamSynchronize(h->stream));
        if (h->coeffs) cudaFree(h->coeffs);
        h->coeffs = nullptr; h->coeffs_cap = 0;
        CK(h, cudaMalloc(&h->coeffs, sizeof(double) * need));
        h->coeffs_cap = need;
    }
    CK(h, cudaMemcpyAsync(h->coeffs, coeffs, sizeof(double) * need, cudaMemcpyHostToDevice, h->stream));
    return step_device(h, nullptr, h->coeffs, n_steps);
}

int pic_set_reward(pic_handle* h, double alpha, double beta, double r_pe_n, double r_ie_n) {
    if (!h) return PIC_EINVAL;
    if (!(r_pe_n > 0) || !(r_ie_n > 0)) return fail(h, PIC_EINVAL, "reward normalisations must be positive");
    h->rw.alpha = alpha; h->rw.beta = beta; h->rw.r_pe_n = r_pe_n; h->rw.r_ie_n = r_ie_n;
    return PIC_OK;
}

int pic_enable_modes(pic_handle* h, int32_t n_modes) {
    if (!h) return PIC_EINVAL;
    if (n_modes < 0 || n_modes > MAX_MODES || n_modes >= h->M / 2)
        return fail(h, PIC_EINVAL, "n_modes must be in 0.." + std::to_string(MAX_MODES) + " and below N_mesh/2");
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    for (double** b : {&h->tw_cos, &h->tw_sin, &h->modes}) { if (*b) cudaFree(*b); *b = nullptr; }
    h->n_modes = n_modes;
    if (n_modes == 0) return PIC_OK;
    std::vector<double> c((size_t)h->M * n_modes), s((size_t)h->M * n_modes);
    for (int j = 0; j < h->M; ++j)
        for (int k = 0; k < n_modes; ++k) {
            // exact argument reduction: (j * (k+1)) mod M keeps the angle in [0, 2 pi)
            const long long jk = ((long long)j * (k + 1)) % h->M;
            const double th = 2.0 * 3.14159265358979323846 * (double)jk / (double)h->M;
            c[(size_t)j * n_modes + k] = cos(th); s[(size_t)j * n_modes + k] = sin(th);
        }
    size_t b = sizeof(double) * (size_t)h->M * n_modes;
    CK(h, cudaMalloc(&h->tw_cos, b)); CK(h, cudaMalloc(&h->tw_sin, b));
    CK(h, cudaMalloc(&h->modes, sizeof(double) * (size_t)h->n_envs * 2 * n_modes));
    CK(h, cudaMemcpy(h->tw_cos, c.data(), b, cudaMemcpyHostToDevice));
    CK(h, cudaMemcpy(h->tw_sin, s.data(), b, cudaMemcpyHostToDevice));
    CK(h, cudaMemset(h->modes, 0, sizeof(double) * (size_t)h->n_envs * 2 * n_modes));
    if (h->have_state) return init_fields(h);          // refresh so that the modes of the current state exist
    return PIC_OK;
}

int pic_get_modes(pic_handle* h, double* modes) {
    if (!h || !modes) return PIC_EINVAL;
    if (h->n_modes < 1) return fail(h, PIC_ESTATE, "pic_enable_modes has not been called");
    if (!h->have_state) return fail(h, PIC_ESTATE, "no state");
    CK(h, cudaMemcpyAsync(modes, h->modes, sizeof(double) * (size_t)h->n_envs * 2 * h->n_modes, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return PIC_OK;
}

int pic_get_mode_trace(pic_handle* h, double* out, int32_t n_steps) {
    if (!h || !out) return PIC_EINVAL;
    if (h->n_modes < 1) return fail(h, PIC_ESTATE, "pic_enable_modes has not been called");
    if (n_steps < 1 || n_steps > h->trace_steps) return fail(h, PIC_EINVAL, "n_steps exceeds the last call's step count");
    CK(h, cudaMemcpyAsync(out, h->mode_trace, sizeof(double) * (size_t)n_steps * h->n_envs * 2 * h->n_modes,
                          cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return PIC_OK;
}

static int phase_hist_run(pic_handle* h) {
    const int nb = h->ph_nb;
    CK(h, cudaMemsetAsync(h->ph_counts, 0, sizeof(unsigned) * (size_t)h->n_envs * nb * nb, h->stream));
    long long gx = (h->N + 255) / 256;
    dim3 grid((unsigned)(gx < 2048 ? gx : 2048), h->n_envs);
    if (h->f32) phase_hist_kernel<float><<<grid, 256, 0, h->stream>>>((const float*)h->x, (const float*)h->v, h->N, h->ld, nb,
                                                                      h->mc.L, h->ph_vmin, h->ph_vmax, h->ph_counts);
    else phase_hist_kernel<double><<<grid, 256, 0, h->stream>>>((const double*)h->x, (const double*)h->v, h->N, h->ld, nb,
                                                                h->mc.L, h->ph_vmin, h->ph_vmax, h->ph_counts);
    h->launches++;
    CK(h, cudaGetLastError());
    return PIC_OK;
}

int pic_phase_hist_config(pic_handle* h, double vmin, double vmax, int32_t nbins) {
    if (!h) return PIC_EINVAL;
    if (nbins < 1 || nbins > 4096 || !(vmax > vmin)) return fail(h, PIC_EINVAL, "need 1 <= nbins <= 4096 and vmax > vmin");
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    if (h->ph_counts) cudaFree(h->ph_counts);
    if (h->ph_feq) cudaFree(h->ph_feq);
    if (h->ph_kl) cudaFree(h->ph_kl);
    h->ph_counts = nullptr; h->ph_feq = nullptr; h->ph_kl = nullptr;
    h->ph_nb = nbins; h->ph_vmin = vmin; h->ph_vmax = vmax;
    CK(h, cudaMalloc(&h->ph_counts, sizeof(unsigned) * (size_t)h->n_envs * nbins * nbins));
    CK(h, cudaMalloc(&h->ph_feq, sizeof(double) * (size_t)nbins * nbins));
    CK(h, cudaMalloc(&h->ph_kl, sizeof(double) * (size_t)h->n_envs));
    CK(h, cudaMemset(h->ph_feq, 0, sizeof(double) * (size_t)nbins * nbins));
    return PIC_OK;
}

int pic_phase_hist(pic_handle* h, uint32_t* counts) {
    if (!h) return PIC_EINVAL;
    if (h->ph_nb < 1) return fail(h, PIC_ESTATE, "pic_phase_hist_config has not been called");
    if (!h->have_state) return fail(h, PIC_ESTATE, "no state");
    CK(h, cudaSetDevice(h->device));
    int rc = phase_hist_run(h);
    if (rc) return rc;
    if (counts) {
        CK(h, cudaMemcpyAsync(counts, h->ph_counts, sizeof(unsigned) * (size_t)h->n_envs * h->ph_nb * h->ph_nb,
                              cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
    }
    return PIC_OK;
}

int pic_set_feq(pic_handle* h, const double* feq) {
    if (!h || !feq) return PIC_EINVAL;
    if (h->ph_nb < 1) return fail(h, PIC_ESTATE, "pic_phase_hist_config has not been called");
    CK(h, cudaMemcpyAsync(h->ph_feq, feq, sizeof(double) * (size_t)h->ph_nb * h->ph_nb, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return PIC_OK;
}

int pic_kl_divergence(pic_handle* h, double* kl) {
    if (!h || !kl) return PIC_EINVAL;
    if (h->ph_nb < 1) return fail(h, PIC_ESTATE, "pic_phase_hist_config has not been called");
    if (!h->have_state) return fail(h, PIC_ESTATE, "no state");
    CK(h, cudaSetDevice(h->device));
    int rc = phase_hist_run(h);
    if (rc) return rc;
    const int nb = h->ph_nb;
    const double dx = h->mc.L / nb, dv = (h->ph_vmax - h->ph_vmin) / nb;
    const double scale = h->mc.n0 / dx / dv / (double)h->Ntotal;            // objective.py:13
    kl_kernel<<<h->n_envs, 256, 0, h->stream>>>(h->ph_counts, h->ph_feq, nb, scale, dx * dv, h->ph_kl);
    h->launches++;
    CK(h, cudaMemcpyAsync(kl, h->ph_kl, sizeof(double) * (size_t)h->n_envs, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return PIC_OK;
}

int pic_refresh_fields(pic_handle* h) {
    if (!h) return PIC_EINVAL;
    if (!h->have_state) return fail(h, PIC_ESTATE, "no state");
    CK(h, cudaSetDevice(h->device));
    return init_fields(h);
}

int pic_sync(pic_handle* h) {
    if (!h) return PIC_EINVAL;
    CK(h, cudaStreamSynchronize(h->stream));
    return PIC_OK;
}

int pic_get_error_flags(pic_handle* h, uint32_t* flags) {
    if (!h || !flags) return PIC_EINVAL;
    CK(h, cudaMemcpyAsync(flags, h->err, sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return PIC_OK;
}

int pic_clear_error_flags(pic_handle* h) {
    if (!h) return PIC_EINVAL;
    CK(h, cudaMemsetAsync(h->err, 0, sizeof(unsigned), h->stream));
    return PIC_OK;
}

int pic_get_device_views(pic_handle* h, pic_device_views* out) {
    if (!h || !out) return PIC_EINVAL;
    out->x = h->x; out->v = h->v; out->ld = h->ld; out->n = h->n; out->E_mesh = h->E; out->diag = h->diag;
    out->elem_size = h->esize;
    return PIC_OK;
}

// ------------------------------------------------------------------------------------------------ sharding
int pic_nccl_unique_id(char* out128) {
    if (!out128) return PIC_EINVAL;
    if (!nccl_api().ok) return fail(nullptr, PIC_ENCCL, "libnccl not available in this process");
    nccl_uid id;
    if (nccl_api().get_uid(&id) != 0) return fail(nullptr, PIC_ENCCL, "ncclGetUniqueId failed");
    memcpy(out128, id.internal, 128);
    return PIC_OK;
}

int pic_comm_init_rank(pic_handle* h, const char* id128, int32_t rank, int32_t world) {
    if (!h || !id128) return PIC_EINVAL;
    if (h->resident) return fail(h, PIC_EUNSUPPORTED, "particle sharding needs streaming mode");
    if (!nccl_api().ok) return fail(h, PIC_ENCCL, "libnccl not available in this process");
    CK(h, cudaSetDevice(h->device));
    nccl_uid id; memcpy(id.internal, id128, 128);
    void* comm = nullptr;
    int r = nccl_api().init_rank(&comm, world, id, rank);
    if (r != 0) return fail(h, PIC_ENCCL, std::string("ncclCommInitRank: ") + (nccl_api().errstr ? nccl_api().errstr(r) : "error"));
    h->comm = comm; h->own_comm = true; h->rank = rank; h->world = world;
    return PIC_OK;
}

int pic_comm_init(pic_handle* h, void* comm, int32_t rank, int32_t world) {
    if (!h || !comm) return PIC_EINVAL;
    if (h->resident) return fail(h, PIC_EUNSUPPORTED, "particle sharding needs streaming mode");
    if (!nccl_api().ok) return fail(h, PIC_ENCCL, "libnccl not available in this process");
    h->comm = comm; h->own_comm = false; h->rank = rank; h->world = world;
    return PIC_OK;
}

int64_t pic_comm_exchange_words(const pic_handle* h, int32_t world) {
    return h ? (int64_t)COMM_SETS * world * comm_slot_len(h) : 0;
}

int pic_comm_init_peer(pic_handle* h, int32_t rank, int32_t world, void* const* exch_ptrs, void* const* flag_ptrs,
                       int64_t exch_words) {
    if (!h || !exch_ptrs || !flag_ptrs) return PIC_EINVAL;
    if (h->resident || h->n_envs != 1) return fail(h, PIC_EUNSUPPORTED, "the fused exchange needs streaming mode with one env");
    if (world < 2 || world > COMM_MAX_WORLD || rank < 0 || rank >= world)
        return fail(h, PIC_EINVAL, "world must be 2.." + std::to_string(COMM_MAX_WORLD));
    if (exch_words < pic_comm_exchange_words(h, world)) return fail(h, PIC_EINVAL, "exchange buffer too small");
    for (int r = 0; r < world; ++r) {
        if (!exch_ptrs[r] || !flag_ptrs[r]) return fail(h, PIC_EINVAL, "null peer pointer");
        h->exch[r] = (unsigned long long*)exch_ptrs[r]; h->cflags[r] = (unsigned long long*)flag_ptrs[r];
    }
    h->rank = rank; h->world = world; h->fused = true; h->seq = 0; h->seq_state = 0; h->exch_mc = nullptr;
    return configure_gather(h);
}

int pic_comm_set_multicast(pic_handle* h, void* exch_multicast) {
    if (!h) return PIC_EINVAL;
    if (!h->fused) return fail(h, PIC_ESTATE, "pic_comm_init_peer has not been called");
    cudaStreamSynchronize(h->stream);
    h->exch_mc = (unsigned long long*)exch_multicast;
    return PIC_OK;
}

int pic_set_stage_actuation(pic_handle* h, const double* ext_dev, const double* coeffs_dev) {
    if (!h) return PIC_EINVAL;
    h->stage_ext = ext_dev; h->stage_coeffs = coeffs_dev;
    return PIC_OK;
}

int pic_run_stage(pic_handle* h, int32_t stage) {
    if (!h) return PIC_EINVAL;
    if (h->resident) return fail(h, PIC_EUNSUPPORTED, "pic_run_stage needs streaming mode");
    CK(h, cudaSetDevice(h->device));
    if (stage == 4 || stage == -1) {
        int rc = run_stage(h, stage, nullptr, nullptr, nullptr);
        if (!rc && stage == 4) h->have_state = true;
        return rc;
    }
    return run_stage(h, stage, h->stage_ext, h->stage_coeffs, nullptr);
}

int pic_stage_density(pic_handle* h, int32_t stage, uint64_t** rho_dev) {
    if (!h || !rho_dev) return PIC_EINVAL;
    int idx = stage == -1 ? 3 : stage;          // -1 and 3: [S][W0], 2 * n_envs * n_mesh values (state + next stage 0)
    if (idx < 0 || idx > 3) return fail(h, PIC_EINVAL, "stage must be -1..3");
    *rho_dev = (uint64_t*)h->rho[idx];
    return PIC_OK;
}

}  // extern "C"
