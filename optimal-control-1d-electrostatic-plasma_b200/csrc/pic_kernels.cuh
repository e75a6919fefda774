// Kernels of the PIC hot path (sm_100a).  See pic_device.cuh for the per-particle arithmetic.
//
//   push_stream_kernel   large-N "streaming" mode: three launches per env step (KICK0, KICK, FINAL; the deposit of the
//                        drift-only first Yoshida sub-stage rides along with the pass that produces the state, and the
//                        stage-1 pass redoes the drift on load).  Persistent CTAs
//                        stream the SoA particle arrays once per pass (16-byte loads/stores), every CTA rebuilds the
//                        mesh field from the previous sub-stage's fixed-point density in its prologue (one block scan,
//                        no separate Poisson launch), gathers/kicks/drifts, and deposits into shared-memory privatised
//                        histograms that are flushed with 64-bit global reductions.
//   env_step_resident_kernel
//                        small-N / batched mode: one CTA per env, particles live in shared memory for the whole
//                        launch, all four sub-stages (and any number of env steps) run inside one launch.
//   field_finalize_kernel
//                        state field after a streaming step: density, self-consistent E, energies.
#pragma once
#include <type_traits>

#include "pic_device.cuh"

#ifndef PIC_PREFETCH_TILES
#define PIC_PREFETCH_TILES 2      // L2 prefetch distance of the streaming kernel, in tiles of its own CTA (0: off)
#endif
#ifndef PIC_PREFETCH_TILES_HBM
#define PIC_PREFETCH_TILES_HBM 0  // the same for the passes that are purely HBM-bound (stage 1, init)
#endif

namespace pic {

// Sub-stages of the streaming mode.  Stage 0 of every step (d == 0, a pure drift, integration.py:71) is deposited
// ahead of time by the kernel that produces the state it starts from (MODE_FINAL / MODE_INIT, see next_stage0) and its
// drift is redone on load by the stage-1 kernel, so an env step is three passes over the particles of 32 bytes per
// particle each: KICK0, KICK, FINAL.
constexpr int MODE_KICK = 1;    // stage 2: kick + drift (integration.py:72-73); stores v ONLY -- its drift is redone on load by MODE_FINAL
constexpr int MODE_FINAL = 2;   // stage 3: the stage-2 drift redone on load, kick + drift + state wrap (pic.py:139) + kinetic sums,
                                // then the stage-0 deposit of the next step
constexpr int MODE_INIT = 3;    // no motion: wrap in place + deposit (pic.py:76 / util.py:51), then the stage-0 deposit of the next step
constexpr int MODE_KICK0 = 4;   // stage 1: the stage-0 drift redone on load (its deposit happened a pass ago), then kick + drift

// ------------------------------------------------------------------ fused density exchange over NVLink
// Particle-sharded mode without a collective library in the step loop.  Every rank owns an exchange buffer that all
// peers can write (symmetric / peer-mapped memory): SETS slot sets x `world` slots x `slot_len` 64-bit words, plus
// one flag word per source rank.  The LAST CTA of a push kernel (ticket counter) copies the rank's finished partial
// density -- and, after a state-producing pass, its kinetic sums -- into slot[set][rank] of EVERY rank with peer
// stores, fences, and then publishes the exchange number `seq` in flag[rank] on every rank.  The next kernel's
// prologue waits until all `world` flags have reached `seq` and reads the density as the sum over the slots in rank
// order: an integer sum, identical on every rank.  The transfer therefore overlaps the tail of the producing kernel
// and the head of the consuming one; there is no separate reduction kernel and no host involvement.
// Two slot sets would suffice (a rank can produce exchange s+2 only after every rank has consumed exchange s); four
// are used.  Waits are bounded: a missing peer raises ERR_COMM_TIMEOUT instead of hanging the GPU.
constexpr int COMM_MAX_WORLD = 8;
constexpr int COMM_SETS = 4;
constexpr unsigned ERR_COMM_TIMEOUT = 4u;

struct CommArgs {
    int world, rank;                                   // world <= 1: disabled
    int slot_len;                                      // words per slot (2 * n_envs * M + 2 * n_envs)
    unsigned long long* exch[COMM_MAX_WORLD];          // exchange buffer of every rank (peer pointers); [rank] is local
    unsigned long long* flags[COMM_MAX_WORLD];         // flag array of every rank; this rank writes entry [rank] of each
    unsigned long long* exch_mc;                       // NVLS multicast mapping of the exchange buffers (one store lands in
                                                       // every rank's copy), or nullptr: one peer store per rank
    unsigned long long seq_in;                         // exchange to consume in the prologue (0: none)
    unsigned long long seq_out;                        // exchange produced by this kernel (0: none)
    int in_offset;                                     // word offset of the density to consume inside a slot
    int out_words;                                     // words this kernel publishes
    unsigned* ticket;                                  // CTA completion counter (zero between kernels)
    const unsigned long long* out_src;                 // local finished partial density, out_words contiguous words
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// one store to a multicast (NVLS) address: the switch replicates it into the copy of every rank of the group
__device__ __forceinline__ void multimem_st(unsigned long long* p, unsigned long long v) {
    asm volatile("multimem.st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// every thread of the CTA calls this; returns after all ranks have published exchange `seq`.
// Returns true (to every thread) when a peer did not arrive within the bound: the caller must then leave the
// particle state alone -- the densities in the slots are incomplete, and a step built on them would overwrite x and v
// with garbage.  The sticky ERR_COMM_TIMEOUT flag makes the host classes raise on their next read.
#ifndef PIC_COMM_SPIN_LIMIT
#define PIC_COMM_SPIN_LIMIT (1ll << 25)              // x (64 ns sleep + one system-scope load): a few seconds
#endif
__device__ __forceinline__ bool comm_wait(const CommArgs& c, unsigned long long seq, unsigned* err) {
    int timed_out = 0;
    if (threadIdx.x < c.world) {
        const unsigned long long* f = c.flags[c.rank] + threadIdx.x;
        long long spins = 0;
        while (ld_acquire_sys(f) < seq) {
            if (++spins > PIC_COMM_SPIN_LIMIT || ((spins & 0xfff) == 0 && (__ldcg(err) & ERR_COMM_TIMEOUT))) {
                atomicOr(err, ERR_COMM_TIMEOUT);       // never hang the GPU; another CTA's verdict is taken over at once
                timed_out = 1;
                break;
            }
            __nanosleep(64);
        }
    }
    return __syncthreads_or(timed_out) != 0;
}

struct PeerSumRho {                                   // density = sum over ranks of their published partial density
    const unsigned long long* slots;                  // slot[set][0] + in_offset of the local exchange buffer
    int slot_len, world;
    __device__ __forceinline__ unsigned long long operator()(int j) const {
        unsigned long long s = 0;
        for (int r = 0; r < world; ++r) s += __ldcg(slots + (size_t)r * slot_len + j);
        return s;
    }
};
__device__ __forceinline__ const unsigned long long* comm_in_slots(const CommArgs& c, unsigned long long seq) {
    return c.exch[c.rank] + (size_t)(seq % COMM_SETS) * c.world * c.slot_len + c.in_offset;
}

struct ActuatorArgs {
    const double* ext;       // mesh mode: [n_envs][M] (device) or nullptr
    const double* coeffs;    // coefficient mode: [n_envs][2m] (device) or nullptr
    const double* bcos;      // [M][m]
    const double* bsin;      // [M][m]
    int m;
};

struct StreamArgs {
    MeshConst mc;
    PartConsts pcs;                    // per-particle constants of both precisions (host-computed)
    void* x;                           // [n_envs][ld] particle positions (R): read and written in place
    void* v;                           // [n_envs][ld] particle velocities (R)
    long long N, ld;
    const unsigned long long* rho_in;  // [n_envs][M] density of the previous sub-stage (MODE_KICK / MODE_FINAL)
    unsigned long long* rho_out;       // [n_envs][M] must be zero on entry
    unsigned long long* rho_next;      // MODE_FINAL / MODE_INIT: density of x1 (must be zero on entry)
    unsigned long long* rho_zero;      // [n_envs][M] or nullptr: cleared for a later sub-stage
    unsigned long long* rho_zero2;     // a second buffer to clear (overlapped finalize: see step_device), or nullptr
    ActuatorArgs act;
    double c, d;
    double c_next;                     // c0 of the Yoshida scheme (MODE_FINAL / MODE_INIT: stage 0 of the next step)
    double c_pre;                      // drift redone on load: MODE_KICK0: c0 (stage 0), MODE_FINAL: c2 (stage 2, whose
                                       // pass stored only v) -- the same three roundings, so the positions are bit-identical
    CommArgs comm;                     // fused exchange over peer memory (world <= 1: off)
    double* partial;                   // [n_envs][gridDim.x][2] per-CTA sum v^2, sum v (MODE_FINAL / MODE_INIT)
    unsigned* err;
    cudaTextureObject_t table_tex;     // TEXG kernels: linear texture over the [n_envs][M] pair table of this sub-stage
};

// Shared-memory plan of one CTA.  SEPARATE_D: the prefix-sum scratch gets its own region (needed when the density
// is read from the shared histogram itself, i.e. the resident kernel); otherwise it aliases the histogram, which is
// idle while the field is rebuilt from the global density.
__host__ __device__ constexpr size_t hist_region_bytes(int M, int ip) {      // max over deposit flavours, 16-aligned
    return (((size_t)M * (ip == IP_TSC ? 20 : 12) + 16 + 15) & ~(size_t)15);
}

template <typename R>
__host__ __device__ constexpr size_t smem_plan_bytes(int M, int threads, bool separate_d, int ip = IP_CIC,
                                                     bool second_hist = false, bool table = true) {
    return (table ? (size_t)M * 2 * sizeof(R) : 0)   // gather pair table (none: the kernel gathers through a texture)
         + hist_region_bytes(M, ip) * (second_hist ? 2 : 1)   // histogram(s)
         + (separate_d ? (size_t)M * 8 : 0)          // D_s
         + (size_t)(field_scratch_doubles(threads) + threads / 32 + 2) * 8;   // field / reduction scratch
}

// byte offsets of the regions inside the dynamic shared memory of a CTA (the gather table sits at offset 0)
struct SmemOffsets { unsigned hist, hist2, D, red, x, v, ext; };
template <typename R>
__host__ __device__ inline SmemOffsets smem_offsets(int M, int threads, bool separate_d, int ip = IP_CIC,
                                                    bool second_hist = false, long long n_resident = 0, bool table = true) {
    SmemOffsets o;
    size_t b = table ? (size_t)M * 2 * sizeof(R) : 0;
    o.hist = (unsigned)b;       b += hist_region_bytes(M, ip);
    o.hist2 = (unsigned)b;      if (second_hist) b += hist_region_bytes(M, ip);
    o.D = separate_d ? (unsigned)b : o.hist;
    if (separate_d) b += (size_t)M * 8;
    o.red = (unsigned)b;
    o.x = (unsigned)smem_plan_bytes<R>(M, threads, separate_d, ip, second_hist, table);   // resident particle state
    o.v = o.x + (unsigned)((n_resident + 1) / 2 * 2 * sizeof(R));
    o.ext = o.v + (unsigned)((n_resident + 1) / 2 * 2 * sizeof(R));                    // resident: E_ext of the step, M doubles
    return o;
}

template <typename R>
struct SmemLayout {
    void* hist; void* hist2; typename PairT<R>::type* E_s; double* D_s; double* red;
    __device__ __forceinline__ SmemLayout(unsigned char* base, const SmemOffsets& o) {
        E_s = (typename PairT<R>::type*)base;
        hist = base + o.hist; hist2 = base + o.hist2; D_s = (double*)(base + o.D); red = (double*)(base + o.red);
    }
    __device__ __forceinline__ SmemLayout(unsigned char* base, int M, bool separate_d, int ip = IP_CIC,
                                          bool second_hist = false, bool table = true)
        : SmemLayout(base, smem_offsets<R>(M, 0, separate_d, ip, second_hist, 0, table)) {}
};

struct GlobalRho {
    const unsigned long long* p;
    __device__ __forceinline__ unsigned long long operator()(int j) const { return __ldcg(p + j); }
};
template <typename H> struct SharedRho {
    const H* h; long long one;
    __device__ __forceinline__ unsigned long long operator()(int j) const { return h->get(j, one); }
};

// what the particles of env `env` see on top of the self-consistent field
__device__ __forceinline__ ExtSrc stage_ext(const ActuatorArgs& act, int env, int M) {
    ExtSrc e;
    e.ext = act.ext ? act.ext + (size_t)env * M : nullptr;
    e.coeff = act.coeffs ? act.coeffs + (size_t)env * 2 * act.m : nullptr;
    e.bcos = act.bcos; e.bsin = act.bsin; e.m = act.m;
    return e;
}

template <int DEP, int IP, typename R = double> struct HistSel { using type = Hist<DEP>; };
template <int DEP, typename R> struct HistSel<DEP, IP_TSC, R> { using type = HistTSC; };
template <> struct HistSel<DEP_SPLIT32, IP_CIC, float> { using type = Hist<DEP_SPLIT32_RARE>; };   // float32: fixed_bits <= 24

// TEXG: the gather goes through the texture pipe (TexTable) instead of a shared-memory table.  The table was written to
// global memory by field_table_kernel, launched between the passes, so these kernels have no field prologue and no
// table in shared memory: the L1 that the smaller shared-memory carve-out frees holds the 64 KB table.
// The pass itself is a device function so that the cooperative step kernel (step_coop_kernel below) can run the three
// passes of an env step back to back inside one launch.  COOP: called from that kernel -- the CTA's rank among the
// worker CTAs and their number come as arguments (the grid holds one more CTA, which does the finalize), there is no
// programmatic dependent launch, and the actuation of the step is passed separately (it changes from step to step
// while `a` stays the kernel parameter, whose constants the hot loop reads straight from the constant bank).
template <typename R, int THREADS, int UNROLL, int MODE, int DEP, bool EXACT_W, int IP = IP_CIC, bool TEXG = false, bool COOP = false>
__device__ __forceinline__ void push_stream_body(const StreamArgs& a, const ActuatorArgs& act, unsigned coop_bid, unsigned coop_nb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned bid = COOP ? coop_bid : blockIdx.x, nb = COOP ? coop_nb : gridDim.x;
    using V = typename RT<R>::vec;
    constexpr int VEC = RT<R>::VEC;
    static_assert(MODE == MODE_KICK || MODE == MODE_KICK0 || MODE == MODE_FINAL || MODE == MODE_INIT, "unknown sub-stage");
    constexpr bool KICK = (MODE != MODE_INIT);
    constexpr bool SUMS = (MODE == MODE_FINAL || MODE == MODE_INIT);      // these also deposit stage 0 of the next step
    const int tid = threadIdx.x, env = blockIdx.y, M = a.mc.M;
    SmemLayout<R> sm(smem_raw, M, false, IP, SUMS, !TEXG);
    using H = typename HistSel<DEP, IP, R>::type;
    H hist; hist.init(sm.hist, M);
    H hist_next; hist_next.init(sm.hist2, M);
    #ifdef PIC_DEVICE_PARTCONST                          // experiment: recompute the constants in registers (round-1 behaviour)
    const PartConst<R> pc = make_part_const<R>(a.mc);
#else
    const PartConst<R>& pc = part_const<R>(a.pcs);
#endif
#ifndef PIC_NO_PDL
    if constexpr (!COOP) {
        griddep_launch_dependents();                    // the next kernel of the step may start launching behind this one
        griddep_wait();                                 // ... and this one touches memory only after its predecessors are done
    }
#endif

    const bool fused = a.comm.world > 1;
    bool dead = false;                                  // fused exchange timed out: leave the particle state untouched
    if (KICK && !TEXG) {                                // D_s aliases the histogram: solve first, then clear
        const ExtSrc ext = stage_ext(act, env, M);
        if (fused) {
            dead = comm_wait(a.comm, a.comm.seq_in, a.err);
            PeerSumRho rho{comm_in_slots(a.comm, a.comm.seq_in), a.comm.slot_len, a.comm.world};
            block_field<R, THREADS, false, true>(rho, sm.E_s, sm.D_s, sm.red, a.mc, ext, nullptr, nullptr, 0.0, 0.0, [] {});
        } else {
            GlobalRho rho{a.rho_in + (size_t)env * M};
            block_field<R, THREADS, false, true>(rho, sm.E_s, sm.D_s, sm.red, a.mc, ext, nullptr, nullptr, 0.0, 0.0, [] {});
        }
    }
    // Texture route under the fused exchange: the consumer of the peers' slots was field_table_kernel.  If its wait timed
    // out it left the sticky flag, and this pass must leave the particle state alone: the whole CTA returns at once
    // (nothing is published either, so the peers time out on their next wait and do the same; the flag is fatal for the
    // handle anyway -- the host classes raise on their next read).  Checked here, once: keeping a `dead` flag alive across
    // the hot loop, or re-reading it after the loop, made ptxas spill in these kernels, which sit at the 64-register limit.
    if constexpr (KICK && TEXG) {
        if (fused && (__ldcg(a.err) & ERR_COMM_TIMEOUT) != 0) return;
    }
    hist.zero(tid, THREADS);
    if (SUMS) hist_next.zero(tid, THREADS);
    __syncthreads();
    if (a.rho_zero) {
        unsigned long long* z = a.rho_zero + (size_t)env * M;
        for (int j = bid * THREADS + tid; j < M; j += nb * THREADS) z[j] = 0ull;
    }
    if (a.rho_zero2) {
        unsigned long long* z = a.rho_zero2 + (size_t)env * M;
        for (int j = bid * THREADS + tid; j < M; j += nb * THREADS) z[j] = 0ull;
    }

    R* xe = (R*)a.x + (size_t)env * a.ld;
    R* ve = (R*)a.v + (size_t)env * a.ld;
    V* xv = (V*)xe;
    V* vv = (V*)ve;
    const long long nvec = a.N / VEC;
    // c_pre: MODE_KICK0 / MODE_FINAL redo on load the drift whose position the previous pass did not store;
    // c_next: MODE_FINAL / MODE_INIT drift the state they produce for the next step's stage-0 deposit
#ifdef PIC_SCHED96      // experiment: the 96 B schedule (stage 2 stores x, MODE_FINAL loads it)
    constexpr bool REDRIFT = (MODE == MODE_KICK0);
    constexpr bool STORE_X = true;
#else
    constexpr bool REDRIFT = (MODE == MODE_KICK0 || MODE == MODE_FINAL);
    constexpr bool STORE_X = (MODE != MODE_KICK);       // stage 2 leaves x alone: MODE_FINAL recomputes x3 from (x2, v2)
#endif
    const R cc = (R)a.c, dd = (R)a.d, cpre = (R)a.c_pre, cnext = (R)a.c_next;
    unsigned err = 0;
    double s2 = 0.0, s1 = 0.0;

    // x, v: updated in place.  probe: this particle refreshes the tile's aggregation hint (deposit_hinted)
    bool agg = false;
    const TexTable<R> table{a.table_tex, env * M};
    auto one = [&](R& x, R& v, auto full_warp, auto probe) {
        constexpr bool FW = decltype(full_warp)::value, PROBE = decltype(probe)::value;
        if (REDRIFT) x = drift<R>(x, v, cpre, pc);      // stage 0 / stage 2 drift (its deposit was done a pass ago)
        if constexpr (TEXG)
            particle_substage<R, IP, KICK, MODE != MODE_INIT, EXACT_W, FW, H, PROBE>(x, v, hist, table, cc, dd, pc, a.mc, SUMS,
                                                                                     err, &agg);
        else
            particle_substage<R, IP, KICK, MODE != MODE_INIT, EXACT_W, FW, H, PROBE>(x, v, hist, sm.E_s, cc, dd, pc, a.mc, SUMS,
                                                                                     err, &agg);
        if (SUMS) {
            s2 += (double)v * (double)v; s1 += (double)v;
            next_stage0<R, IP, EXACT_W, FW>(x, v, hist_next, cnext, pc, a.mc, err, &agg);
        }
    };
    // float32, CIC: two particles per floating-point instruction (f32x2 in pic_device.cuh); full tiles only.  px, pv point
    // at two consecutive particles of a vector.  Same bits as two calls of `one`: the packed lanes round like the scalar
    // instructions, a pair with a particle that needs the careful path is redone by the scalar code, the kinetic sums are
    // accumulated in particle order, and the deposits are integers.
#ifdef PIC_NO_F32X2
    constexpr bool PAIRS = false;
#else
    // stages 1 and 2 only: measured at N = 1e9, 3.31 -> 3.08 ms and 3.15 -> 2.68 ms; the stage-3 pass (two more positions per
    // particle, both particles' careful blocks inlined) got slower with pairs (5.39 -> 7.65 ms) and keeps the scalar code
#ifdef PIC_F32X2_FINAL
    constexpr bool PAIRS = std::is_same<R, float>::value && IP == IP_CIC && !EXACT_W && KICK && !TEXG;
#else
    constexpr bool PAIRS = std::is_same<R, float>::value && IP == IP_CIC && !EXACT_W && KICK && !TEXG && !SUMS;
#endif
#endif
    auto pair = [&](R* px, R* pv, auto probe) {
        if constexpr (PAIRS) {
            constexpr bool PROBE = decltype(probe)::value;
            float2 x = make_float2(px[0], px[1]);
            const float2 v = make_float2(pv[0], pv[1]);
            if (REDRIFT) x = f32x2::drift(x, v, cpre, pc);
            float2 xn, vn, wr, wr1;
            int il[2], jl[2] = {0, 0};
            long long Wa[2], Wn[2] = {0, 0}, Wb;
            bool slow[2], slown[2] = {false, false};
            f32x2::particle_fast(x, v, xn, vn, il[0], il[1], wr, sm.E_s, cc, dd, pc, M, slow[0], slow[1]);
            if (SUMS) {
                const float2 x1 = f32x2::drift(xn, vn, cnext, pc);
                float2 f1;
                f32x2::fast_cell(x1, pc, M, jl[0], jl[1], f1, slown[0], slown[1]);
                wr1 = f32x2::weight_r(x1, f1, pc);
            }
            R xo[2] = {xn.x, xn.y}, vo[2] = {vn.x, vn.y};
            Wa[0] = fix_weight((double)wr.x, a.mc.fix_scale); Wa[1] = fix_weight((double)wr.y, a.mc.fix_scale);
            if (SUMS) { Wn[0] = fix_weight((double)wr1.x, a.mc.fix_scale); Wn[1] = fix_weight((double)wr1.y, a.mc.fix_scale); }
            if (__builtin_expect(slow[0] | slow[1] | slown[0] | slown[1], 0)) {      // rare; no warp-wide step inside
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    if (slow[k]) {                      // the whole sub-stage of this particle with the reference semantics
                        R xs = px[k], vs = pv[k];
                        if (REDRIFT) xs = drift<R>(xs, vs, cpre, pc);
                        particle_careful<R, IP, true, true, EXACT_W>(xs, vs, il[k], Wa[k], Wb, sm.E_s, cc, dd, pc, a.mc, SUMS, err);
                        xo[k] = xs; vo[k] = vs;
                        if (SUMS) next_stage0_compute<R, IP, EXACT_W>(xs, vs, cnext, pc, a.mc, err, jl[k], Wn[k], Wb);
                    } else if (SUMS && slown[k]) {      // only the stage-0 position of the next step
                        next_stage0_compute<R, IP, EXACT_W>(xo[k], vo[k], cnext, pc, a.mc, err, jl[k], Wn[k], Wb);
                    }
                }
            }
            __syncwarp();                               // reconverge before the warp-wide deposits
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (k == 0) deposit_hinted<IP, PROBE>(hist, il[k], Wa[k], 0, a.mc.fix_one, agg);
                else deposit_hinted<IP, false>(hist, il[k], Wa[k], 0, a.mc.fix_one, agg);
                if (SUMS) {
                    s2 += (double)vo[k] * (double)vo[k]; s1 += (double)vo[k];
                    deposit_hinted<IP, false>(hist_next, jl[k], Wn[k], 0, a.mc.fix_one, agg);
                }
                px[k] = xo[k]; pv[k] = vo[k];
            }
        }
    };
    auto vec_pair = [&](long long i, auto full_warp) {      // one 16-byte vector of x and of v
        V xq = ld_stream(xv + i), vq = ld_stream(vv + i);
        R* px = reinterpret_cast<R*>(&xq);
        R* pv = reinterpret_cast<R*>(&vq);
#pragma unroll
        for (int e = 0; e < VEC; ++e) one(px[e], pv[e], full_warp, std::false_type{});
        if (STORE_X) st_stream(xv + i, xq);
        if (KICK) st_stream(vv + i, vq);
    };

    // flush the CTA-private histogram(s): integer sums are associative, so the result does not depend on CTA order
    unsigned long long* out = a.rho_out + (size_t)env * M;
    unsigned long long* outn = SUMS ? a.rho_next + (size_t)env * M : nullptr;
    auto flush = [&]() {
        for (int j = tid; j < M; j += THREADS) {
            unsigned long long val = hist.get(j, a.mc.fix_one);
            if (val) atomicAdd(out + j, val);
        }
        if (SUMS) {
            for (int j = tid; j < M; j += THREADS) {
                unsigned long long val = hist_next.get(j, a.mc.fix_one);
                if (val) atomicAdd(outn + j, val);
            }
        }
    };

    // full tiles: every lane of every warp has work, so warp-wide primitives may use the full mask
    constexpr long long TILE = (long long)THREADS * UNROLL;
    const long long n_tiles = dead ? 0 : nvec / TILE;
    // The tile that will be loaded PIC_PREFETCH_TILES iterations from now is requested into the L2 by one thread
    // (two bulk prefetches: its x and its v vectors), so the 16-byte loads below find their lines on chip instead of
    // waiting out an HBM round trip with nothing but the other warps of the CTA to cover it.  The first tiles are
    // requested before the loop, i.e. they stream in under whatever remains of the prologue.
    // Measured (N = 1e9): 4.16 -> 3.84 ms for the 24-byte pass and 6.34 -> 6.07 ms for the final pass, but 5.23 -> 6.6 ms
    // for the 32-byte stage-1 pass, which already runs at 94 % of the HBM copy rate and only loses from read bursts
    // running ahead of its writes -- so the passes that are bound by HBM alone do without.
    constexpr int PF = (MODE == MODE_KICK || MODE == MODE_FINAL) ? PIC_PREFETCH_TILES : PIC_PREFETCH_TILES_HBM;
#ifndef PIC_TEX_LD_NOL1
#define PIC_TEX_LD_NOL1 1
#endif
    constexpr bool LD_NOL1 = PIC_TEX_LD_NOL1 != 0;      // texture-gather kernels: the particle stream bypasses the L1
    auto prefetch_tile = [&](long long tile) {
        if (PF > 0 && tid == 0 && tile < n_tiles) {
            prefetch_l2_bulk(xv + tile * TILE, (unsigned)(TILE * sizeof(V)));
            prefetch_l2_bulk(vv + tile * TILE, (unsigned)(TILE * sizeof(V)));
        }
    };
#pragma unroll
    for (int d = 0; d < PF; ++d) prefetch_tile((long long)bid + (long long)d * nb);
    for (long long tile = bid; tile < n_tiles; tile += nb) {
        prefetch_tile(tile + (long long)PF * nb);
        const long long base = tile * TILE + tid;
        V xs[UNROLL], vs[UNROLL];
        if constexpr (TEXG && KICK) {
            // Texture gather: the fetches of all the tile's particles are issued before the first is consumed (a texture
            // fetch takes several times as long as a shared load, and the atomics of the deposit are ordering points the
            // compiler will not move a later particle's fetch across).
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) { xs[u] = LD_NOL1 ? ld_stream_nol1(xv + base + u * THREADS) : ld_stream(xv + base + u * THREADS);
                                               vs[u] = LD_NOL1 ? ld_stream_nol1(vv + base + u * THREADS) : ld_stream(vv + base + u * THREADS); }
            R fs[UNROLL * VEC]; bool slowg[UNROLL * VEC]; Fetched<R, IP> qs[UNROLL * VEC];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                R* px = reinterpret_cast<R*>(&xs[u]);
                R* pv = reinterpret_cast<R*>(&vs[u]);
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    const int p = u * VEC + e;
                    if (REDRIFT) px[e] = drift<R>(px[e], pv[e], cpre, pc);
                    int il;
                    slowg[p] = fast_cell<R>(px[e], pc, M, il, fs[p]);
                    il = (int)min((unsigned)il, (unsigned)(M - 1));      // keep the fetch legal on the slow path
                    qs[p] = fetch_field<R, IP>(il, table, M);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                R* px = reinterpret_cast<R*>(&xs[u]);
                R* pv = reinterpret_cast<R*>(&vs[u]);
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    const int p = u * VEC + e;
                    if (p == 0) particle_substage_pre<R, IP, true, EXACT_W, true, H, true>(px[e], pv[e], fs[p], slowg[p], qs[p], hist, table, cc, dd, pc, a.mc, SUMS, err, &agg);
                    else particle_substage_pre<R, IP, true, EXACT_W, true, H, false>(px[e], pv[e], fs[p], slowg[p], qs[p], hist, table, cc, dd, pc, a.mc, SUMS, err, &agg);
                    if (SUMS) {
                        s2 += (double)pv[e] * (double)pv[e]; s1 += (double)pv[e];
                        next_stage0<R, IP, EXACT_W, true>(px[e], pv[e], hist_next, cnext, pc, a.mc, err, &agg);
                    }
                }
                if (STORE_X) st_stream(xv + base + u * THREADS, xs[u]);
                st_stream(vv + base + u * THREADS, vs[u]);
            }
        } else {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) { xs[u] = ld_stream(xv + base + u * THREADS); vs[u] = ld_stream(vv + base + u * THREADS); }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                R* px = reinterpret_cast<R*>(&xs[u]);
                R* pv = reinterpret_cast<R*>(&vs[u]);
                if constexpr (PAIRS) {
#pragma unroll
                    for (int e = 0; e < VEC; e += 2) {
                        if (u == 0 && e == 0) pair(px + e, pv + e, std::true_type{});
                        else pair(px + e, pv + e, std::false_type{});
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        if (u == 0 && e == 0) one(px[e], pv[e], std::true_type{}, std::true_type{});
                        else one(px[e], pv[e], std::true_type{}, std::false_type{});
                    }
                }
                if (STORE_X) st_stream(xv + base + u * THREADS, xs[u]);
                if (KICK) st_stream(vv + base + u * THREADS, vs[u]);
            }
        }
    }
    // ragged remainder (< one tile of vectors) and the scalar tail (N not a multiple of the vector width)
    if (!dead && bid == (unsigned)(n_tiles % nb)) {
        for (long long i = n_tiles * TILE + tid; i < nvec; i += THREADS) vec_pair(i, std::false_type{});
    }
    if (!dead && bid == 0) {
        long long i = nvec * VEC + tid;
        if (i < a.N) {
            R x = xe[i], v = ve[i];
            one(x, v, std::false_type{}, std::false_type{});
            if (STORE_X) xe[i] = x;
            if (KICK) ve[i] = v;
        }
    }
    __syncthreads();

    flush();
    if (SUMS) {
        double t2 = block_sum<THREADS>(s2, sm.red);
        double t1 = block_sum<THREADS>(s1, sm.red);
        if (tid == 0) {
            double* p = a.partial + ((size_t)env * nb + bid) * 2;
            p[0] = t2; p[1] = t1;
        }
    }
    if (err) atomicOr(a.err, err);
    if (fused) {                                        // single env per handle in this mode
        __shared__ double s_extra[2];                   // the last CTA also publishes the rank's kinetic sums
        __shared__ unsigned s_is_last;
        __threadfence();
        __syncthreads();
        if (tid == 0) s_is_last = atomicAdd(a.comm.ticket, 1u) == nb - 1 ? 1u : 0u;
        __syncthreads();
        if (s_is_last) {
            __threadfence();
            int n_extra = 0;
            if (SUMS) {
                double q2 = 0.0, q1 = 0.0;
                for (int i = tid; i < (int)nb; i += THREADS) { q2 += __ldcg(a.partial + 2 * i); q1 += __ldcg(a.partial + 2 * i + 1); }
                q2 = block_sum<THREADS>(q2, sm.red);
                q1 = block_sum<THREADS>(q1, sm.red);
                if (tid == 0) { s_extra[0] = q2; s_extra[1] = q1; }
                __syncthreads();
                n_extra = 2;
            }

            const size_t slot = ((size_t)(a.comm.seq_out % COMM_SETS) * a.comm.world + a.comm.rank) * a.comm.slot_len;
            constexpr int BATCH = 8;
            if (a.comm.exch_mc) {
                // NVLS: ONE store per word to the multicast mapping; the switch writes it into slot[set][rank] of every rank
                unsigned long long* dst = a.comm.exch_mc + slot;
                for (int j0 = tid; j0 < a.comm.out_words; j0 += THREADS * BATCH) {
                    unsigned long long w[BATCH];
#pragma unroll
                    for (int b = 0; b < BATCH; ++b) {
                        const int j = j0 + b * THREADS;
                        w[b] = j < a.comm.out_words ? __ldcg(a.comm.out_src + j) : 0ull;
                    }
#pragma unroll
                    for (int b = 0; b < BATCH; ++b) {
                        const int j = j0 + b * THREADS;
                        if (j < a.comm.out_words) multimem_st(dst + j, w[b]);
                    }
                }
                if (tid < n_extra) multimem_st(dst + a.comm.out_words + tid, (unsigned long long)__double_as_longlong(s_extra[tid]));
            } else {
            // each word is read once (all loads of a batch in flight together) and then posted to every peer
            for (int j0 = tid; j0 < a.comm.out_words; j0 += THREADS * BATCH) {
                unsigned long long w[BATCH];
#pragma unroll
                for (int b = 0; b < BATCH; ++b) {
                    const int j = j0 + b * THREADS;
                    w[b] = j < a.comm.out_words ? __ldcg(a.comm.out_src + j) : 0ull;
                }
                for (int r = 0; r < a.comm.world; ++r) {
                    unsigned long long* dst = a.comm.exch[r] + slot;
#pragma unroll
                    for (int b = 0; b < BATCH; ++b) {
                        const int j = j0 + b * THREADS;
                        if (j < a.comm.out_words) dst[j] = w[b];
                    }
                }
            }
            if (tid < n_extra)
                for (int r = 0; r < a.comm.world; ++r)
                    a.comm.exch[r][slot + a.comm.out_words + tid] = (unsigned long long)__double_as_longlong(s_extra[tid]);
            }
            __threadfence_system();
            __syncthreads();
            if (tid < a.comm.world) st_release_sys(a.comm.flags[tid] + a.comm.rank, a.comm.seq_out);
            if (tid == 0) *a.comm.ticket = 0u;
        }
    }
}

template <typename R, int THREADS, int UNROLL, int MODE, int DEP, bool EXACT_W, int IP = IP_CIC, bool TEXG = false>
__global__ void __launch_bounds__(THREADS) push_stream_kernel(const StreamArgs a) {
    push_stream_body<R, THREADS, UNROLL, MODE, DEP, EXACT_W, IP, TEXG, false>(a, a.act, 0u, 0u);
}

// ----------------------------------------------------------------- field table (texture gather)
// One CTA per env: the field solve that the streaming kernels otherwise repeat in every CTA's prologue, done once, and
// the gather table (E_j + ext_j, E_{j+1} + ext_{j+1}) written to global memory for the TEXG kernels of the next pass.
// The same block_field instance on the same integers: the table holds the same bits as the shared-memory one.
struct FieldTableArgs {
    MeshConst mc;
    const unsigned long long* rho_in;  // [n_envs][M]
    ActuatorArgs act;
    void* table;                       // [n_envs][M] pairs of R
    CommArgs comm;                     // fused exchange: the density is the sum of the ranks' slots of exchange seq_in
    unsigned* err;
};

template <typename R, int THREADS>
__global__ void __launch_bounds__(THREADS) field_table_kernel(const FieldTableArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using P = typename PairT<R>::type;
    const int env = blockIdx.x, M = a.mc.M, tid = threadIdx.x;
#ifndef PIC_NO_PDL
    griddep_launch_dependents();
    griddep_wait();
#endif
    SmemLayout<R> sm(smem_raw, M, false);
    const ExtSrc ext = stage_ext(a.act, env, M);
    if (a.comm.world > 1) {
        // the consumer side of the fused exchange for a texture-route pass: a peer that never arrives leaves the sticky
        // ERR_COMM_TIMEOUT flag, which makes the pass that follows leave the particle state alone
        comm_wait(a.comm, a.comm.seq_in, a.err);
        PeerSumRho rho{comm_in_slots(a.comm, a.comm.seq_in), a.comm.slot_len, a.comm.world};
        block_field<R, THREADS, false, true>(rho, sm.E_s, sm.D_s, sm.red, a.mc, ext, nullptr, nullptr, 0.0, 0.0, [] {});
    } else {
        GlobalRho rho{a.rho_in + (size_t)env * M};
        block_field<R, THREADS, false, true>(rho, sm.E_s, sm.D_s, sm.red, a.mc, ext, nullptr, nullptr, 0.0, 0.0, [] {});
    }
    P* out = (P*)a.table + (size_t)env * M;
    for (int j = tid; j < M; j += THREADS) out[j] = sm.E_s[j];
}

// ----------------------------------------------------------------- finalize
struct FinalizeArgs {
    MeshConst mc;
    RewardConst rw;
    const double* coeffs;              // [n_envs][2m] action applied in this step, or nullptr
    int two_m;
    const double* tw_cos; const double* tw_sin; int n_modes;   // spectral read-out tables ([M][n_modes]) or nullptr
    double* modes;                     // [n_envs][2 n_modes] or nullptr
    int step_done;                     // 1: this finalize closes an env step (emit reward), 0: initial state
    const unsigned long long* rho;     // [n_envs][M] state density (fixed point)
    unsigned long long* rho_zero;      // [n_envs][M] or nullptr
    double* n_out;                     // [n_envs][M]
    double* E_out;                     // [n_envs][M]
    double* diag;                      // [n_envs][DIAG_N]
    const double* partial;             // [n_envs][n_partial][2] or nullptr
    double* vsum;                      // [n_envs][2] local sum v^2, sum v (all-reduced by the host when sharded)
    int n_partial;
    CommArgs comm;                     // fused exchange: consume the state density + kinetic sums of all ranks
    unsigned long long* rho_reduced;   // fused exchange: where the summed state density is stored (the S buffer)
    double* trace_row;                 // nullptr or [n_envs][DIAG_N]: copy of the diagnostics record of this step
    unsigned* err;
};

// The finalize itself is a device function (one CTA per env calls it): field_finalize_kernel below, and the finalize CTA
// of the cooperative step kernel, which passes the action and the trace row of the step it closes.
template <int THREADS>
__device__ __forceinline__ void field_finalize_body(const FinalizeArgs& a, int env, const double* coeffs, double* trace_row) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int M = a.mc.M, tid = threadIdx.x;
    SmemLayout<double> sm(smem_raw, M, false);
    const ExtSrc none{nullptr, nullptr, nullptr, nullptr, 0};
    const ModeOut mo{a.tw_cos, a.tw_sin, a.modes ? a.modes + (size_t)env * 2 * a.n_modes : nullptr, a.n_modes};
    const bool fused = a.comm.world > 1;
    double s2 = 0.0, s1 = 0.0;
    FieldTotals t;
    if (fused) {                                       // single env: density and kinetic sums come from the slots
        comm_wait(a.comm, a.comm.seq_in, a.err);
        const unsigned long long* slots = comm_in_slots(a.comm, a.comm.seq_in);
        PeerSumRho rho{slots, a.comm.slot_len, a.comm.world};
        if (tid == 0) {                                // rank order: the same doubles in the same order on every rank
            const int kin = 2 * M;                     // words 2M, 2M+1 of a slot: sum v^2, sum v of that rank
            for (int r = 0; r < a.comm.world; ++r) {
                s2 += __longlong_as_double((long long)__ldcg(slots + (size_t)r * a.comm.slot_len + kin));
                s1 += __longlong_as_double((long long)__ldcg(slots + (size_t)r * a.comm.slot_len + kin + 1));
            }
        }
        for (int j = tid; j < M; j += THREADS) a.rho_reduced[j] = rho(j);
        t = block_field<double, THREADS, true, true>(rho, sm.E_s, sm.D_s, sm.red, a.mc, none, a.n_out, a.E_out, s2, s1, [] {}, mo, a.err);
    } else {
        GlobalRho rho{a.rho + (size_t)env * M};
        if (a.partial) {
            const double* p = a.partial + (size_t)env * a.n_partial * 2;
            for (int i = tid; i < a.n_partial; i += THREADS) { s2 += p[2 * i]; s1 += p[2 * i + 1]; }
        }
        t = block_field<double, THREADS, true, true>(rho, sm.E_s, sm.D_s, sm.red, a.mc, none,
                                         a.n_out + (size_t)env * M, a.E_out + (size_t)env * M, s2, s1, [] {}, mo, a.err);
    }
    if (a.rho_zero) for (int j = tid; j < M; j += THREADS) a.rho_zero[(size_t)env * M + j] = 0ull;
    if (tid == 0) {
        double* d = a.diag + (size_t)env * DIAG_N;
        if (a.step_done) {                             // reward of this transition: field energy of the PRE-step state
            const double ie = coeffs ? input_energy(coeffs + (size_t)env * a.two_m, a.two_m, a.rw.L) : 0.0;
            d[DIAG_REWARD] = reward_of(a.rw, d[DIAG_PE_MESH], ie);
            d[DIAG_INPUT_E] = ie;
        } else {
            d[DIAG_REWARD] = 0.0; d[DIAG_INPUT_E] = 0.0;
        }
        d[DIAG_PE_MESH] = 0.5 * t.e2 * a.mc.dx;
        d[DIAG_SUM_E2] = t.e2;
        if (a.partial || fused) {
            a.vsum[env * 2] = t.s1; a.vsum[env * 2 + 1] = t.s2;
            d[DIAG_KE] = 0.5 * t.s1; d[DIAG_SUM_V] = t.s2;
        }
        if (trace_row)
            for (int k = 0; k < DIAG_N; ++k) trace_row[(size_t)env * DIAG_N + k] = d[k];
    }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) field_finalize_kernel(const FinalizeArgs a) {
#ifndef PIC_NO_PDL
    griddep_launch_dependents();
    griddep_wait();
#endif
    field_finalize_body<THREADS>(a, blockIdx.x, a.coeffs, a.trace_row);
}

// ----------------------------------------------------------------- cooperative step (mid-size envs)
// One launch per call instead of four per env step.  For envs of 1e4 .. a few 1e7 particles the streaming step is not
// bound by its particles but by what sits between them: three kernel boundaries and a finalize launch per step (about
// 28 us for an env whose particles take a few us).  Here the three passes of a step run back to back inside ONE
// cooperative launch, separated by grid barriers, for any number of steps; the CTAs of column blockIdx.x < n_workers are
// the passes' CTAs (a CTA walks the same tiles in every pass, so particles never cross CTAs: only the densities and the
// kinetic partial sums do, through the L2), and one more CTA per env does the finalize of step s while the workers are
// in the first pass of step s + 1, which needs nothing the finalize produces -- the same overlap step_device arranges
// with a side stream, and the same buffer-clearing pattern (StreamArgs of the `overlap` flavour).  Same device functions,
// same order of every floating-point sum => bit-identical to the kernel-per-pass path with n_workers CTAs.
struct CoopArgs {
    StreamArgs s[3];                   // Yoshida stages 1, 2, 3; act.coeffs points at the action of the first step
    FinalizeArgs f;                    // coeffs / trace_row are taken from the fields below instead
    int n_steps;
    long long coeff_step_stride;       // elements between the actions of consecutive steps (n_envs * 2m)
    double* trace;                     // [n_steps][n_envs][DIAG_N]
    unsigned long long* barrier;       // monotonic arrival counter
    unsigned long long barrier_base;   // its value when this launch starts (the host keeps count)
    unsigned n_workers;                // gridDim.x - 1
};

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// every thread of every CTA of the (co-resident: cooperative launch) grid calls this
__device__ __forceinline__ void grid_barrier(unsigned long long* ctr, unsigned long long target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(ctr) : "memory");
        while (ld_acquire_gpu(ctr) < target) { }
        __threadfence();
    }
    __syncthreads();
}

template <typename R, int THREADS, int UNROLL, int DEP, bool EXACT_W>
__global__ void __launch_bounds__(THREADS) step_coop_kernel(const CoopArgs c) {
    const unsigned nw = c.n_workers, bid = blockIdx.x;
    const bool worker = bid < nw;
    const unsigned long long n_ctas = (unsigned long long)gridDim.x * gridDim.y;
    unsigned long long target = c.barrier_base;
    for (int s = 0; ; ++s) {
        ActuatorArgs act = c.s[0].act;
        if (act.coeffs) act.coeffs += (size_t)s * c.coeff_step_stride;
        if (worker) {
            if (s < c.n_steps)
                push_stream_body<R, THREADS, UNROLL, MODE_KICK0, DEP, EXACT_W, IP_CIC, false, true>(c.s[0], act, bid, nw);
        } else if (s > 0) {             // the finalize of the previous step, beside the workers' first pass of this one
            field_finalize_body<THREADS>(c.f, blockIdx.y, c.f.coeffs ? c.f.coeffs + (size_t)(s - 1) * c.coeff_step_stride : nullptr,
                                         c.trace + (size_t)(s - 1) * gridDim.y * DIAG_N);
        }
        if (s == c.n_steps) break;
        grid_barrier(c.barrier, target += n_ctas);
        if (worker) push_stream_body<R, THREADS, UNROLL, MODE_KICK, DEP, EXACT_W, IP_CIC, false, true>(c.s[1], act, bid, nw);
        grid_barrier(c.barrier, target += n_ctas);
        if (worker) push_stream_body<R, THREADS, UNROLL, MODE_FINAL, DEP, EXACT_W, IP_CIC, false, true>(c.s[2], act, bid, nw);
        grid_barrier(c.barrier, target += n_ctas);
    }
}

// ----------------------------------------------------------------- resident
struct ResidentArgs {
    MeshConst mc;
    PartConsts pcs;                    // per-particle constants of both precisions (host-computed)
    RewardConst rw;
    const double* tw_cos; const double* tw_sin; int n_modes;   // spectral read-out tables or nullptr
    double* modes;                     // [n_envs][2 n_modes] after the last step, or nullptr
    double* mode_trace;                // nullptr or [n_steps][n_envs][2 n_modes]
    void* x; void* v;                  // [n_envs][ld]
    long long N, ld;
    int n_steps;                       // 0 => init only (wrap + deposit + field)
    ActuatorArgs act;                  // ext: [n_envs][M]; coeffs: [n_steps][n_envs][2m]
    long long coeff_step_stride;       // elements between consecutive steps of coeffs (0 => same every step)
    long long ext_step_stride;         // same for ext
    double c[4], d[4];
    double* n_out;                     // [n_envs][M] state density after the last step
    double* E_out;                     // [n_envs][M] self-consistent field after the last step
    double* diag;                      // [n_envs][DIAG_N] after the last step
    double* trace;                     // nullptr or [n_steps][n_envs][DIAG_N]
    unsigned long long* rho_out;       // nullptr or [n_envs][M] fixed-point state density (for parity tests)
    unsigned* err;
    SmemOffsets lay;                   // shared-memory plan, computed on the host (smem_offsets): the kernel re-derives
                                       // addresses inside its particle loops, and a constant-bank load is the cheapest way
};

// One CTA per env; the particle state of the env lives in SHARED memory for the whole launch (any number of env
// steps): a run-time loop over the particles (any N that fits), two conflict-free 8-byte loads and stores per
// particle and sub-stage on top of the random gather / atomics.  (A register-resident variant was measured 1.4x
// slower: 64 registers per thread at 1024 threads spill, and at 512 threads only one CTA fits per SM, so nothing
// overlaps the barriers of the field solve.  With the state in shared memory two 512-thread CTAs share an SM.)
template <typename R>
__host__ __device__ constexpr size_t resident_smem_bytes(int M, int threads, long long n, int ip = IP_CIC) {
    return smem_plan_bytes<R>(M, threads, true, ip) + (size_t)((n + 1) / 2 * 2) * 2 * sizeof(R) + (size_t)M * 8;
}

template <typename R, int THREADS, int DEP, bool EXACT_W, int IP = IP_CIC>
__global__ void __launch_bounds__(THREADS, THREADS <= 512 ? 1024 / THREADS : 1) env_step_resident_kernel(const ResidentArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, env = blockIdx.x, M = a.mc.M, N = (int)a.N;
    SmemLayout<R> sm(smem_raw, a.lay);
    R* x_s = (R*)(smem_raw + a.lay.x);
    R* v_s = (R*)(smem_raw + a.lay.v);
    double* ext_s = (double*)(smem_raw + a.lay.ext);
    using H = typename HistSel<DEP, IP, R>::type;
    H hist; hist.init(sm.hist, M);
    #ifdef PIC_DEVICE_PARTCONST                          // experiment: recompute the constants in registers (round-1 behaviour)
    const PartConst<R> pc = make_part_const<R>(a.mc);
#else
    const PartConst<R>& pc = part_const<R>(a.pcs);
#endif
    R* xe = (R*)a.x + (size_t)env * a.ld;
    R* ve = (R*)a.v + (size_t)env * a.ld;
    for (int i = tid; i < N; i += THREADS) { x_s[i] = xe[i]; v_s[i] = ve[i]; }
    hist.zero(tid, THREADS);
    __syncthreads();
    unsigned err = 0;
    SharedRho<H> rho{&hist, a.mc.fix_one};
    double* n_out = a.n_out + (size_t)env * M;
    double* E_out = a.E_out + (size_t)env * M;
    const ExtSrc none{nullptr, nullptr, nullptr, nullptr, 0};
    constexpr int FT = FieldShape<THREADS>::FT;
    auto clear_hist = [&]() {
        if (FT == THREADS) hist.zero(tid, THREADS);
        else hist.zero(tid - FT, THREADS - FT);
    };
    auto kinetic = [&](double& s2, double& s1) {
        s2 = 0.0; s1 = 0.0;
        for (int i = tid; i < N; i += THREADS) { const double vj = (double)v_s[i]; s2 += vj * vj; s1 += vj; }
    };
    double pe_pre = 0.0;
    if (tid == 0 && a.n_steps > 0) pe_pre = a.diag[(size_t)env * DIAG_N + DIAG_PE_MESH];
    auto write_diag = [&](const FieldTotals& t, int step, double input_e) {
        if (tid == 0) {
            double rec[DIAG_N];
            rec[DIAG_KE] = 0.5 * t.s1; rec[DIAG_PE_MESH] = 0.5 * t.e2 * a.mc.dx; rec[DIAG_SUM_V] = t.s2; rec[DIAG_SUM_E2] = t.e2;
            rec[DIAG_REWARD] = 0.0; rec[DIAG_INPUT_E] = 0.0;
            if (step >= 0) {
                rec[DIAG_INPUT_E] = input_e;
                rec[DIAG_REWARD] = reward_of(a.rw, pe_pre, rec[DIAG_INPUT_E]);
                pe_pre = rec[DIAG_PE_MESH];
            }
            double* d = a.diag + (size_t)env * DIAG_N;
#pragma unroll
            for (int k = 0; k < DIAG_N; ++k) d[k] = rec[k];
            if (a.trace && step >= 0) {
                double* tr = a.trace + ((size_t)step * gridDim.x + env) * DIAG_N;
#pragma unroll
                for (int k = 0; k < DIAG_N; ++k) tr[k] = rec[k];
            }
        }
    };
    auto dump_rho = [&]() {
        if (a.rho_out) for (int j = tid; j < M; j += THREADS) a.rho_out[(size_t)env * M + j] = rho(j);
    };

    if (a.n_steps == 0) {
        for (int i = tid; i < N; i += THREADS) {
            R x = x_s[i], v = v_s[i];
            particle_substage<R, IP, false, false, EXACT_W, false>(x, v, hist, sm.E_s, (R)0, (R)0, pc, a.mc, true, err);
            x_s[i] = x;
        }
        __syncthreads();
        dump_rho();
        double s2, s1;
        kinetic(s2, s1);
        const ModeOut mo{a.tw_cos, a.tw_sin, a.modes ? a.modes + (size_t)env * 2 * a.n_modes : nullptr, a.n_modes};
        write_diag(block_field<R, THREADS>(rho, sm.E_s, sm.D_s, sm.red, a.mc, none, n_out, E_out, s2, s1, clear_hist, mo, a.err),
                   -1, 0.0);
    }

    for (int step = 0; step < a.n_steps; ++step) {
        ActuatorArgs act = a.act;
        if (act.coeffs) act.coeffs += (size_t)step * a.coeff_step_stride;
        if (act.ext) act.ext += (size_t)step * a.ext_step_stride;
        const ExtSrc ext_g = stage_ext(act, env, M);
        const bool last = step == a.n_steps - 1;
        double input_e = 0.0;                                       // loaded now, used after the last field solve
        if (tid == 0 && ext_g.coeff) input_e = input_energy(ext_g.coeff, 2 * a.act.m, a.rw.L);
        // E_external is held for all sub-stages of a step (pic.py:131-137): evaluate it on the mesh once, here, where its
        // global loads overlap the stage-0 particle loop, instead of inside each of the three kick-stage field solves.
        // The first reader is behind the block barriers of the stage-0 solve.
        ExtSrc ext = none;
        if (ext_g.any()) {
            for (int j = tid; j < M; j += THREADS) ext_s[j] = ext_g.at(j);
            ext.ext = ext_s;
        }
        {                                                           // stage 0: drift only (integration.py:71)
            const R cc = (R)a.c[0];
#pragma unroll 2
            for (int i = tid; i < N; i += THREADS) {
                R x = x_s[i], v = v_s[i];
                particle_substage<R, IP, false, true, EXACT_W, false>(x, v, hist, sm.E_s, cc, (R)0, pc, a.mc, false, err);
                x_s[i] = x;
            }
            __syncthreads();
            block_field<R, THREADS, false>(rho, sm.E_s, sm.D_s, sm.red, a.mc, ext, nullptr, nullptr, 0.0, 0.0, clear_hist);
        }
#pragma unroll 1
        for (int st = 1; st < 3; ++st) {                            // stages 1, 2: kick + drift
            const R cc = (R)a.c[st], dd = (R)a.d[st];
#pragma unroll 2
            for (int i = tid; i < N; i += THREADS) {
                R x = x_s[i], v = v_s[i];
                particle_substage<R, IP, true, true, EXACT_W, false>(x, v, hist, sm.E_s, cc, dd, pc, a.mc, false, err);
                x_s[i] = x; v_s[i] = v;
            }
            __syncthreads();
            block_field<R, THREADS, false>(rho, sm.E_s, sm.D_s, sm.red, a.mc, ext, nullptr, nullptr, 0.0, 0.0, clear_hist);
        }
        {                                                           // stage 3: kick + drift + state wrap + kinetic sums
            const R cc = (R)a.c[3], dd = (R)a.d[3];
            double s2 = 0.0, s1 = 0.0;
#pragma unroll 2
            for (int i = tid; i < N; i += THREADS) {
                R x = x_s[i], v = v_s[i];
                particle_substage<R, IP, true, true, EXACT_W, false>(x, v, hist, sm.E_s, cc, dd, pc, a.mc, true, err);
                x_s[i] = x; v_s[i] = v;
                s2 += (double)v * (double)v; s1 += (double)v;
            }
            __syncthreads();
            if (last) dump_rho();
            double* mout = nullptr;
            if (a.n_modes > 0) {
                if (a.mode_trace) mout = a.mode_trace + ((size_t)step * gridDim.x + env) * 2 * a.n_modes;
                else if (last && a.modes) mout = a.modes + (size_t)env * 2 * a.n_modes;
            }
            const ModeOut mo{a.tw_cos, a.tw_sin, mout, a.n_modes};
            write_diag(block_field<R, THREADS>(rho, sm.E_s, sm.D_s, sm.red, a.mc, none, last ? n_out : nullptr,
                                               last ? E_out : nullptr, s2, s1, clear_hist, mo, a.err), step, input_e);
        }
    }
    for (int i = tid; i < N; i += THREADS) { xe[i] = x_s[i]; ve[i] = v_s[i]; }
    if (err) atomicOr(a.err, err);
}

// ----------------------------------------------------------------- resident, one env over a CTA cluster
// The same env step with the env spread over CL CTAs of a thread-block cluster (CL SMs, or CL slots of the same SMs):
// CTA r keeps the r-th slice of the particles in its shared memory and deposits into its own histogram; the field
// solve of every sub-stage reads ALL the cluster's histograms through distributed shared memory (each CTA solves the
// whole mesh redundantly -- the mesh is tiny -- so nothing has to be sent back), and the per-CTA kinetic sums are
// exchanged the same way.  Why:
//   * finer scheduling units.  With one CTA per env two 512-thread CTAs share an SM, and a batch that does not fill a
//     whole number of waves (512 envs on 2 x 148 slots = 1.73 waves) pays for 2; half-env CTAs of 256 threads run four
//     to an SM, the last wave is 3/4 full instead of 1/2, and an SM with three co-resident CTAs loses far less
//     throughput than one with a single CTA of two;
//   * envs that do not fit ONE CTA's shared memory (N > ~13 000 particles at float64; the SAC shape N = 10 000,
//     N_mesh = 500 fits only one CTA per SM) stay in the resident kernel -- one launch per call, the state never leaves
//     the chip -- instead of falling back to three streaming launches per env step.
// Histograms are double-buffered: sub-stage s deposits into buffer s & 1 and the other one -- last read by the peers
// during the solve of sub-stage s - 1, which they have left before arriving at this sub-stage's cluster barrier -- is
// cleared meanwhile.  That makes ONE cluster barrier per sub-stage enough (4 per env step).
// Integer densities: x, v, fields are bit-identical to the one-CTA kernel; only the kinetic sums (float) differ in
// their last bits because they are added up in a different order.
template <typename H, int CL> struct ClusterRho {
    H h[CL];                       // every rank's histogram of the current buffer (generic pointers into peer smem)
    long long one;
    __device__ __forceinline__ unsigned long long operator()(int j) const {
        unsigned long long s = 0;
#pragma unroll
        for (int r = 0; r < CL; ++r) s += h[r].get(j, one);
        return s;
    }
};

__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {       // barrier.cluster with release / acquire semantics
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// generic address of `p` (a shared-memory object of this CTA) in the shared memory of cluster rank r
template <typename T> __device__ __forceinline__ T* map_to_rank(T* p, unsigned r) {
    unsigned long long out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"((unsigned long long)p), "r"(r));
    return (T*)out;
}

template <typename R>
__host__ __device__ constexpr size_t cluster_smem_bytes(int M, int threads, long long n_slice, int cl, int ip = IP_CIC) {
    // gather table + two histograms + D_s + scratch + particle slice + E_ext of the step + per-warp kinetic partials
    return smem_plan_bytes<R>(M, threads, true, ip, true) + (size_t)((n_slice + 1) / 2 * 2) * 2 * sizeof(R) + (size_t)M * 8 +
           (size_t)(threads / 32) * 2 * 8;
}

template <typename R, int THREADS, int CL, bool EXACT_W, int IP = IP_CIC>
__global__ void __launch_bounds__(THREADS, THREADS <= 512 ? 1024 / THREADS : 1) env_step_cluster_kernel(const ResidentArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, M = a.mc.M;
    const unsigned crank = cluster_ctarank();
    const int env = blockIdx.x / CL;
    const int n_lo = (int)(a.N * crank / CL), n_hi = (int)(a.N * (crank + 1) / CL), N = n_hi - n_lo;   // this CTA's slice
    SmemLayout<R> sm(smem_raw, a.lay);
    R* x_s = (R*)(smem_raw + a.lay.x);
    R* v_s = (R*)(smem_raw + a.lay.v);
    double* ext_s = (double*)(smem_raw + a.lay.ext);
    double* part_s = ext_s + M;                               // [NW][2] per-warp sum v^2, sum v of the final sub-stage
    using H = typename HistSel<DEP_SPLIT32, IP, R>::type;
    // Two histogram buffers; sub-stage s of an env step uses buffer s & 1 (four sub-stages per step: every step starts
    // on buffer 0), so the buffer is a compile-time constant everywhere below and everything stays in registers.
    H hist0, hist1;
    hist0.init(sm.hist, M); hist1.init(sm.hist2, M);
    ClusterRho<H, CL> rho0, rho1;
    rho0.one = rho1.one = a.mc.fix_one;
#pragma unroll
    for (int r = 0; r < CL; ++r) {
        rho0.h[r] = hist0; rho0.h[r].w = map_to_rank(hist0.w, (unsigned)r);
        rho1.h[r] = hist1; rho1.h[r].w = map_to_rank(hist1.w, (unsigned)r);
    }
    #ifdef PIC_DEVICE_PARTCONST                          // experiment: recompute the constants in registers (round-1 behaviour)
    const PartConst<R> pc = make_part_const<R>(a.mc);
#else
    const PartConst<R>& pc = part_const<R>(a.pcs);
#endif
    R* xe = (R*)a.x + (size_t)env * a.ld + n_lo;
    R* ve = (R*)a.v + (size_t)env * a.ld + n_lo;
    for (int i = tid; i < N; i += THREADS) { x_s[i] = xe[i]; v_s[i] = ve[i]; }
    hist0.zero(tid, THREADS); hist1.zero(tid, THREADS);
    __syncthreads();
    unsigned err = 0;
    const bool lead = crank == 0;                             // rank 0 writes the env's outputs
    double* n_out = a.n_out + (size_t)env * M;
    double* E_out = a.E_out + (size_t)env * M;
    const ExtSrc none{nullptr, nullptr, nullptr, nullptr, 0};
    constexpr int FT = FieldShape<THREADS>::FT;
    auto clear = [&](H& h) {                                  // called by the threads block_field leaves idle
        if (FT == THREADS) h.zero(tid, THREADS);
        else h.zero(tid - FT, THREADS - FT);
    };
    // per-warp kinetic partials -> shared memory (before the cluster barrier); summed over ranks and warps in a fixed
    // order by warp 0 of every CTA after it, so that all CTAs of the cluster hold the same doubles
    auto publish_kinetic = [&](double s2, double s1) {
        s2 = warp_sum(s2); s1 = warp_sum(s1);
        if (lane == 0) { part_s[2 * warp] = s2; part_s[2 * warp + 1] = s1; }
    };
    auto gather_kinetic = [&](double& s2, double& s1) {       // valid in warp 0
        s2 = 0.0; s1 = 0.0;
        if (warp == 0) {
            for (int k = lane; k < CL * NW; k += 32) {
                const double* p = map_to_rank(part_s, (unsigned)(k / NW)) + 2 * (k % NW);
                s2 += p[0]; s1 += p[1];
            }
            s2 = warp_sum(s2); s1 = warp_sum(s1);
        }
    };
    double pe_pre = 0.0;
    if (tid == 0 && a.n_steps > 0) pe_pre = a.diag[(size_t)env * DIAG_N + DIAG_PE_MESH];
    auto write_diag = [&](const FieldTotals& t, double s2, double s1, int step, double input_e) {
        if (tid == 0 && lead) {
            double rec[DIAG_N];
            rec[DIAG_KE] = 0.5 * s2; rec[DIAG_PE_MESH] = 0.5 * t.e2 * a.mc.dx; rec[DIAG_SUM_V] = s1; rec[DIAG_SUM_E2] = t.e2;
            rec[DIAG_REWARD] = 0.0; rec[DIAG_INPUT_E] = 0.0;
            if (step >= 0) {
                rec[DIAG_INPUT_E] = input_e;
                rec[DIAG_REWARD] = reward_of(a.rw, pe_pre, rec[DIAG_INPUT_E]);
                pe_pre = rec[DIAG_PE_MESH];
            }
            double* d = a.diag + (size_t)env * DIAG_N;
#pragma unroll
            for (int k = 0; k < DIAG_N; ++k) d[k] = rec[k];
            if (a.trace && step >= 0) {
                double* tr = a.trace + ((size_t)step * (gridDim.x / CL) + env) * DIAG_N;
#pragma unroll
                for (int k = 0; k < DIAG_N; ++k) tr[k] = rec[k];
            }
        }
    };
    auto dump_rho = [&](const ClusterRho<H, CL>& rho) {
        if (a.rho_out && lead) for (int j = tid; j < M; j += THREADS) a.rho_out[(size_t)env * M + j] = rho(j);
    };
    // one kick/drift sub-stage into histogram `hc` (summed over the cluster by `rc`), clearing the other buffer `ho`
    auto kick_stage = [&](H& hc, const ClusterRho<H, CL>& rc, H& ho, R cc, R dd, const ExtSrc& ext) {
#pragma unroll 2
        for (int i = tid; i < N; i += THREADS) {
            R x = x_s[i], v = v_s[i];
            particle_substage<R, IP, true, true, EXACT_W, false>(x, v, hc, sm.E_s, cc, dd, pc, a.mc, false, err);
            x_s[i] = x; v_s[i] = v;
        }
        cluster_sync_all();
        block_field<R, THREADS, false>(rc, sm.E_s, sm.D_s, sm.red, a.mc, ext, nullptr, nullptr, 0.0, 0.0, [&] { clear(ho); });
    };

    if (a.n_steps == 0) {                                            // initial state: wrap + deposit + field (buffer 0)
        double s2 = 0.0, s1 = 0.0;
        for (int i = tid; i < N; i += THREADS) {
            R x = x_s[i], v = v_s[i];
            particle_substage<R, IP, false, false, EXACT_W, false>(x, v, hist0, sm.E_s, (R)0, (R)0, pc, a.mc, true, err);
            x_s[i] = x;
            s2 += (double)v * (double)v; s1 += (double)v;
        }
        publish_kinetic(s2, s1);
        cluster_sync_all();
        dump_rho(rho0);
        gather_kinetic(s2, s1);
        const ModeOut mo{a.tw_cos, a.tw_sin, (a.modes && lead) ? a.modes + (size_t)env * 2 * a.n_modes : nullptr, a.n_modes};
        const FieldTotals t = block_field<R, THREADS>(rho0, sm.E_s, sm.D_s, sm.red, a.mc, none, lead ? n_out : nullptr,
                                                      lead ? E_out : nullptr, 0.0, 0.0, [&] { clear(hist1); }, mo, a.err);
        write_diag(t, s2, s1, -1, 0.0);
    }

    for (int step = 0; step < a.n_steps; ++step) {
        ActuatorArgs act = a.act;
        if (act.coeffs) act.coeffs += (size_t)step * a.coeff_step_stride;
        if (act.ext) act.ext += (size_t)step * a.ext_step_stride;
        const ExtSrc ext_g = stage_ext(act, env, M);
        const bool last = step == a.n_steps - 1;
        double input_e = 0.0;
        if (tid == 0 && ext_g.coeff) input_e = input_energy(ext_g.coeff, 2 * a.act.m, a.rw.L);
        ExtSrc ext = none;
        if (ext_g.any()) {                                           // E_external of the step on the mesh, once (pic.py:131-137)
            for (int j = tid; j < M; j += THREADS) ext_s[j] = ext_g.at(j);
            ext.ext = ext_s;
        }
        {                                                            // stage 0: drift only (integration.py:71); buffer 0
            const R cc = (R)a.c[0];
#pragma unroll 2
            for (int i = tid; i < N; i += THREADS) {
                R x = x_s[i], v = v_s[i];
                particle_substage<R, IP, false, true, EXACT_W, false>(x, v, hist0, sm.E_s, cc, (R)0, pc, a.mc, false, err);
                x_s[i] = x;
            }
            cluster_sync_all();
            block_field<R, THREADS, false>(rho0, sm.E_s, sm.D_s, sm.red, a.mc, ext, nullptr, nullptr, 0.0, 0.0, [&] { clear(hist1); });
        }
        kick_stage(hist1, rho1, hist0, (R)a.c[1], (R)a.d[1], ext);   // stage 1: kick + drift; buffer 1
        kick_stage(hist0, rho0, hist1, (R)a.c[2], (R)a.d[2], ext);   // stage 2; buffer 0
        {                                                            // stage 3: kick + drift + state wrap + kinetic sums; buffer 1
            const R cc = (R)a.c[3], dd = (R)a.d[3];
            double s2 = 0.0, s1 = 0.0;
#pragma unroll 2
            for (int i = tid; i < N; i += THREADS) {
                R x = x_s[i], v = v_s[i];
                particle_substage<R, IP, true, true, EXACT_W, false>(x, v, hist1, sm.E_s, cc, dd, pc, a.mc, true, err);
                x_s[i] = x; v_s[i] = v;
                s2 += (double)v * (double)v; s1 += (double)v;
            }
            publish_kinetic(s2, s1);
            cluster_sync_all();
            if (last) dump_rho(rho1);
            gather_kinetic(s2, s1);
            double* mout = nullptr;
            if (a.n_modes > 0 && lead) {
                if (a.mode_trace) mout = a.mode_trace + ((size_t)step * (gridDim.x / CL) + env) * 2 * a.n_modes;
                else if (last && a.modes) mout = a.modes + (size_t)env * 2 * a.n_modes;
            }
            const ModeOut mo{a.tw_cos, a.tw_sin, mout, a.n_modes};
            const FieldTotals t = block_field<R, THREADS>(rho1, sm.E_s, sm.D_s, sm.red, a.mc, none,
                                                          (last && lead) ? n_out : nullptr, (last && lead) ? E_out : nullptr,
                                                          0.0, 0.0, [&] { clear(hist0); }, mo, a.err);
            write_diag(t, s2, s1, step, input_e);
        }
    }
    for (int i = tid; i < N; i += THREADS) { xe[i] = x_s[i]; ve[i] = v_s[i]; }
    if (err) atomicOr(a.err, err);
    cluster_sync_all();                      // no CTA leaves while a peer may still be reading its shared memory
}

// ------------------------------------------------------------- small helpers
template <typename T, typename U>
__global__ void convert_kernel(const T* __restrict__ in, U* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (U)in[i];
}

}  // namespace pic
