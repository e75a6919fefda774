// Device-side building blocks of the 1D electrostatic PIC step (sm_100a).
//
// Everything here restates *what* the reference computes
//   deposit     src/env/interpolate.py:4-20  (CIC)   / src/env/util.py:48-62
//   field       src/env/util.py:99-103, src/env/solve.py:27-53, src/env/util.py:7-46
//   gather      src/env/util.py:106
//   push        src/env/integration.py:22-47,60-75, src/env/pic.py:125-146
// but not *how*: no dense matrices, no Thomas solve, no temporaries in memory.
//
// Numerical contract (DESIGN.md "Numerics"):
//   * wrap and cell index are bit-exact with np.mod(np.mod(x,L),L) and
//     floor(x/dx) (IEEE division) -- rn intrinsics keep ptxas from contracting
//     index-critical expressions into FMAs.
//   * push arithmetic replays the reference's operation order with rn
//     intrinsics, so given the same mesh field x and v are bit-identical.
//   * the deposit accumulates llrint(w * 2^k) in integers (associative =>
//     independent of thread / CTA / GPU count); the field is one prefix sum.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pic {

constexpr int DIAG_KE = 0;        // 0.5 * sum v^2            (src/env/util.py:144)
constexpr int DIAG_PE_MESH = 1;   // 0.5 * sum E_mesh^2 * dx  (src/control/objective.py:31)
constexpr int DIAG_SUM_V = 2;     // sum v (momentum)
constexpr int DIAG_SUM_E2 = 3;    // sum E_mesh^2
constexpr int DIAG_REWARD = 4;    // reward of the transition that produced this state (src/control/rl/reward.py:71-76)
constexpr int DIAG_INPUT_E = 5;   // sum(a^2) * L / 4 of the action applied (reward.py:52-54)
constexpr int DIAG_N = 6;
constexpr int MAX_MODES = 8;      // Fourier modes of E_mesh emitted per step (spectrum.py:17)

// Reward.compute_reward (reward.py:71-76): alpha * max(1 - PE_mesh(s_t)/r_pe_n, 0) + beta * max(1 - IE(a_t)/r_ie_n, 0)
struct RewardConst { double alpha, beta, r_pe_n, r_ie_n, L; };

__device__ __forceinline__ double input_energy(const double* __restrict__ coeff, int two_m, double L) {
    double s = 0.0;
    for (int k = 0; k < two_m; ++k) s += coeff[k] * coeff[k];
    return s * L * 0.25;
}
__device__ __forceinline__ double reward_of(const RewardConst& rc, double pe_pre, double ie) {
    const double r_pe = fmax(1.0 - pe_pre / rc.r_pe_n, 0.0), r_ie = fmax(1.0 - ie / rc.r_ie_n, 0.0);
    return r_pe * rc.alpha + r_ie * rc.beta;
}

// Optional spectral read-out of the self-consistent field: E_k = fft(E_mesh)[k] / N_mesh * 2 for k = 1..m
// (src/interpret/spectrum.py:17; what run_feedback.py and the behaviour-cloning target of ddpg.py:429-431 use).
struct ModeOut {
    const double* tw_cos;   // [M][m] cos(2 pi j k / M)
    const double* tw_sin;   // [M][m]
    double* out;            // [2m]: Re_1..Re_m, Im_1..Im_m for this env, or nullptr
    int m;
};

constexpr unsigned ERR_INDEX_RANGE = 1u;   // floor(x/dx) fell outside [0, N_mesh) (the reference would raise in np.bincount)
constexpr unsigned ERR_NONFINITE = 2u;     // non-finite position reached the deposit
constexpr unsigned ERR_DENSITY_RANGE = 8u; // a cell's fixed-point sum left [0, 2^63): more than ~8x the mean density

// ---------------------------------------------------------------- real traits
template <typename R> struct RT;

template <> struct RT<double> {
    using vec = double2;                       // 16-byte vector = 2 particles
    static constexpr int VEC = 2;
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double flo(double a) { return floor(a); }
    static __device__ __forceinline__ double rnd(double a) { return rint(a); }
    static __device__ __forceinline__ double mod(double a, double b) { return fmod(a, b); }
    static __device__ __forceinline__ double abs(double a) { return fabs(a); }
    static __device__ __forceinline__ int toint(double a) { return __double2int_rd(a); }
    static __device__ __forceinline__ double fromint(int a) { return __int2double_rn(a); }
    // round-to-nearest-even integer of q (|q| < 2^31) on the fp64 add pipe: adding 2^52+2^51 leaves rint(q) in the
    // low mantissa word.  No F2I / FRND (those run at a fraction of the DADD rate).
    // `ok` is false unless 0 <= rint(q) < 2^32 (the high word of the sum is then exactly the magic's): rejects
    // negative, huge and non-finite q with one integer compare.
    static __device__ __forceinline__ double rint_magic(double q, int& i, bool& ok) {
        const double MAGIC = 6755399441055744.0;
        double t = __dadd_rn(q, MAGIC);
        i = __double2loint(t);
        ok = __double2hiint(t) == 0x43380000;
        return __dsub_rn(t, MAGIC);
    }
    // floor(q) the same way with a round-down add (a rounding-mode flag on the same DADD): no nearest-integer fix-up
    static __device__ __forceinline__ double floor_magic(double q, int& i, bool& ok) {
        const double MAGIC = 6755399441055744.0;
        double t = __dadd_rd(q, MAGIC);
        i = __double2loint(t);
        ok = __double2hiint(t) == 0x43380000;
        return __dsub_rn(t, MAGIC);
    }
    // floor(x * a) for 0 <= x * a < 2^32 by ONE fused multiply-add rounded down against the magic constant (the exact
    // product plus 2^52+2^51 rounds down to an integer + magic); the integer sits in the low mantissa word.
    // `hi_ok`: the high word is the magic's, i.e. 0 <= floor < 2^32 and the input was finite.
    static __device__ __forceinline__ int floor_mul_magic(double x, double a, bool& hi_ok, double& t) {
        t = __fma_rd(x, a, 6755399441055744.0);
        hi_ok = __double2hiint(t) == 0x43380000;
        return __double2loint(t);
    }
    static __device__ __forceinline__ double unmagic(double t) { return __dsub_rn(t, 6755399441055744.0); }
};

template <> struct RT<float> {
    using vec = float4;                        // 16-byte vector = 4 particles
    static constexpr int VEC = 4;
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float flo(float a) { return floorf(a); }
    static __device__ __forceinline__ float rnd(float a) { return rintf(a); }
    static __device__ __forceinline__ float mod(float a, float b) { return fmodf(a, b); }
    static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
    static __device__ __forceinline__ int toint(float a) { return __float2int_rd(a); }
    static __device__ __forceinline__ float fromint(int a) { return __int2float_rn(a); }
    static __device__ __forceinline__ float rint_magic(float q, int& i, bool& ok) {       // ok: 0 <= rint(q) < 2^22
        const float MAGIC = 12582912.0f;
        float t = __fadd_rn(q, MAGIC);
        i = __float_as_int(t) - __float_as_int(MAGIC);
        ok = (__float_as_int(t) >> 22) == (0x4B400000 >> 22);
        return __fsub_rn(t, MAGIC);
    }
    static __device__ __forceinline__ float floor_magic(float q, int& i, bool& ok) {      // ok: 0 <= floor(q) < 2^22
        const float MAGIC = 12582912.0f;
        float t = __fadd_rd(q, MAGIC);
        i = __float_as_int(t) - __float_as_int(MAGIC);
        ok = (__float_as_int(t) >> 22) == (0x4B400000 >> 22);
        return __fsub_rn(t, MAGIC);
    }
    static __device__ __forceinline__ int floor_mul_magic(float x, float a, bool& hi_ok, float& t) {   // 0 <= floor < 2^22
        t = __fmaf_rd(x, a, 12582912.0f);
        hi_ok = (__float_as_int(t) >> 22) == (0x4B400000 >> 22);
        return __float_as_int(t) - 0x4B400000;
    }
    static __device__ __forceinline__ float unmagic(float t) { return __fsub_rn(t, 12582912.0f); }
};

// llrint(w * 2^k) for |w * 2^k| < 2^51 without a conversion instruction: one fused multiply-add against the magic
// constant rounds (w*2^k is exact) to an integer held in the mantissa.
__device__ __forceinline__ long long fix_weight(double w, double fix_scale) {
    const double MAGIC = 6755399441055744.0;
    double t = __fma_rn(w, fix_scale, MAGIC);
    return __double_as_longlong(t) - __double_as_longlong(MAGIC);
}

// gather table entry: (E_j, E_{j+1 mod M}) side by side so that the CIC gather is ONE 16-byte (8-byte in f32)
// shared load instead of two loads with independent bank conflicts
template <typename R> struct PairT;
template <> struct PairT<double> { using type = double2; };
template <> struct PairT<float> { using type = float2; };

// The same table read through the TEXTURE pipe (streaming kernels with gather = texture): a linear texture object over a
// global copy of the pair table, written by field_table_kernel before the pass.  The fetch runs on the SM's texture
// pipe and is served by the L1, so it no longer competes with the deposit's shared atomics for the LSU data pipe --
// the unit that bounds the SM-side passes (a random 16-byte LDS costs 10.6 LSU wavefronts per warp next to 3.7 per
// atomic; tools/tex_probe.cu: 6 atomics + LDS gather 3.41 ms per 1e9 particles, + texture gather 2.30 ms, no gather 2.29).
template <typename R> struct TexTable;
template <> struct TexTable<double> {
    cudaTextureObject_t tex; int base;                 // base: first texel of this env's table
    __device__ __forceinline__ double2 operator[](int i) const {
        const int4 q = tex1Dfetch<int4>(tex, base + i);
        return make_double2(__hiloint2double(q.y, q.x), __hiloint2double(q.w, q.z));
    }
};
template <> struct TexTable<float> {
    cudaTextureObject_t tex; int base;
    __device__ __forceinline__ float2 operator[](int i) const { return tex1Dfetch<float2>(tex, base + i); }
};

// ------------------------------------------------------------ kernel parameters
struct MeshConst {
    int M;                 // N_mesh
    double L, dx, inv_dx;  // dx = L / N_mesh (src/env/pic.py:36)
    double dx2, inv2dx;    // dx * dx, 1 / (2 dx): field-solve constants
    double n0;
    double scale;          // n0 * L / N / dx, left to right (src/env/interpolate.py:18)
    double fix_scale;      // 2^k
    double inv_fix;        // 2^-k
    long long fix_one;     // 2^k as an integer: what one particle deposits in total
    double idx_thr;        // |q - rint(q)| below this => redo the cell index with IEEE division
    double dt;
    int field_g, field_nvw; // field solve: cells per chunk (virtual lane) = ceil(N_mesh / 1024), virtual warps in use
    long long range_floor; // a cell sum below this has wrapped past 2^63 (ERR_DENSITY_RANGE).  CIC sums are never
                           // negative: -4 * 2^k; TSC's middle weight goes down to -0.25 per particle, so a clumped cell
                           // is legitimately negative: -2^61, far below anything the 8x headroom rule admits
};

template <typename R> struct PartConst {   // per-particle constants in the particle precision
    R L, twoL, dx, inv_dx, dt, idx_thr;
    R inv_lo, inv_hi;      // (1/dx) (1 -+ eps), eps = idx_thr / N_mesh: the bracket of fast_cell
};

template <typename R>
__host__ __device__ inline PartConst<R> make_part_const(const MeshConst& m) {
    PartConst<R> c;
    c.L = (R)m.L; c.twoL = (R)(2.0 * m.L); c.dx = (R)m.dx; c.inv_dx = (R)1 / c.dx; c.dt = (R)m.dt;
    c.idx_thr = (R)m.idx_thr;
    const R eps = (R)(m.idx_thr / (double)m.M);
    c.inv_lo = c.inv_dx * ((R)1 - eps);
    c.inv_hi = c.inv_dx * ((R)1 + eps);
    return c;
}

// Both precisions side by side, computed once on the HOST and passed inside the kernel arguments: the hot loops then
// read these constants straight from the constant bank (fp64 instructions take a constant-bank operand) instead of
// holding ~14 registers of loop invariants that every kernel would otherwise recompute at its start -- the streaming
// and resident kernels live at the 64-register limit.
struct PartConsts { PartConst<double> d; PartConst<float> f; };
template <typename R> __device__ __forceinline__ const PartConst<R>& part_const(const PartConsts& p);
template <> __device__ __forceinline__ const PartConst<double>& part_const<double>(const PartConsts& p) { return p.d; }
template <> __device__ __forceinline__ const PartConst<float>& part_const<float>(const PartConsts& p) { return p.f; }

// ---------------------------------------------------------------- wrap / cell
// np.mod(np.mod(x, L), L): util.py:51 followed by interpolate.py:6 (and pic.py:139 for the state itself).
// np.mod = fmod, then +L when the remainder is non-zero and negative.  For |x| < 2L every branch below is the
// exact value of that sequence: [L,2L) -> x-L is exact (Sterbenz); (-L,0) -> fl(x+L), which can round to L and
// is then sent to 0 by the second mod.
template <typename R>
__device__ __noinline__ R wrap_pos_far(R x, R L) {                      // cold: |x| >= 2L or non-finite (-> -1)
    if (!(RT<R>::abs(x) <= (R)3.0e38)) return (R)-1;
    R r = RT<R>::mod(x, L);
    if (r != (R)0) { if (r < (R)0) r = RT<R>::add(r, L); } else r = (R)0;
    R r2 = RT<R>::mod(r, L);                         // second mod: only r == L changes
    if (r2 != (R)0) { if (r2 < (R)0) r2 = RT<R>::add(r2, L); } else r2 = (R)0;
    return r2;
}

template <typename R>
__device__ __forceinline__ R wrap_pos(R x, const PartConst<R>& c, unsigned& err) {
    if (x >= (R)0 && x < c.L) return x;
    if (x >= c.L && x < c.twoL) return RT<R>::sub(x, c.L);
    if (x < (R)0 && x > -c.L) {
        R r = RT<R>::add(x, c.L);
        return (r == c.L) ? (R)0 : r;
    }
    R r = wrap_pos_far<R>(x, c.L);
    if (r < (R)0) { err |= ERR_NONFINITE; r = (R)0; }
    return r;
}

// cold: exact quotient for positions within rounding distance of a cell edge
template <typename R>
__device__ __noinline__ R floor_div_exact(R xw, R dx) { return RT<R>::flo(RT<R>::div(xw, dx)); }

struct Cell {
    int il, ir;          // left / right cell; ir is NOT reduced mod M (tables are padded with one wrap cell)
};

// floor(xw / dx) with correctly rounded division (interpolate.py:8).  Fast path: q = xw * (1/dx), r = rint(q) by the
// magic-number add, floor = r or r-1 by the sign of q - r.  Whenever q is within idx_thr of an integer (where the
// reciprocal multiply could land on the other side of it) the quotient is redone with IEEE division, so the index
// is always the reference's.  xw must be finite and in [0, L] (wrap_pos guarantees it).  Returns floor as R in f.
template <typename R>
__device__ __forceinline__ int cell_index(R xw, const PartConst<R>& c, int M, R& f, unsigned& err) {
    R q = RT<R>::mul(xw, c.inv_dx);
    int il; bool ok;
    R r = RT<R>::rint_magic(q, il, ok);
    R diff = RT<R>::sub(q, r);
    if (diff < (R)0) { r = RT<R>::sub(r, (R)1); il -= 1; }
    if (!ok || RT<R>::abs(diff) <= c.idx_thr) {          // rare: exact quotient
        r = floor_div_exact<R>(xw, c.dx);
        il = RT<R>::toint(r);
    }
    if (__builtin_expect((unsigned)il >= (unsigned)M, 0)) {   // the reference raises in np.bincount here
        err |= ERR_INDEX_RANGE;
        il = il < 0 ? 0 : M - 1;
        r = RT<R>::fromint(il);
    }
    f = r;
    return il;
}

// the two CIC weights (interpolate.py:11-12); EXACT_W selects true divisions instead of multiplying by 1/dx
template <typename R, bool EXACT_W>
__device__ __forceinline__ Cell cell_weights(R xw, const PartConst<R>& c, int M, R& wl, R& wr, unsigned& err) {
    R f;
    int il = cell_index<R>(xw, c, M, f, err);
    R fr = RT<R>::add(f, (R)1);
    R nl = RT<R>::sub(RT<R>::mul(fr, c.dx), xw);     // indx_r * dx - x
    R nr = RT<R>::sub(xw, RT<R>::mul(f, c.dx));      // x - indx_l * dx
    if (EXACT_W) { wl = RT<R>::div(nl, c.dx); wr = RT<R>::div(nr, c.dx); }
    else         { wl = RT<R>::mul(nl, c.inv_dx); wr = RT<R>::mul(nr, c.inv_dx); }
    Cell k; k.il = il; k.ir = il + 1;
    return k;
}

// ------------------------------------------------------------------ deposits
// 64-bit integer accumulation in shared memory.  sm_100a has no native 64-bit (or floating point) shared
// atomic add -- ptxas emits an ATOMS.CAST.SPIN loop -- so two flavours are provided:
//   DEP_CAS64   : atomicAdd on unsigned long long (CAS loop): 2^k - W_r to cell i_l, W_r to cell i_l + 1
//                 (table of M+1 cells, the wrap cell M is folded into cell 0 when read).
//   DEP_SPLIT32 : native 32-bit ATOMS.ADD only.  Each particle touches ONLY its own cell i_l: a particle count
//                 cnt[i_l] += 1 and the 64-bit sum S[i_l] += W_r kept as two 32-bit words (the low-word add
//                 returns the old value, so exactly one thread sees each carry and forwards it to the high word).
//                 The cell density in fixed point is then  cnt[j] 2^k - S[j] + S[j-1]  -- three atomics per
//                 particle instead of four and no wrap cell.
// Both give the identical integer density (tests/test_gpu_parity.py::test_bitwise_reproducibility_across_kernels).
constexpr int DEP_CAS64 = 0;
constexpr int DEP_SPLIT32 = 1;

template <int DEP> struct Hist;

template <> struct Hist<DEP_CAS64> {
    unsigned long long* h;
    int M;
    static __host__ __device__ constexpr size_t bytes(int M) { return (size_t)(M + 1) * 8; }
    __device__ __forceinline__ void init(void* base, int M_) { h = (unsigned long long*)base; M = M_; }
    __device__ __forceinline__ void zero(int tid, int nthreads) {
        for (int j = tid; j <= M; j += nthreads) h[j] = 0ull;
    }
    __device__ __forceinline__ void deposit(int il, long long Wr, long long one) {
        atomicAdd(&h[il], (unsigned long long)(one - Wr));
        atomicAdd(&h[il + 1], (unsigned long long)Wr);
    }
    // `count` particles of cell il with right-weight sum S
    __device__ __forceinline__ void deposit_group(int il, unsigned long long S, unsigned count, long long one) {
        atomicAdd(&h[il], (unsigned long long)count * (unsigned long long)one - S);
        atomicAdd(&h[il + 1], S);
    }
    __device__ __forceinline__ unsigned long long get(int j, long long) const {
        return j == 0 ? h[0] + h[M] : h[j];
    }
};

// native 32-bit shared atomics by address in the shared window (inline PTX: keeps ptxas from wrapping constant
// increments in its own MATCH.ANY-based warp aggregation, and needs no generic-to-shared conversion per access)
#ifdef PIC_ATOM_NOCLOBBER      // experiment: let the compiler interleave the particles of a thread across the atomics
#define PIC_ATOM_CLOBBER
#else
#define PIC_ATOM_CLOBBER : "memory"
#endif
__device__ __forceinline__ void red_shared_u32(unsigned addr, unsigned v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) PIC_ATOM_CLOBBER);
}
__device__ __forceinline__ unsigned atom_shared_u32(unsigned addr, unsigned v) {
    unsigned old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) PIC_ATOM_CLOBBER);
    return old;
}
// the same at a constant byte offset from addr (folded into the instruction's immediate)
template <int OFF> __device__ __forceinline__ void red_shared_u32_at(unsigned addr, unsigned v) {
    asm volatile("red.shared.add.u32 [%0+%2], %1;" ::"r"(addr), "r"(v), "n"(OFF) PIC_ATOM_CLOBBER);
}
template <int OFF> __device__ __forceinline__ unsigned atom_shared_u32_at(unsigned addr, unsigned v) {
    unsigned old;
    asm volatile("atom.shared.add.u32 %0, [%1+%3], %2;" : "=r"(old) : "r"(addr), "r"(v), "n"(OFF) PIC_ATOM_CLOBBER);
    return old;
}

// the same, issued only by the lanes whose addend is non-zero (a predicated ATOMS, no branch): lanes that are
// predicated off cost no data-pipe wavefront
template <int OFF> __device__ __forceinline__ void red_shared_u32_at_nonzero(unsigned addr, unsigned v) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p red.shared.add.u32 [%0+%2], %1;\n\t}"
                 ::"r"(addr), "r"(v), "n"(OFF) PIC_ATOM_CLOBBER);
}

// Layout: three words per cell side by side (cnt, S.lo, S.hi), so one address computation serves all three atomics
// (immediate offsets 0 / 4 / 8).  The stride of 3 words is coprime with the 32 banks: random cells spread over the
// banks exactly as three separate arrays would, and a flush that walks the cells in order is conflict-free.
// RARE_HI (Hist<DEP_SPLIT32_RARE>, the float32 kernels): with at most 32 fractional bits a particle's weight fits the
// low word, the high word only ever receives carries, and those are rare (one deposit in 2^(32-k+1) on average) -- the
// third atomic is predicated on its addend and all but disappears from the LSU data pipe.  Same integers either way.
template <bool RARE_HI> struct HistSplit32 {
    unsigned* w;
    unsigned w_a;                        // the same array as a shared-window byte address
    int M;
    static __host__ __device__ constexpr size_t bytes(int M) { return (size_t)M * 12 + 8; }
    __device__ __forceinline__ void init(void* base, int M_) {
        M = M_; w = (unsigned*)base;
        w_a = (unsigned)__cvta_generic_to_shared(w);
    }
    __device__ __forceinline__ void zero(int tid, int nthreads) {
        for (int j = tid; j < 3 * M; j += nthreads) w[j] = 0u;
    }
    __device__ __forceinline__ void deposit_group(int il, unsigned long long S, unsigned count, long long) {
        const unsigned wl = (unsigned)S, wh = (unsigned)(S >> 32), cell = w_a + 12u * (unsigned)il;
        const unsigned old = atom_shared_u32_at<4>(cell, wl);
        red_shared_u32_at<0>(cell, count);
        unsigned hi_add;                               // wh + carry-out of (old + wl): one add with carry-out, one with carry-in
        asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %3, 0;\n\t}" : "=r"(hi_add) : "r"(old), "r"(wl), "r"(wh));
        if (RARE_HI) red_shared_u32_at_nonzero<8>(cell, hi_add);
        else red_shared_u32_at<8>(cell, hi_add);
    }
    __device__ __forceinline__ void deposit(int il, long long Wr, long long one) {
        deposit_group(il, (unsigned long long)Wr, 1u, one);
    }
    __device__ __forceinline__ unsigned long long S(int j) const {
        return ((unsigned long long)w[3 * j + 2] << 32) | (unsigned long long)w[3 * j + 1];
    }
    __device__ __forceinline__ unsigned long long get(int j, long long one) const {
        return (unsigned long long)w[3 * j] * (unsigned long long)one - S(j) + S(j == 0 ? M - 1 : j - 1);
    }
};
constexpr int DEP_SPLIT32_RARE = 2;       // internal: selected by HistSel for float32, never requested through the C ABI
template <> struct Hist<DEP_SPLIT32> : HistSplit32<false> {};
template <> struct Hist<DEP_SPLIT32_RARE> : HistSplit32<true> {};

// TSC (3-point) deposit, src/env/interpolate.py:22-44: a particle in cell m gives W_l to cell m-1, W_r to cell m+1 and
// 2^k - W_l - W_r to its own cell (the reference's three weights sum to one; the middle one may be negative, which
// the modular integer arithmetic handles).  Same single-cell trick as the CIC split deposit: the particle only
// touches cell m -- cnt[m] += 1, A[m] += W_l, B[m] += W_r (two 64-bit sums in 32-bit words) -- and
//   density[j] = cnt[j] 2^k - A[j] - B[j] + A[j+1] + B[j-1].
constexpr int IP_CIC = 0;
constexpr int IP_TSC = 1;

struct HistTSC {
    unsigned* w;                          // cnt | alo | ahi | blo | bhi, M words each
    unsigned base_a;
    int M;
    static __host__ __device__ constexpr size_t bytes(int M) { return (size_t)M * 20 + 16; }
    __device__ __forceinline__ void init(void* base, int M_) { M = M_; w = (unsigned*)base; base_a = (unsigned)__cvta_generic_to_shared(w); }
    __device__ __forceinline__ void zero(int tid, int nthreads) { for (int j = tid; j < 5 * M; j += nthreads) w[j] = 0u; }
    __device__ __forceinline__ void add64(unsigned lo_addr, unsigned hi_addr, unsigned long long S) {
        const unsigned wl = (unsigned)S, wh = (unsigned)(S >> 32);
        const unsigned old = atom_shared_u32(lo_addr, wl);
        red_shared_u32(hi_addr, wh + ((old + wl) < old ? 1u : 0u));
    }
    __device__ __forceinline__ void deposit_group(int il, unsigned long long SA, unsigned long long SB, unsigned count) {
        const unsigned off = base_a + 4u * (unsigned)il, st = 4u * (unsigned)M;
        red_shared_u32(off, count);
        add64(off + st, off + 2 * st, SA);
        add64(off + 3 * st, off + 4 * st, SB);
    }
    __device__ __forceinline__ unsigned long long A(int j) const { return ((unsigned long long)w[2 * M + j] << 32) | w[M + j]; }
    __device__ __forceinline__ unsigned long long B(int j) const { return ((unsigned long long)w[4 * M + j] << 32) | w[3 * M + j]; }
    __device__ __forceinline__ unsigned long long get(int j, long long one) const {
        const int jp = j == M - 1 ? 0 : j + 1, jm = j == 0 ? M - 1 : j - 1;
        return (unsigned long long)w[j] * (unsigned long long)one - A(j) - B(j) + A(jp) + B(jm);
    }
};

// ------------------------------------------------------------ block primitives
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* scratch /* THREADS/32 + 1 doubles */) {
    constexpr int NW = THREADS / 32;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();                       // scratch may still be read by a previous call
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    double t = 0.0;
    if (w == 0) {
        t = lane < NW ? scratch[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) scratch[NW] = t;
    }
    __syncthreads();
    return scratch[NW];
}

// E_ext on the mesh from Fourier coefficients: basis_cos @ a + basis_sin @ b on the actuator's own node table
// (src/control/actuator.py:62; tables are uploaded from the host so node positions are the reference's linspace).
__device__ __forceinline__ double actuator_field_at(int j, int m, const double* __restrict__ bcos,
                                                    const double* __restrict__ bsin,
                                                    const double* __restrict__ coeff /* [2m]: cos then sin */) {
    double ec = 0.0, es = 0.0;
    for (int k = 0; k < m; ++k) {
        ec = __dadd_rn(ec, __dmul_rn(bcos[j * m + k], coeff[k]));
        es = __dadd_rn(es, __dmul_rn(bsin[j * m + k], coeff[m + k]));
    }
    return __dadd_rn(ec, es);
}

// What the particles see on top of the self-consistent field (util.py:102-103): nothing, a mesh vector, or the
// actuator's Fourier series evaluated on the fly.
struct ExtSrc {
    const double* ext;     // [M] or nullptr
    const double* coeff;   // [2m] or nullptr (takes precedence)
    const double* bcos;    // [M][m]
    const double* bsin;    // [M][m]
    int m;
    __device__ __forceinline__ bool any() const { return ext != nullptr || coeff != nullptr; }
    __device__ __forceinline__ double at(int j) const {
        if (coeff) return actuator_field_at(j, m, bcos, bsin, coeff);
        return ext[j];
    }
};

// ------------------------------------------------------------------ field solve
// Periodic 3-point Poisson + centred difference in closed form (DESIGN.md "Field solve"):
//   b_j = n_j - n0,  S_j = sum_{i<=j} b_i,  D_j = dx^2 (S_j - mean S),  E_j = -(D_j + D_{j-1}) / (2 dx)
// which is the reference's Thomas/Sherman-Morrison solve of laplacian @ phi = b followed by -grad @ phi
// (src/env/util.py:99-100) without forming phi.
//
// Block-cooperative with TWO block barriers (A after the chunk prefix sums, C before the result is used).  Two
// instances, picked by the size of the MESH (see block_field below):
//
// Large meshes (block_field_body).  The floating-point sums are organised over VIRTUAL lanes, not over the threads that
// happen to run them, so that every launch shape produces the same bits: the mesh is cut into 1024 chunks of
// g = ceil(N_mesh / 1024) consecutive cells (four at 4096); 32 consecutive chunks form a virtual warp.  Within a chunk
// the cells are added left to right, within a virtual warp the chunk sums are scanned with the shuffle tree, across
// virtual warps the totals are added in order.  A field warp of the CTA takes the virtual warps w, w + NWF, ... one
// after the other: with 1024 threads every lane owns one chunk, with 256 four.
// Each virtual warp publishes two numbers -- its sum of b and its share of sum_j S_j = sum_i b_i (M - i), a weighted
// sum that does not depend on the scan -- from which every field thread derives its offsets, the grand total and
// sum_j S_j without further communication (S_{j-1} of a chunk's first cell is its exclusive offset, so D_{j-1} never
// comes from a neighbour).  Every thread must call this.
//
//   rho      : M fixed-point cell sums (functor; global, shared or distributed shared memory)
//   E_s      : shared, M pairs (E_j + ext_j, E_{j+1 mod M} + ext_{j+1 mod M}) for the gather
//   D_s      : shared scratch, M doubles;  red: shared scratch, field_scratch_doubles(THREADS) doubles
//   ext      : external field source added to what particles see (util.py:102-103)
//   n_out/E_out : nullptr or global outputs of the density (interpolate.py:18) / self-consistent field
//   x1, x2   : two per-thread values summed over the block on the way (kinetic sums)
//   modes    : optional Fourier read-out of the self-consistent field (written by warp 0 after barrier C)
// Returns {sum_j E_j^2 (self-consistent field), sum x1, sum x2}; valid in warp 0 only.  TOTALS = false (the field is
// only needed for the next kick): the three block sums and their shuffles are skipped, the result is zeros.
struct FieldTotals { double e2, s1, s2; };

constexpr int FIELD_VWARPS = 32;                                  // virtual warps (1024 virtual lanes) at most
template <int THREADS> struct FieldShape {
    static constexpr int FT = THREADS < 256 ? THREADS : 256;     // default number of field threads
    static constexpr int NWF = FT / 32, NW = THREADS / 32;
};
// scratch: [3 x FIELD_VWARPS] scan triples | [FIELD_VWARPS] sum E^2 per virtual warp | [2 x NW] kinetic partials |
//          [2 MAX_MODES x NW] mode partials
__host__ __device__ constexpr int field_scratch_doubles(int threads) {
    return 4 * FIELD_VWARPS + (2 + 2 * MAX_MODES) * (threads / 32);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// FT_: number of field threads (a multiple of 32, >= 32).  The resident kernels keep the default (<= 256); the
// streaming prologue and the finalize kernel have nothing else to do and use every thread: at N_mesh = 4096 a lane's
// chain of loads, int64 -> double conversions and dependent adds is then 4 cells long instead of 16, which was most of
// the fixed cost of a pass (12 of 20 us).
template <typename R, int THREADS, bool TOTALS, int FT_, int ROUNDS, typename RhoLoad, typename IdleWork>
__device__ __forceinline__ FieldTotals block_field_body(RhoLoad rho, typename PairT<R>::type* E_s, double* D_s, double* red,
                                                        const MeshConst& mc, const ExtSrc& ext, double* __restrict__ n_out,
                                                        double* __restrict__ E_out, double x1, double x2,
                                                        IdleWork idle_work, const ModeOut modes, unsigned* range_err) {
    constexpr int FT = FT_, NWF = FT_ / 32, NW = FieldShape<THREADS>::NW;   // ROUNDS: virtual warps per field warp, at most
    const int M = mc.M, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const bool field_thread = tid < FT;
    const int g = mc.field_g, nvw = mc.field_nvw;                     // cells per chunk (virtual lane), virtual warps in use
    double* red_e2 = red + 3 * FIELD_VWARPS;                          // [FIELD_VWARPS]
    double* red_kin = red_e2 + FIELD_VWARPS;                          // [NW][2]
    double* red_mode = red_kin + 2 * NW;                              // [NW][2 * MAX_MODES]
    const bool has_ext = ext.any();
    const bool want_modes = modes.out != nullptr;

    double excl[ROUNDS];
    if (field_thread) {
        bool wrapped = false;
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int vw = w + r * NWF;                               // warp-uniform
            excl[r] = 0.0;
            if (vw < nvw) {
                const int j0 = (vw * 32 + lane) * g, j1 = min(M, j0 + g);
                double run = 0.0, ws = 0.0;
                for (int j = j0; j < j1; ++j) {
                    const long long rj = (long long)rho(j);
                    wrapped |= rj < mc.range_floor;                   // wrapped past 2^63
                    // rn intrinsics / explicit fma throughout: this function is inlined into every kernel flavour, and
                    // ptxas must not be free to contract a multiply-add in one of them and not in another
                    const double nj = __dmul_rn(__dmul_rn((double)rj, mc.inv_fix), mc.scale);
                    if (n_out) n_out[j] = nj;
                    const double b = __dsub_rn(nj, mc.n0);
                    run = __dadd_rn(run, b);
                    D_s[j] = run;                                     // inclusive prefix inside the chunk
                    ws = __fma_rn(b, (double)(M - j), ws);            // b_j enters S_j, S_{j+1}, ..., S_{M-1}
                }
                double inc = run;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    double t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc = __dadd_rn(inc, t);
                }
                excl[r] = __dsub_rn(inc, run);
                const double a = warp_sum(ws);
                const double wt = __shfl_sync(0xffffffffu, inc, 31);
                if (lane == 0) { red[3 * vw] = wt; red[3 * vw + 1] = a; }
            }
        }
        if (range_err && wrapped) atomicOr(range_err, ERR_DENSITY_RANGE);
    }
    __syncthreads();                                   // (A) triples published; all reads of rho are done

    double kin1 = x1, kin2 = x2;
    if (field_thread) {
        double total = 0.0, sumS = 0.0, myoff[ROUNDS];
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) myoff[r] = 0.0;
        for (int k = 0; k < nvw; ++k) {
#pragma unroll
            for (int r = 0; r < ROUNDS; ++r) if (k == w + r * NWF) myoff[r] = total;
            sumS = __dadd_rn(sumS, red[3 * k + 1]);
            total = __dadd_rn(total, red[3 * k]);
        }
        const double meanS = sumS / (double)M;
        const double dx2 = mc.dx2, inv2dx = mc.inv2dx;
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int vw = w + r * NWF;
            if (vw < nvw) {
                const int j0 = (vw * 32 + lane) * g, j1 = min(M, j0 + g);
                const double off = __dadd_rn(myoff[r], excl[r]);
                double prevS = j0 == 0 ? total : off;  // S_{j0-1}; periodic: S_{-1} = S_{M-1} = total
                double e2 = 0.0;
                for (int j = j0; j < j1; ++j) {
                    const int jm = j == 0 ? M - 1 : j - 1;
                    const double S = __dadd_rn(D_s[j], off);
                    const double E = __dmul_rn(-__dadd_rn(__dmul_rn(dx2, __dsub_rn(S, meanS)), __dmul_rn(dx2, __dsub_rn(prevS, meanS))), inv2dx);
                    prevS = S;
                    if (E_out) E_out[j] = E;
                    if (TOTALS) e2 = __fma_rn(E, E, e2);
                    if (want_modes) D_s[j] = E;         // S_j is dead from here on; keep E_j for the read-out below
                    const double Et = has_ext ? __dadd_rn(E, ext.at(j)) : E;
                    E_s[j].x = (R)Et;
                    E_s[jm].y = (R)Et;
                }
                if (TOTALS) {                           // sum E^2 per VIRTUAL warp: the same tree in every flavour
                    e2 = warp_sum(e2);
                    if (lane == 0) red_e2[vw] = e2;
                }
            }
        }
        if (want_modes) {                               // cold: one mode at a time keeps the register footprint small
            for (int k = 0; k < modes.m; ++k) {
                double re = 0.0, im = 0.0;
                for (int r = 0; r < ROUNDS; ++r) {
                    const int vw = w + r * NWF;
                    if (vw < nvw) {
                        const int j0 = (vw * 32 + lane) * g, j1 = min(M, j0 + g);
                        for (int j = j0; j < j1; ++j) {
                            re = __fma_rn(D_s[j], modes.tw_cos[j * modes.m + k], re);
                            im = __fma_rn(-D_s[j], modes.tw_sin[j * modes.m + k], im);
                        }
                    }
                }
                re = warp_sum(re); im = warp_sum(im);
                if (lane == 0) { red_mode[w * 2 * MAX_MODES + k] = re; red_mode[w * 2 * MAX_MODES + MAX_MODES + k] = im; }
            }
        }
    } else {
        idle_work();
    }
    if (FT == THREADS) idle_work();                    // no spare threads: everyone does it after its share
    if (TOTALS) {
        kin1 = warp_sum(kin1); kin2 = warp_sum(kin2);
        if (lane == 0) { red_kin[2 * w] = kin1; red_kin[2 * w + 1] = kin2; }
    }
    __syncthreads();                                   // (C) gather table, idle work and partial sums complete

    FieldTotals t{0.0, 0.0, 0.0};
    if (TOTALS && w == 0) {
        t.e2 = warp_sum(lane < nvw ? red_e2[lane] : 0.0);
        t.s1 = warp_sum(lane < NW ? red_kin[2 * lane] : 0.0);
        t.s2 = warp_sum(lane < NW ? red_kin[2 * lane + 1] : 0.0);
        if (want_modes && lane < 2 * modes.m) {         // lane < m: Re_{lane+1}; m <= lane < 2m: Im_{lane-m+1}
            const int k = lane < modes.m ? lane : lane - modes.m;
            const int slot = lane < modes.m ? k : MAX_MODES + k;
            double acc = 0.0;
#pragma unroll
            for (int ww = 0; ww < NWF; ++ww) acc += red_mode[ww * 2 * MAX_MODES + slot];
            modes.out[lane] = acc / (double)M * 2.0;
        }
    }
    return t;
}

// Small meshes (N_mesh <= FIELD_SMALL_MESH): 256 field threads own contiguous runs of ceil(N_mesh / 256) <= 4 cells.
// Same structure as the large-mesh body with the chunk = one thread's run and ONE virtual warp per field warp, written
// out straight-line: this is what the resident kernels run four times per env step (a third of their time).
template <typename R, int THREADS, bool TOTALS, typename RhoLoad, typename IdleWork>
__device__ __forceinline__ FieldTotals block_field_small(RhoLoad rho, typename PairT<R>::type* E_s, double* D_s, double* red,
                                                         const MeshConst& mc, const ExtSrc& ext, double* __restrict__ n_out,
                                                         double* __restrict__ E_out, double x1, double x2,
                                                         IdleWork idle_work, const ModeOut modes, unsigned* range_err) {
    constexpr int FT = FieldShape<THREADS>::FT, NWF = FieldShape<THREADS>::NWF, NW = FieldShape<THREADS>::NW;
    const int M = mc.M, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const bool field_thread = tid < FT;
    const int cpt = (M + FT - 1) / FT;
    const int j0 = tid * cpt, j1 = min(M, j0 + cpt);
    double* red2 = red + 3 * NWF;
    double* red3 = red2 + 3 * NW;                       // [NWF][2 * MAX_MODES] mode partials
    const bool has_ext = ext.any();
    const bool want_modes = modes.out != nullptr;

    // rn intrinsics / explicit fma on everything that reaches E: this function is inlined into every kernel flavour, and
    // ptxas must not be free to contract a multiply-add in one of them and not in another (the flavours promise
    // identical bits)
    double run = 0.0, excl = 0.0, ext0 = 0.0;
    if (field_thread) {
        if (has_ext && j0 < j1) ext0 = ext.at(j0);     // global loads: in flight across the scan and barrier (A)
        double ws = 0.0;
        for (int j = j0; j < j1; ++j) {
            const long long rj = (long long)rho(j);
            if (range_err && rj < mc.range_floor) atomicOr(range_err, ERR_DENSITY_RANGE);   // wrapped past 2^63
            const double nj = __dmul_rn(__dmul_rn((double)rj, mc.inv_fix), mc.scale);
            if (n_out) n_out[j] = nj;
            const double b = __dsub_rn(nj, mc.n0);
            run = __dadd_rn(run, b);
            D_s[j] = run;                              // inclusive prefix inside this thread's run of cells
            ws = __fma_rn(b, (double)(M - j), ws);     // b_j enters S_j, S_{j+1}, ..., S_{M-1}
        }
        double inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc = __dadd_rn(inc, t);
        }
        excl = __dsub_rn(inc, run);
        const double a = warp_sum(ws);
        const double wt = __shfl_sync(0xffffffffu, inc, 31);
        if (lane == 0) { red[3 * w] = wt; red[3 * w + 1] = a; }
    }
    __syncthreads();                                   // (A) warp triples published; all reads of rho are done

    double e2 = 0.0;
    if (field_thread) {
        double total = 0.0, sumS = 0.0, myoff = 0.0;
#pragma unroll
        for (int k = 0; k < NWF; ++k) {
            if (k == w) myoff = total;
            sumS = __dadd_rn(sumS, red[3 * k + 1]);
            total = __dadd_rn(total, red[3 * k]);
        }
        const double meanS = sumS / (double)M;
        const double off = __dadd_rn(myoff, excl);
        const double dx2 = mc.dx2, inv2dx = mc.inv2dx;
        double prevS = j0 == 0 ? total : off;          // S_{j0-1}; periodic: S_{-1} = S_{M-1} = total
        for (int j = j0; j < j1; ++j) {
            const int jm = j == 0 ? M - 1 : j - 1;
            const double S = __dadd_rn(D_s[j], off);
            const double E = __dmul_rn(-__dadd_rn(__dmul_rn(dx2, __dsub_rn(S, meanS)), __dmul_rn(dx2, __dsub_rn(prevS, meanS))), inv2dx);
            prevS = S;
            if (E_out) E_out[j] = E;
            if (TOTALS) e2 = __fma_rn(E, E, e2);
            if (want_modes) D_s[j] = E;                 // S_j is dead from here on; keep E_j for the read-out below
            const double Et = has_ext ? __dadd_rn(E, j == j0 ? ext0 : ext.at(j)) : E;
            E_s[j].x = (R)Et;
            E_s[jm].y = (R)Et;
        }
        if (want_modes) {                               // cold: one mode at a time keeps the register footprint small
            for (int k = 0; k < modes.m; ++k) {
                double re = 0.0, im = 0.0;
                for (int j = j0; j < j1; ++j) {
                    re = __fma_rn(D_s[j], modes.tw_cos[j * modes.m + k], re);
                    im = __fma_rn(-D_s[j], modes.tw_sin[j * modes.m + k], im);
                }
                re = warp_sum(re); im = warp_sum(im);
                if (lane == 0) { red3[w * 2 * MAX_MODES + k] = re; red3[w * 2 * MAX_MODES + MAX_MODES + k] = im; }
            }
        }
    } else {
        idle_work();
    }
    if (FT == THREADS) idle_work();                    // no spare threads: everyone does it after its share
    if (TOTALS) {
        e2 = warp_sum(e2); x1 = warp_sum(x1); x2 = warp_sum(x2);
        if (lane == 0) { red2[3 * w] = e2; red2[3 * w + 1] = x1; red2[3 * w + 2] = x2; }
    }
    __syncthreads();                                   // (C) gather table, idle work and partial sums complete

    FieldTotals t{0.0, 0.0, 0.0};
    if (TOTALS && w == 0) {
        t.e2 = warp_sum(lane < NW ? red2[3 * lane] : 0.0);
        t.s1 = warp_sum(lane < NW ? red2[3 * lane + 1] : 0.0);
        t.s2 = warp_sum(lane < NW ? red2[3 * lane + 2] : 0.0);
        if (want_modes && lane < 2 * modes.m) {         // lane < m: Re_{lane+1}; m <= lane < 2m: Im_{lane-m+1}
            const int k = lane < modes.m ? lane : lane - modes.m;
            const int slot = lane < modes.m ? k : MAX_MODES + k;
            double acc = 0.0;
#pragma unroll
            for (int ww = 0; ww < NWF; ++ww) acc += red3[ww * 2 * MAX_MODES + slot];
            modes.out[lane] = acc / (double)M * 2.0;
        }
    }
    return t;
}

// The field solve.  LARGE_OK = false (resident and cluster kernels: N_mesh <= FIELD_SMALL_MESH is a precondition that
// pic_create enforces) compiles the small-mesh instance only; the streaming prologue and the finalize kernel take any
// mesh and pick by size (CTA-uniform branch).  Both instances order their sums by the MESH, never by the launch shape:
// 256 runs of <= 4 cells up to 1024 cells, 1024 chunks of ceil(N_mesh / 1024) cells above -- so every kernel flavour,
// thread count and GPU count builds the identical field from the identical integer density.
constexpr int FIELD_SMALL_MESH = 1024;
template <typename R, int THREADS, bool TOTALS = true, bool LARGE_OK = false, typename RhoLoad, typename IdleWork>
__device__ __forceinline__ FieldTotals block_field(RhoLoad rho, typename PairT<R>::type* E_s, double* D_s, double* red,
                                                   const MeshConst& mc, const ExtSrc& ext, double* __restrict__ n_out,
                                                   double* __restrict__ E_out, double x1, double x2,
                                                   IdleWork idle_work, const ModeOut modes = ModeOut{nullptr, nullptr, nullptr, 0},
                                                   unsigned* range_err = nullptr) {
    if constexpr (LARGE_OK) {
        constexpr int ROUNDS = (FIELD_VWARPS + THREADS / 32 - 1) / (THREADS / 32);
        if (mc.M > FIELD_SMALL_MESH)
            return block_field_body<R, THREADS, TOTALS, THREADS, ROUNDS>(rho, E_s, D_s, red, mc, ext, n_out, E_out, x1, x2,
                                                                        idle_work, modes, range_err);
    }
    return block_field_small<R, THREADS, TOTALS>(rho, E_s, D_s, red, mc, ext, n_out, E_out, x1, x2, idle_work, modes, range_err);
}

// ------------------------------------------------------------------- push step
// One Yoshida sub-stage for one particle (integration.py:22-47 with f = [v; -E], pic.py:125-129):
//   kick  : v <- v + (d * (-E_p)) * dt     E_p gathered at the CURRENT (wrapped) position
//   drift : x <- x + (c * v) * dt          x itself is carried unwrapped between sub-stages
template <typename R>
__device__ __forceinline__ R drift(R x, R v, R cc, const PartConst<R>& c) {
    return RT<R>::add(x, RT<R>::mul(RT<R>::mul(cc, v), c.dt));
}

// CIC deposit in fixed point: the right weight is rounded to k fractional bits, W_r = llrint(w_r 2^k), and the left
// cell receives 2^k - W_r, so every particle deposits exactly 2^k: total charge is conserved to the bit and the sum
// over cells is independent of summation order.  (The reference's w_l differs from 1 - w_r by at most one rounding
// of 1.0, i.e. 1.1e-16 -- four orders below the stated density tolerance.)  See deposit_weights below.

// ---------------------------------------------------------------- fast path
// The functions above are the reference semantics with every rare case handled in place (far wrap, positions within
// rounding distance of a cell edge, indices outside the mesh, non-finite input).  Executed as is they cost ~120
// instructions per particle, a third of them branch / reconvergence overhead around cases that almost never occur.
// The hot loops therefore run `particle_fast`: straight-line code whose result is exact whenever, for the position
// it looks at, x * (1/dx) is farther than idx_thr from an integer and its floor is a cell of the mesh -- which
// together imply 0 <= x < L (no wrap needed) and floor(x * (1/dx)) == floor(x / dx).  Anything else (including NaN
// and Inf, through the unordered compare) raises one flag and the particle is redone from its original (x, v) by
// the full-semantics code in a cold block.  Same bits either way.
template <typename R>
__device__ __forceinline__ bool fast_cell(R xw, const PartConst<R>& c, int M, int& il, R& f) {
    // Bracket instead of threshold: floor(x a_lo) and floor(x a_hi) with a_lo/hi = (1/dx)(1 -+ eps) are two fused
    // multiply-adds (round-down, against the magic constant).  eps = 2^-49 (2^-22 in float32) is 16x (4x) the relative
    // distance between the correctly rounded quotient x/dx and x (1/dx), so when both floors agree every real number
    // in between -- x/dx as IEEE division rounds it included -- has that floor: il == floor(x/dx), bit for bit.  They
    // disagree only within eps of a cell edge (redone exactly on the careful path).  Three fp64-pipe instructions
    // (2 DFMA + the subtraction that turns the integer back into a double) where multiply, magic add, two
    // subtractions, an add and a compare were needed to measure the distance to the nearest integer.
    // (The high word of the upper product needs no test of its own: with the lower product in range, the upper one is
    //  at most a few units above it, so equal low words mean equal values.  NaN, infinities, negative and huge x all
    //  fail the test of the lower product.)
    bool ok_lo, ok_hi;
    R t_lo, t_hi;
    il = RT<R>::floor_mul_magic(xw, c.inv_lo, ok_lo, t_lo);
    const int ih = RT<R>::floor_mul_magic(xw, c.inv_hi, ok_hi, t_hi);
    f = RT<R>::unmagic(t_lo);
    return !ok_lo | (il != ih) | ((unsigned)il >= (unsigned)M);
}

// the three TSC weights from the in-cell distance d = (x - m dx)/dx, formulas followed literally (interpolate.py:28-32)
template <typename R>
__device__ __forceinline__ void tsc_weights(R d, R& wl, R& wm, R& wr) {
    const R t = RT<R>::sub((R)1.5, d), u = RT<R>::sub(d, (R)1), q = RT<R>::sub(d, (R)0.5);
    wl = RT<R>::mul((R)0.5, RT<R>::mul(t, t));
    wm = RT<R>::sub((R)0.75, RT<R>::mul(u, u));
    wr = RT<R>::mul((R)0.5, RT<R>::mul(q, q));
}

// gathered field at a particle of cell il (already inside the mesh) with in-cell numerators, util.py:106 / :110.
// Split into the table read (fetch_field) and the arithmetic (apply_gather) so that the texture-gather kernels can issue
// the reads of a whole tile before they consume the first of them (the texture pipe's latency is several times a
// shared load's); gather_field is the two back to back.
template <typename R, int IP> struct Fetched { typename PairT<R>::type e1, e0; };   // e1 = (E_m, E_{m+1}); e0 = (E_{m-1}, E_m): TSC only

template <typename R, int IP, typename ET>
__device__ __forceinline__ Fetched<R, IP> fetch_field(int il, const ET& E_s, int M) {
    Fetched<R, IP> q;
    if (IP == IP_TSC) q.e0 = E_s[il == 0 ? M - 1 : il - 1];
    q.e1 = E_s[il];
    return q;
}

template <typename R, int IP, bool EXACT_W>
__device__ __forceinline__ R apply_gather(R x, R f, const Fetched<R, IP>& q, const PartConst<R>& c) {
    R nr = RT<R>::sub(x, RT<R>::mul(f, c.dx));
    R wr = EXACT_W ? RT<R>::div(nr, c.dx) : RT<R>::mul(nr, c.inv_dx);
    if (IP == IP_TSC) {
        R wl, wm, wrr;
        tsc_weights<R>(wr, wl, wm, wrr);
        return RT<R>::add(RT<R>::add(RT<R>::mul(wl, q.e0.x), RT<R>::mul(wm, q.e1.x)), RT<R>::mul(wrr, q.e1.y));
    }
    R fr = RT<R>::add(f, (R)1);
    R nl = RT<R>::sub(RT<R>::mul(fr, c.dx), x);
    R wl = EXACT_W ? RT<R>::div(nl, c.dx) : RT<R>::mul(nl, c.inv_dx);
    return RT<R>::add(RT<R>::mul(wl, q.e1.x), RT<R>::mul(wr, q.e1.y));
}

template <typename R, int IP, bool EXACT_W, typename ET>
__device__ __forceinline__ R gather_field(R x, int il, R f, const ET& E_s, const PartConst<R>& c, int M) {
    return apply_gather<R, IP, EXACT_W>(x, f, fetch_field<R, IP>(il, E_s, M), c);
}

// fixed-point deposit weights of a particle at wrapped position xw in cell with floor f:
// CIC: Wa = W_r (Wb unused);  TSC: Wa = W_l, Wb = W_r
template <typename R, int IP, bool EXACT_W>
__device__ __forceinline__ void deposit_weights(R xw, R f, const PartConst<R>& c, const MeshConst& mc, long long& Wa,
                                                long long& Wb) {
    R nr = RT<R>::sub(xw, RT<R>::mul(f, c.dx));
    R wr = EXACT_W ? RT<R>::div(nr, c.dx) : RT<R>::mul(nr, c.inv_dx);
    if (IP == IP_TSC) {
        R wl, wm, wrr;
        tsc_weights<R>(wr, wl, wm, wrr);
        Wa = fix_weight((double)wl, mc.fix_scale);
        Wb = fix_weight((double)wrr, mc.fix_scale);
    } else {
        Wa = fix_weight((double)wr, mc.fix_scale);
        Wb = 0;
    }
}

template <typename R, int IP, bool KICK, bool MOVE, bool EXACT_W, typename ET>
__device__ __forceinline__ bool particle_fast(R x, R v, R& xn, R& vn, int& il_dep, long long& Wa, long long& Wb,
                                              const ET& E_s, R cc, R dd, const PartConst<R>& c, const MeshConst& mc) {
    bool slow = false;
    vn = v;
    if (KICK) {
        int il; R f;
        slow = fast_cell<R>(x, c, mc.M, il, f);
        il = (int)min((unsigned)il, (unsigned)(mc.M - 1));      // keep the gather address legal on the slow path
        R Ep = gather_field<R, IP, EXACT_W>(x, il, f, E_s, c, mc.M);
        vn = RT<R>::add(v, RT<R>::mul(RT<R>::mul(dd, -Ep), c.dt));
    }
    xn = MOVE ? RT<R>::add(x, RT<R>::mul(RT<R>::mul(cc, vn), c.dt)) : x;
    R f2;
    slow |= fast_cell<R>(xn, c, mc.M, il_dep, f2);
    deposit_weights<R, IP, EXACT_W>(xn, f2, c, mc, Wa, Wb);
    return slow;
}

// Reference semantics, every case handled (cold block; reached by ~1e-5 .. 1e-2 of the particles).
// wrap_state: the position written back is the wrapped one (pic.py:139 / util.py:51), else the unwrapped drift.
template <typename R, int IP, bool KICK, bool MOVE, bool EXACT_W, typename ET>
__device__ __forceinline__ void particle_careful(R& x, R& v, int& il_dep, long long& Wa, long long& Wb,
                                                 const ET& E_s, R cc, R dd, const PartConst<R>& c, const MeshConst& mc,
                                                 bool wrap_state, unsigned& err) {
    if (KICK) {
        R xk = wrap_pos<R>(x, c, err), fk;
        int ik = cell_index<R>(xk, c, mc.M, fk, err);
        R Ep = gather_field<R, IP, EXACT_W>(xk, ik, fk, E_s, c, mc.M);
        v = RT<R>::add(v, RT<R>::mul(RT<R>::mul(dd, -Ep), c.dt));
    }
    if (MOVE) x = drift<R>(x, v, cc, c);
    R xw = wrap_pos<R>(x, c, err);
    R f;
    il_dep = cell_index<R>(xw, c, mc.M, f, err);
    deposit_weights<R, IP, EXACT_W>(xw, f, c, mc, Wa, Wb);
    if (wrap_state) x = xw;
}

// 64-bit sum over the full warp of values below 2^51 (two 25/26-bit halves, each REDUX sum stays below 2^31)
__device__ __forceinline__ unsigned long long warp_sum_u51(unsigned long long W) {
    const unsigned a = __reduce_add_sync(0xffffffffu, (unsigned)(W & 0x1FFFFFFu));
    const unsigned b = __reduce_add_sync(0xffffffffu, (unsigned)(W >> 25));
    return (unsigned long long)a + ((unsigned long long)b << 25);
}

// Deposit of one particle per lane, all 32 lanes of the warp calling together.  When every lane hits the same
// cell (cell-sorted particles) the warp sums its weights with REDUX instructions and one lane issues the
// atomics -- instead of 32 serialised same-address atomics.  Integer sums: grouping never changes the result.
template <int IP, typename H>
__device__ __forceinline__ void deposit_one(H& hist, int il, long long Wa, long long Wb, long long one) {
    if constexpr (IP == IP_TSC) hist.deposit_group(il, (unsigned long long)Wa, (unsigned long long)Wb, 1u);
    else hist.deposit(il, Wa, one);
}
template <int IP, typename H>
__device__ __forceinline__ void deposit_full_warp(H& hist, int il, long long Wa, long long Wb, long long one) {
    const int il0 = __shfl_sync(0xffffffffu, il, 0);
    if (__all_sync(0xffffffffu, il == il0)) {
        const unsigned long long SA = warp_sum_u51((unsigned long long)Wa);
        unsigned long long SB = 0;
        if constexpr (IP == IP_TSC) SB = warp_sum_u51((unsigned long long)Wb);
        if ((threadIdx.x & 31) == 0) {
            if constexpr (IP == IP_TSC) hist.deposit_group(il, SA, SB, 32u);
            else hist.deposit_group(il, SA, 32u, one);
        }
    } else {
        deposit_one<IP>(hist, il, Wa, Wb, one);
    }
}

// Aggregation hint.  Warp-aggregating the deposit only pays when the particles are cell-sorted (every lane of a warp in
// the same cell, where 32 same-address atomics would serialise); for particles in random order the uniformity vote is
// pure overhead -- 4-5 instructions per deposit, 4 % of the streaming passes.  Sortedness is a property of the particle
// ORDER, so one probe per tile is enough: the first particle of a tile votes (PROBE) and its verdict `agg` (warp-
// uniform) decides whether the other deposits of the tile take the aggregated route (which re-checks per particle,
// so the hint can only cost time, never correctness) or go straight to their own atomics.
template <int IP, bool PROBE, typename H>
__device__ __forceinline__ void deposit_hinted(H& hist, int il, long long Wa, long long Wb, long long one, bool& agg) {
#ifdef PIC_NO_VOTE
    deposit_one<IP>(hist, il, Wa, Wb, one);
#else
    if (PROBE) agg = __all_sync(0xffffffffu, il == __shfl_sync(0xffffffffu, il, 0));
    if (agg) deposit_full_warp<IP>(hist, il, Wa, Wb, one);
    else deposit_one<IP>(hist, il, Wa, Wb, one);
#endif
}

// One sub-stage for one particle: fast path, careful fallback, deposit.  x, v are updated in place.
// FULL_WARP: the caller guarantees that all 32 lanes of the warp execute this call (enables the aggregated deposit,
// steered by the per-tile hint `agg`; PROBE: this call refreshes the hint).
// ET: the gather table -- a pointer to the shared-memory pair table, or a TexTable (anything with operator[](cell)).
template <typename R, int IP, bool KICK, bool MOVE, bool EXACT_W, bool FULL_WARP, typename H, bool PROBE = false,
          typename ET = const typename PairT<R>::type*>
__device__ __forceinline__ void particle_substage(R& x, R& v, H& hist, const ET& E_s,
                                                  R cc, R dd, const PartConst<R>& c, const MeshConst& mc,
                                                  bool wrap_state, unsigned& err, bool* agg = nullptr) {
    R xn, vn; int il; long long Wa, Wb;
    const bool slow = particle_fast<R, IP, KICK, MOVE, EXACT_W>(x, v, xn, vn, il, Wa, Wb, E_s, cc, dd, c, mc);
    if (__builtin_expect(slow, 0)) {
        particle_careful<R, IP, KICK, MOVE, EXACT_W>(x, v, il, Wa, Wb, E_s, cc, dd, c, mc, wrap_state, err);
    } else {
        x = xn; v = vn;                                   // inside [0, L): wrapped == unwrapped
    }
    if (FULL_WARP) {
        __syncwarp();                                     // reconverge after the cold branch before the warp-wide part
        deposit_hinted<IP, PROBE>(hist, il, Wa, Wb, mc.fix_one, *agg);
    } else {
        deposit_one<IP>(hist, il, Wa, Wb, mc.fix_one);
    }
}

// The same sub-stage for a particle whose gather was issued ahead of time (texture-gather kernels): `f`, `slow_gather`
// and `q` are what fast_cell and fetch_field returned for the particle's position x.  Same arithmetic, same bits.
template <typename R, int IP, bool MOVE, bool EXACT_W, bool FULL_WARP, typename H, bool PROBE, typename ET>
__device__ __forceinline__ void particle_substage_pre(R& x, R& v, R f, bool slow_gather, const Fetched<R, IP>& q, H& hist,
                                                      const ET& E_s, R cc, R dd, const PartConst<R>& c, const MeshConst& mc,
                                                      bool wrap_state, unsigned& err, bool* agg) {
    const R Ep = apply_gather<R, IP, EXACT_W>(x, f, q, c);
    const R vn = RT<R>::add(v, RT<R>::mul(RT<R>::mul(dd, -Ep), c.dt));
    const R xn = MOVE ? RT<R>::add(x, RT<R>::mul(RT<R>::mul(cc, vn), c.dt)) : x;
    int il; R f2; long long Wa, Wb;
    const bool slow = slow_gather | fast_cell<R>(xn, c, mc.M, il, f2);
    deposit_weights<R, IP, EXACT_W>(xn, f2, c, mc, Wa, Wb);
    if (__builtin_expect(slow, 0)) {
        particle_careful<R, IP, true, MOVE, EXACT_W>(x, v, il, Wa, Wb, E_s, cc, dd, c, mc, wrap_state, err);
    } else {
        x = xn; v = vn;
    }
    if (FULL_WARP) {
        __syncwarp();
        deposit_hinted<IP, PROBE>(hist, il, Wa, Wb, mc.fix_one, *agg);
    } else {
        deposit_one<IP>(hist, il, Wa, Wb, mc.fix_one);
    }
}

// Stage 0 of the NEXT env step, done while the particle is still in registers: the first Yoshida sub-stage has d = 0
// (integration.py:71), i.e. it is a pure drift x1 = x + (c0 v) dt of the state the current step just produced and
// depends on nothing else -- not on the next action, not on any field.  Only its DEPOSIT has to happen ahead of
// time (stage 1 needs the field of x1 before it can kick), so x1 is deposited into a second histogram here and
// thrown away; the stage-1 pass recomputes it from the stored state with the same three operations (bit-identical).
// This removes one whole pass over the particles from every env step: 96 instead of 120 bytes per particle-step.
template <typename R, int IP, bool EXACT_W, bool FULL_WARP, typename H>
__device__ __forceinline__ void next_stage0(R x_state, R v, H& hist_next, R c0, const PartConst<R>& c, const MeshConst& mc,
                                            unsigned& err, bool* agg = nullptr) {
    const R x1 = RT<R>::add(x_state, RT<R>::mul(RT<R>::mul(c0, v), c.dt));
    int il; R f; long long Wa, Wb;
    const bool slow = fast_cell<R>(x1, c, mc.M, il, f);
    deposit_weights<R, IP, EXACT_W>(x1, f, c, mc, Wa, Wb);
    if (__builtin_expect(slow, 0)) {
        const R xw = wrap_pos<R>(x1, c, err);
        il = cell_index<R>(xw, c, mc.M, f, err);
        deposit_weights<R, IP, EXACT_W>(xw, f, c, mc, Wa, Wb);
    }
    if (FULL_WARP) {
        __syncwarp();
        deposit_hinted<IP, false>(hist_next, il, Wa, Wb, mc.fix_one, *agg);
    } else {
        deposit_one<IP>(hist_next, il, Wa, Wb, mc.fix_one);
    }
}

// next_stage0 without the deposit: cell and fixed-point weights of x1 = x_state + (c0 v) dt (used by the packed float32
// path when a pair falls back to the scalar code; same statements as next_stage0 above)
template <typename R, int IP, bool EXACT_W>
__device__ __forceinline__ void next_stage0_compute(R x_state, R v, R c0, const PartConst<R>& c, const MeshConst& mc,
                                                    unsigned& err, int& il, long long& Wa, long long& Wb) {
    const R x1 = RT<R>::add(x_state, RT<R>::mul(RT<R>::mul(c0, v), c.dt));
    R f;
    const bool slow = fast_cell<R>(x1, c, mc.M, il, f);
    deposit_weights<R, IP, EXACT_W>(x1, f, c, mc, Wa, Wb);
    if (__builtin_expect(slow, 0)) {
        const R xw = wrap_pos<R>(x1, c, err);
        il = cell_index<R>(xw, c, mc.M, f, err);
        deposit_weights<R, IP, EXACT_W>(xw, f, c, mc, Wa, Wb);
    }
}

// ------------------------------------------------------------- packed float32 pairs
// float32 mode, CIC: TWO particles per floating-point instruction (sm_100's FFMA2 / FADD2 / FMUL2 on aligned register
// pairs; the 16-byte particle vector already holds four).  Every lane of a packed instruction is rounded exactly as the
// scalar instruction it replaces (rn everywhere, rd in the cell bracket; a - b is the single-rounding fma(b, -1, a)), so
// the pair path produces the same bits as the scalar path: the float32 passes execute the same ~80-130 instructions per
// particle as the float64 ones on half the bytes and are bound by instruction issue, not by HBM or the LSU pipe.
// Cell extraction, the table loads and the atomics stay per lane; a particle that needs the careful path is redone by the
// scalar code (its partner keeps the packed result).
namespace f32x2 {
__device__ __forceinline__ float2 bc(float a) { return make_float2(a, a); }
// a + b as fma(a, 1, b): the same single rounding, but ptxas (12.9) contracts a mul.rn.f32x2 feeding an add.rn.f32x2 into one
// FFMA2 despite the explicit rounding modifiers -- which would change the bits -- and it does not contract into an fma
__device__ __forceinline__ float2 add(float2 a, float2 b) { return __ffma2_rn(make_float2(-a.x, -a.y), bc(-1.0f), b); }
__device__ __forceinline__ float2 mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 sub(float2 a, float2 b) { return __ffma2_rn(b, bc(-1.0f), a); }   // round(a - b), one rounding

// fast_cell for two positions; bad0 / bad1: that lane must take the careful path
__device__ __forceinline__ void fast_cell(float2 xw, const PartConst<float>& c, int M, int& il0, int& il1, float2& f,
                                          bool& bad0, bool& bad1) {
    const float2 t_lo = __ffma2_rd(xw, bc(c.inv_lo), bc(12582912.0f));
    const float2 t_hi = __ffma2_rd(xw, bc(c.inv_hi), bc(12582912.0f));
    f = add(t_lo, bc(-12582912.0f));
    const int a0 = __float_as_int(t_lo.x), a1 = __float_as_int(t_lo.y);
    il0 = a0 - 0x4B400000; il1 = a1 - 0x4B400000;
    bad0 = ((a0 >> 22) != (0x4B400000 >> 22)) | (a0 != __float_as_int(t_hi.x)) | ((unsigned)il0 >= (unsigned)M);
    bad1 = ((a1 >> 22) != (0x4B400000 >> 22)) | (a1 != __float_as_int(t_hi.y)) | ((unsigned)il1 >= (unsigned)M);
}

__device__ __forceinline__ float2 drift(float2 x, float2 v, float cc, const PartConst<float>& c) {
    return add(x, mul(mul(bc(cc), v), bc(c.dt)));
}

// right weights of two wrapped positions with floors f (deposit_weights, CIC, EXACT_W = false)
__device__ __forceinline__ float2 weight_r(float2 xw, float2 f, const PartConst<float>& c) {
    return mul(sub(xw, mul(f, bc(c.dx))), bc(c.inv_dx));
}

// particle_fast<float, IP_CIC, true, true, false> for two particles: gather, kick, drift, cell + weight of the new
// position; slow0 / slow1: that particle must be redone by particle_careful
__device__ __forceinline__ void particle_fast(float2 x, float2 v, float2& xn, float2& vn, int& il0, int& il1, float2& wr_dep,
                                              const float2* __restrict__ E_s, float cc, float dd,
                                              const PartConst<float>& c, int M, bool& slow0, bool& slow1) {
    int g0, g1; float2 f;
    fast_cell(x, c, M, g0, g1, f, slow0, slow1);
    g0 = (int)min((unsigned)g0, (unsigned)(M - 1)); g1 = (int)min((unsigned)g1, (unsigned)(M - 1));
    const float2 e0 = E_s[g0], e1 = E_s[g1];
    const float2 wr = weight_r(x, f, c);                                        // (x - f dx) (1/dx)
    const float2 wl = mul(sub(mul(add(f, bc(1.0f)), bc(c.dx)), x), bc(c.inv_dx));   // ((f + 1) dx - x) (1/dx)
    const float2 Ep = add(mul(wl, make_float2(e0.x, e1.x)), mul(wr, make_float2(e0.y, e1.y)));
    vn = add(v, mul(mul(bc(dd), make_float2(-Ep.x, -Ep.y)), bc(c.dt)));
    xn = add(x, mul(mul(bc(cc), vn), bc(c.dt)));
    float2 f2; bool b0, b1;
    fast_cell(xn, c, M, il0, il1, f2, b0, b1);
    slow0 |= b0; slow1 |= b1;
    wr_dep = weight_r(xn, f2, c);
}
}  // namespace f32x2

// One thread asks the L2 to fetch a contiguous range ahead of the loads that will consume it (UBLKPF.L2): no registers,
// no shared memory, one instruction per range.  p 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Programmatic dependent launch: the kernels of a streaming step are launched with the programmatic-serialization
// attribute, so the next kernel's CTAs may be scheduled (and run whatever precedes their own wait) while this grid is
// still draining; `griddep_wait` returns once every grid this one depends on has completed and its writes are visible.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// streaming loads / stores (no reuse: keep the particle stream out of L1, evict-first in L2)
#ifndef PIC_LD_FLAVOR
#define PIC_LD_FLAVOR 0
#endif
#ifndef PIC_ST_FLAVOR
#define PIC_ST_FLAVOR 0
#endif
template <typename V> __device__ __forceinline__ V ld_stream(const V* p) {
#if PIC_LD_FLAVOR == 0
    return __ldcs(p);
#elif PIC_LD_FLAVOR == 1
    return *p;
#elif PIC_LD_FLAVOR == 2
    return __ldcg(p);
#else
    return __ldg(p);
#endif
}
// the same, but not allocated in the L1 at all: the texture-gather kernels keep the L1 for the field table
__device__ __forceinline__ double2 ld_stream_nol1(const double2* p) {
    double2 r;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld_stream_nol1(const float4* p) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
template <typename V> __device__ __forceinline__ void st_stream(V* p, const V& v) {
#if PIC_ST_FLAVOR == 0
    __stcs(p, v);
#elif PIC_ST_FLAVOR == 1
    *p = v;
#elif PIC_ST_FLAVOR == 2
    __stcg(p, v);
#else
    __stwt(p, v);
#endif
}

}  // namespace pic
