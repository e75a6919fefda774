/* pic_b200.h -- C ABI of the B200-native 1D electrostatic PIC step.
 *
 * Drop-in boundary for ONE path of ZINZINBIN/Optimal-Control-1D-Electrostatic-Plasma: what
 * `PIC.update_state` (src/env/pic.py:131-146) does and the getters around it.  The reference is pure Python and
 * has no FFI of its own; these entry points are what a ctypes binding inside `src/env/pic.py` would call
 * (INTEGRATION.md shows that binding).  Plain pointers and sizes only -- no torch, numpy or C++ types.
 *
 * Conventions
 *   - every function returns 0 on success, a negative PIC_E* code on failure; pic_last_error(h) gives the text
 *     (pass NULL for errors of pic_create);
 *   - one handle = one CUDA device + one stream; a handle is not thread-safe;
 *   - work is enqueued asynchronously on the handle's stream; functions that fill HOST buffers synchronise it;
 *   - particle arrays are SoA, env-major: x[env][particle], v[env][particle], float64 on the host side whatever the
 *     device precision; mesh arrays are [env][N_mesh] float64;
 *   - there is no CPU fallback: without a CUDA device pic_create fails with PIC_ENODEVICE.
 */
#ifndef PIC_B200_H
#define PIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden; these are its exports */
#endif

#define PIC_B200_ABI_VERSION 4

enum { PIC_OK = 0, PIC_EINVAL = -1, PIC_ENODEVICE = -2, PIC_ECUDA = -3, PIC_ENOMEM = -4, PIC_ESTATE = -5,
       PIC_ENUMERIC = -6, PIC_ENCCL = -7, PIC_EUNSUPPORTED = -8 };

enum { PIC_F64 = 0, PIC_F32 = 1 };                       /* device precision of particles               */
enum { PIC_MODE_AUTO = 0, PIC_MODE_RESIDENT = 1, PIC_MODE_STREAMING = 2 };
enum { PIC_DEPOSIT_AUTO = -1, PIC_DEPOSIT_CAS64 = 0, PIC_DEPOSIT_SPLIT32 = 1 };
enum { PIC_INTERP_CIC = 0, PIC_INTERP_TSC = 1 };         /* interpol of PIC.__init__: interpolate.py:4 / :22 */

/* per-env diagnostics record (pic_get_diag / pic_get_trace), doubles */
enum { PIC_DIAG_KE = 0,        /* 0.5*sum(v^2)            src/env/util.py:144                       */
       PIC_DIAG_PE_MESH = 1,   /* 0.5*sum(E_mesh^2)*dx    src/control/objective.py:31 (reward term) */
       PIC_DIAG_SUM_V = 2,     /* sum(v)                                                            */
       PIC_DIAG_SUM_E2 = 3,    /* sum(E_mesh^2)                                                     */
       PIC_DIAG_REWARD = 4,    /* Reward.compute_reward of the transition into this state, reward.py:71-76 */
       PIC_DIAG_INPUT_E = 5,   /* sum(a^2)*L/4 of the applied action, reward.py:52-54 (0 for mesh-vector control) */
       PIC_DIAG_N = 6 };

typedef struct pic_handle pic_handle;
#define PIC_STREAM_OWN ((void*)(intptr_t)-1)

/* Mirrors the keyword set of PIC.__init__ (src/env/pic.py:13-27, as passed at run_wo_oc.py:79-92) that the step
 * itself needs, plus the batching / sharding / device knobs the reference does not have. */
typedef struct pic_config {
    int64_t n_particles;        /* particles per env stored on THIS handle (the local shard when sharded)      */
    int64_t n_particles_total;  /* N of the whole env, enters n0*L/N/dx (interpolate.py:18); 0 => n_particles  */
    int32_t n_mesh;             /* N_mesh                                                                      */
    int32_t n_envs;             /* independent envs advanced together (>= 1)                                   */
    double  n0;                 /* mean density                                                                */
    double  L;                  /* box length                                                                  */
    double  dt;                 /* time step AFTER the CFL clip of pic.py:71-72 (see pic_clip_dt)              */
    int32_t precision;          /* PIC_F64 | PIC_F32                                                           */
    int32_t mode;               /* PIC_MODE_*                                                                  */
    int32_t deposit;            /* PIC_DEPOSIT_*                                                               */
    int32_t fixed_bits;         /* fractional bits of the integer deposit; 0 => chosen from N / N_mesh         */
    int32_t exact_weights;      /* 1 => CIC weights with IEEE division as interpolate.py:11-12 (slower)        */
    int32_t device;             /* CUDA device ordinal                                                         */
    int32_t max_mode;           /* actuator modes m (src/control/actuator.py:5); 0 => mesh-vector actuation only */
    void*   stream;             /* cudaStream_t to enqueue on; NULL => the legacy default stream;
                                 * PIC_STREAM_OWN => the handle creates (and destroys) a non-blocking stream of its own */
    int32_t interpolation;      /* PIC_INTERP_CIC | PIC_INTERP_TSC (float64 + split32 deposit only)            */
} pic_config;

/* --- lifetime ---------------------------------------------------------------------------------------------- */
int pic_abi_version(void);
const char* pic_build_info(void);                    /* arch, flags, deposit flavours compiled in              */
double pic_clip_dt(double dt, int64_t n_particles_total, double L);     /* src/env/pic.py:71-72               */
int pic_create(const pic_config* cfg, pic_handle** out);                /* PIC.__init__ minus sampling         */
int pic_destroy(pic_handle* h);
const char* pic_last_error(const pic_handle* h);
int pic_set_stream(pic_handle* h, void* cuda_stream);
int pic_get_stream(pic_handle* h, void** cuda_stream);     /* the stream the handle enqueues on (e.g. to record events) */

/* --- state -------------------------------------------------------------------------------------------------- */
/* PIC.initialize after sampling (pic.py:66-77): takes x, v (host float64, [n_envs][n_particles]), wraps x in
 * place (util.py:51), deposits and solves the self-consistent field so every getter is valid. */
int pic_set_state(pic_handle* h, const double* x, const double* v);
int pic_set_state_device(pic_handle* h, const void* x_dev, const void* v_dev);   /* device precision, same layout */
/* Device-side replacement of the host samplers (src/env/dist.py) + perturbation (pic.py:68) + pic_set_state's
 * field build: kind 0 = bump-on-tail(a, v0, sigma), 1 = two-stream(v0, sigma).  Philox, counter = global particle
 * index, so shards of one env (global_offset, n_global) draw disjoint pieces of the same population; env_offset is
 * the global index of this handle's first env (env-sharded batches draw the same envs as the unsharded batch). */
int pic_sample_state(pic_handle* h, int32_t kind, double a, double v0, double sigma, double A, int32_t n_mode,
                     uint64_t seed, int64_t global_offset, int64_t n_global, int64_t env_offset);
/* PIC.get_state / .x / .v (pic.py:165-167): copies to host float64 and synchronises. */
int pic_get_state(pic_handle* h, double* x, double* v);
/* .n, .E_mesh (pic.py:101,117): [n_envs][n_mesh] each; either pointer may be NULL. */
int pic_get_fields(pic_handle* h, double* n, double* E_mesh);
/* fixed-point state density, [n_envs][n_mesh] uint64; value = sum(w) * 2^fixed_bits.  Bit-reproducible. */
int pic_get_density_fixed(pic_handle* h, uint64_t* rho, int32_t* fixed_bits);
/* [n_envs][PIC_DIAG_N] for the current state (get_energy / get_electric_energy are derived on the host side). */
int pic_get_diag(pic_handle* h, double* diag);
/* the same plus the sticky device error flags (pic_get_error_flags) in ONE synchronisation: what the host classes
 * call after every step so that a flagged step raises where the reference's np.bincount would (interpolate.py:16) */
int pic_get_diag_flags(pic_handle* h, double* diag, uint32_t* flags);
/* per-step records of the last pic_step_* call: [n_steps][n_envs][PIC_DIAG_N] */
int pic_get_trace(pic_handle* h, double* trace, int32_t n_steps);
/* cell index floor(x/dx), weights and gathered field of the current state (pic.py:102-123); [n_envs][n_particles]
 * each, any pointer may be NULL.  CIC: indx = indx_l, weight_m = 0.  TSC: indx = indx_m (indx_l/r = indx_m -+ 1 mod
 * N_mesh) and the three weights of interpolate.py:30-32. */
int pic_get_cells(pic_handle* h, int32_t* indx, double* weight_l, double* weight_r, double* E_particles,
                  double* weight_m);

/* --- the hot path -------------------------------------------------------------------------------------------- */
/* PIC.update_state(E_external) (pic.py:131-146) n_steps times with the same external mesh field.
 * E_ext: NULL (no control) or host float64 [n_envs][n_mesh] -- the (N_mesh,1) vector of util.py:102-103. */
int pic_step_mesh(pic_handle* h, const double* E_ext, int32_t n_steps);
/* Actuator basis tables of src/control/actuator.py:23-24, host float64 [n_mesh][m] each, m == cfg.max_mode. */
int pic_set_actuator_basis(pic_handle* h, const double* basis_cos, const double* basis_sin, int32_t m);
/* Fast path of E_field.compute_E (actuator.py:62) + update_state: coeffs host float64 [n_steps][n_envs][2m],
 * cos coefficients first (run_ddpg.py:283: action[:m] -> cos, action[m:] -> sin). */
int pic_step_coeffs(pic_handle* h, const double* coeffs, int32_t n_steps);
/* Same two calls with DEVICE pointers (no copies, nothing synchronises). */
int pic_step_mesh_device(pic_handle* h, const double* E_ext_dev, int32_t n_steps);
int pic_step_coeffs_device(pic_handle* h, const double* coeffs_dev, int32_t n_steps);
/* Device-side Reward (src/control/rl/reward.py:5-34,71-76): every step records
 *   alpha * max(1 - PE_mesh(s_t)/r_pe_n, 0) + beta * max(1 - (sum a_t^2 L/4)/r_ie_n, 0)
 * with s_t the state BEFORE the step (the trainers' convention, ddpg.py:425-455).  Defaults: alpha = beta = 1,
 * r_pe_n = 1, r_ie_n = 10 L/4 (the reference's n_actions = 10). */
int pic_set_reward(pic_handle* h, double alpha, double beta, double r_pe_n, double r_ie_n);
/* Spectral read-out: first n_modes (<= 8) Fourier modes of the self-consistent mesh field with the normalisation of
 * src/interpret/spectrum.py:17 (fft/N_mesh*2), k = 1..n_modes, layout [env][Re_1..Re_m, Im_1..Im_m].  The linear
 * feedback law of run_feedback.py and the behaviour-cloning target of ddpg.py:429-431 are (-Re, +Im) of these. */
int pic_enable_modes(pic_handle* h, int32_t n_modes);
int pic_get_modes(pic_handle* h, double* modes);                          /* current state                    */
int pic_get_mode_trace(pic_handle* h, double* modes, int32_t n_steps);    /* per step of the last pic_step_*  */
/* Phase-space diagnostics of src/control/objective.py:8-18 (the KL cost of Reward.compute_kl_divergence,
 * reward.py:43-46): 2-D histogram over x in [0, L] x v in [vmin, vmax] with nbins x nbins bins and np.histogram2d's
 * edge rules; f = counts * n0/dx/dv/N; KL = sum(rel_entr(f, f_eq + 1e-12)) * dx * dv.  f_eq is supplied by the caller
 * (estimate_f of the initial state, reward.py:18), host float64 [nbins][nbins]. */
int pic_phase_hist_config(pic_handle* h, double vmin, double vmax, int32_t nbins);
int pic_phase_hist(pic_handle* h, uint32_t* counts /* host [n_envs][nbins][nbins] or NULL */);
int pic_set_feq(pic_handle* h, const double* feq);
int pic_kl_divergence(pic_handle* h, double* kl /* host [n_envs] */);
int pic_sync(pic_handle* h);
/* sticky flags raised on the device (bit 0: cell index out of range, bit 1: non-finite position, bit 2: a peer
 * of the fused exchange did not arrive in time, bit 3: a cell's integer density overflowed -- the fixed-point format
 * leaves 8x the mean per-cell density of headroom, far more when N/N_mesh is small; lower fixed_bits for denser
 * clumps) */
int pic_get_error_flags(pic_handle* h, uint32_t* flags);
int pic_clear_error_flags(pic_handle* h);

/* --- zero-copy views for the policy side (device pointers owned by the handle) ------------------------------- */
typedef struct pic_device_views {
    void*   x;          /* [n_envs][ld] device precision */
    void*   v;
    int64_t ld;         /* env stride in elements */
    double* n;          /* [n_envs][n_mesh] */
    double* E_mesh;     /* [n_envs][n_mesh] */
    double* diag;       /* [n_envs][PIC_DIAG_N] */
    int32_t elem_size;  /* 8 or 4 */
} pic_device_views;
int pic_get_device_views(pic_handle* h, pic_device_views* out);
/* The views are READ-ONLY by contract: density, field, diagnostics and the pre-deposited first sub-stage of the next
 * step were all built from the x, v they alias.  A caller that does write particles through them (a reset kernel, a
 * torch policy) must call this before the next step: it re-runs what pic_set_state does after its copy (wrap,
 * deposit, field, next-step pre-deposit). */
int pic_refresh_fields(pic_handle* h);

/* --- particle sharding over several GPUs (one process per GPU) ------------------------------------------------ */
/* Each rank holds n_particles of the n_particles_total; after every sub-stage deposit the fixed-point density is
 * summed over ranks (integer sum => identical on every rank and independent of the rank count).
 * nccl_comm: an initialised ncclComm_t for this device; the library resolves ncclAllReduce from the already
 * loaded libnccl (the one torch ships) at run time. */
int pic_comm_init(pic_handle* h, void* nccl_comm, int32_t rank, int32_t world_size);
/* Or let the library build the communicator: rank 0 calls pic_nccl_unique_id, the 128 bytes are broadcast by any
 * means (torch.distributed object broadcast, MPI, a file), then every rank calls pic_comm_init_rank. */
int pic_nccl_unique_id(char* out128);
int pic_comm_init_rank(pic_handle* h, const char* id128, int32_t rank, int32_t world_size);
/* Fused exchange over NVLink peer memory -- no collective library in the step loop.  The caller provides, for every
 * rank r of the node, a device pointer to r's exchange buffer (pic_comm_exchange_words() 64-bit words each) and to
 * r's flag array (world 64-bit words, ZERO-initialised before the first step on every rank), all mapped into this
 * process (CUDA IPC, cuMem fabric handles, or torch symmetric memory).  The last CTA of every push kernel then
 * writes the rank's finished partial density straight into every peer's buffer and raises a flag; the next kernel's
 * prologue waits for the flags and sums the slots in rank order (integer sum: bit-identical on every rank and for
 * every GPU count).  world <= 8.  A peer that never arrives raises error flag bit 2 instead of hanging the GPU. */
int64_t pic_comm_exchange_words(const pic_handle* h, int32_t world_size);
int pic_comm_init_peer(pic_handle* h, int32_t rank, int32_t world_size, void* const* exchange_ptrs,
                       void* const* flag_ptrs, int64_t exchange_words);
/* Optional, after pic_comm_init_peer: the NVLS MULTICAST mapping of the same exchange buffers (cuMulticast* /
 * torch symmetric memory's multicast_ptr): the publishing CTA then issues ONE store per word, which the NVSwitch
 * replicates into the slot of every rank, instead of one peer store per word and rank.  Same slots, same flags, same
 * bits.  nullptr switches back to per-rank peer stores. */
int pic_comm_set_multicast(pic_handle* h, void* exchange_multicast);
/* Alternative used when the collective is driven from the host side (e.g. torch.distributed): run sub-stage
 * `stage` (1..3, 4 = finalize, -1 = init deposit; 0 is a no-op because the deposit of the drift-only stage 0 is done
 * ahead of time by stage 3 / init of the state it starts from, and stage 1 redoes the drift itself on load) and leave the local density in the buffer returned
 * by pic_stage_density for the caller to all-reduce in place: [n_envs][n_mesh] uint64 for stages 1 and 2,
 * 2 x [n_envs][n_mesh] (state density followed by the next step's stage-0 density) for stages 3 and -1. */
int pic_run_stage(pic_handle* h, int32_t stage);
int pic_stage_density(pic_handle* h, int32_t stage, uint64_t** rho_dev);
int pic_set_stage_actuation(pic_handle* h, const double* E_ext_dev, const double* coeffs_dev);

/* --- tuning / introspection ---------------------------------------------------------------------------------- */
int pic_get_launch_info(pic_handle* h, int32_t* mode, int32_t* threads, int32_t* per_thread, int32_t* grid_x,
                        int32_t* smem_bytes, int32_t* fixed_bits, int32_t* deposit);
/* streaming mode: threads per CTA, 16-byte vectors in flight per thread, CTAs per SM (0 = as many as fit);
 * resident mode: threads per CTA, CTAs PER ENV (1 = one CTA per env; 2, 4, 8 = the env spread over a thread-block
 * cluster whose CTAs exchange their histograms through distributed shared memory), third argument unused.
 * pic_get_launch_info reports the same quantity in *per_thread.  0 / negative = keep. */
int pic_set_tuning(pic_handle* h, int32_t threads, int32_t unroll_or_cluster, int32_t ctas_per_sm);
/* streaming mode: where the kick's field gather (util.py:106) reads the mesh field from.  PIC_GATHER_SHARED: every CTA
 * rebuilds the field in its prologue and gathers from shared memory.  PIC_GATHER_TEXTURE: the field is solved once per
 * sub-stage by a one-CTA launch and gathered through the texture pipe (off the LSU pipe that the deposit's shared
 * atomics saturate).  PIC_GATHER_TEXTURE_STAGES(mask): texture for the Yoshida stages whose bit is set (bit 0: stage
 * 1, bit 1: stage 2, bit 2: stage 3), shared for the others.  Identical bits whatever the route.  PIC_GATHER_AUTO
 * (default): texture where it was measured faster (stage 3 of large envs).  pic_get_gather reports the route in
 * effect (SHARED, TEXTURE or TEXTURE_STAGES(mask)). */
#define PIC_GATHER_AUTO 0
#define PIC_GATHER_SHARED 1
#define PIC_GATHER_TEXTURE 2
#define PIC_GATHER_TEXTURE_STAGES(mask) (0x10 | ((mask) & 0x7))
int pic_set_gather(pic_handle* h, int32_t route);
int pic_get_gather(pic_handle* h, int32_t* route);
/* streaming mode, one GPU: run whole env steps (any n_steps of a call) as ONE cooperative launch -- the three passes
 * separated by grid barriers, the finalize of a step on a CTA of its own beside the first pass of the next -- instead
 * of four launches per step.  Pays for mid-size envs (1e4 .. a few 1e7 particles), whose step is bound by what sits
 * between the passes, not by the particles.  Identical bits in x, v, densities and fields (the kinetic partial sums are
 * added over workers instead of CTAs, as with any other grid size).  PIC_COOP_AUTO (default): used for calls of two or
 * more steps on up to 2^24 particles per handle when available (a lone step gains nothing: a cooperative launch costs
 * more than a plain one, and there is no next pass to hide the finalize behind); PIC_COOP_ON: every call, fails with
 * PIC_EUNSUPPORTED when unavailable (it needs CIC, the split32 deposit, exact_weights = 0, the 1024 x 2 shape, the
 * shared-memory gather, no spectral read-out, no sharding).  pic_get_coop reports whether a multi-step call takes this
 * path and with how many pass CTAs per env. */
#define PIC_COOP_AUTO (-1)
#define PIC_COOP_OFF 0
#define PIC_COOP_ON 1
int pic_set_coop(pic_handle* h, int32_t mode);
int pic_get_coop(pic_handle* h, int32_t* in_effect, int32_t* workers);
int64_t pic_kernel_launch_count(const pic_handle* h);        /* kernels enqueued by this handle so far */

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* PIC_B200_H */
