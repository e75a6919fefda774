/* Driving the PIC step through the C ABI from plain C (no Python, no torch):
 *
 *   gcc -O2 -Iinclude examples/step_from_c.c -o build/step_from_c \
 *       -Loptimal-control-1d-electrostatic-plasma_b200/lib -lpic_b200 \
 *       -Wl,-rpath,$PWD/optimal-control-1d-electrostatic-plasma_b200/lib -lm
 *   build/step_from_c
 *
 * Same sequence as run_wo_oc.py:79-120 of the reference: build the env, initialise it from sampled particles,
 * call update_state in a loop and read the energies.  Prints total energy drift and exits 0 when it stays small.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "pic_b200.h"

static double uniform01(unsigned long long* s) {           /* xorshift64*, good enough for a demo */
    *s ^= *s >> 12; *s ^= *s << 25; *s ^= *s >> 27;
    return (double)((*s * 2685821657736338717ULL) >> 11) / 9007199254740992.0;
}

int main(void) {
    const int64_t N = 5000;
    const int32_t M = 250, steps = 200;
    const double L = 50.0, two_pi = 6.283185307179586;
    pic_config cfg = {0};
    cfg.n_particles = N; cfg.n_mesh = M; cfg.n_envs = 1; cfg.n0 = 1.0; cfg.L = L;
    cfg.dt = pic_clip_dt(0.05, N, L);                       /* pic.py:71-72 */
    cfg.deposit = PIC_DEPOSIT_AUTO;

    pic_handle* h = NULL;
    if (pic_create(&cfg, &h) != PIC_OK) { fprintf(stderr, "pic_create: %s\n", pic_last_error(NULL)); return 2; }

    double* x = malloc(sizeof(double) * N);
    double* v = malloc(sizeof(double) * N);
    unsigned long long seed = 42;
    for (int64_t i = 0; i < N; ++i) {                       /* two cold counter-streaming beams, Box-Muller */
        const double u1 = uniform01(&seed) + 1e-300, u2 = uniform01(&seed);
        x[i] = L * uniform01(&seed);
        v[i] = (i & 1 ? 3.0 : -3.0) + 0.5 * sqrt(-2.0 * log(u1)) * cos(two_pi * u2);
        v[i] *= 1.0 + 0.1 * sin(two_pi * x[i] / L);         /* pic.py:68 */
    }
    if (pic_set_state(h, x, v) != PIC_OK) { fprintf(stderr, "pic_set_state: %s\n", pic_last_error(h)); return 2; }

    double d[PIC_DIAG_N], e0 = 0.0, e = 0.0;
    for (int32_t s = 0; s <= steps; ++s) {
        if (s > 0 && pic_step_mesh(h, NULL, 1) != PIC_OK) { fprintf(stderr, "pic_step_mesh: %s\n", pic_last_error(h)); return 2; }
        pic_get_diag(h, d);
        e = d[PIC_DIAG_KE] + d[PIC_DIAG_PE_MESH] * (double)N / L;   /* util.py:119-146 */
        if (s == 0) e0 = e;
        if (s % 50 == 0) printf("step %4d  KE %.6e  PE %.6e  total %.9e\n", s, d[PIC_DIAG_KE], d[PIC_DIAG_PE_MESH] * (double)N / L, e);
    }
    uint32_t flags = 0;
    pic_get_error_flags(h, &flags);
    const double drift = fabs(e - e0) / e0;
    printf("relative energy drift over %d steps: %.3e, error flags %u, kernels launched %lld\n", steps, drift, flags,
           (long long)pic_kernel_launch_count(h));
    pic_destroy(h);
    free(x); free(v);
    return (drift < 1e-2 && flags == 0) ? 0 : 1;
}
