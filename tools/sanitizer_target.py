"""Tiny workload for compute-sanitizer (one tool per run): every kernel family once, small sizes."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pic_b200  # noqa: E402

L = 50.0
rng = np.random.RandomState(0)
for mode, N, M in (("resident", 1500, 64), ("streaming", 9001, 128)):
    for interpol in ("CIC", "TSC"):
        x = rng.uniform(-L, 2 * L, N); v = rng.normal(size=N)
        eng = pic_b200.Engine(N, M, L, 0.05, mode=mode, max_mode=2, interpol=interpol)
        act = pic_b200.E_field(L, M, 2)
        eng.set_actuator_basis(act.basis_cos, act.basis_sin)
        eng.enable_modes(2)
        eng.set_state(x[None], v[None])
        eng.step_coeffs(rng.uniform(-1, 1, (3, 1, 4)), 3)
        eng.step_mesh(0.1 * np.sin(np.arange(M))[None], 2)
        eng.get_state(); eng.get_cells(want_wm=True); eng.get_fields(); eng.get_trace(2); eng.get_modes()
        eng.phase_hist_config(-10, 10, 32); eng.phase_hist(); eng.kl_divergence()
        print(mode, interpol, "ok", eng.get_diag()[0, :2], "flags", eng.error_flags(), flush=True)
        eng.close()
e32 = pic_b200.Engine(2000, 64, L, 0.05, precision="f32")
e32.sample_state("two-stream", seed=3); e32.step_mesh(None, 2); print("f32 ok", e32.get_diag()[0, :2])
