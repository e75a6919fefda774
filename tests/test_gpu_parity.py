"""Parity of the CUDA path (through the C ABI / the PIC duck-type) against the golden vectors of the reference and
against the oracle.  Tolerances (DESIGN.md "Numerics"): cell indices bit-exact; fp64 per step |dx|,|dv| <= 1e-12,
|dE_mesh| <= 1e-12*max|E|; 500-step energy trace rel 1e-9."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pic_oracle as O  # noqa: E402  (checker only)


def _engine(N, M, L, dt, **kw):
    from pic_b200 import Engine
    return Engine(N, M, L, dt, **kw)


CONFIGS = [dict(mode="resident", deposit="cas64"), dict(mode="resident", deposit="split32"),
           dict(mode="streaming", deposit="cas64"), dict(mode="streaming", deposit="split32")]


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: c["mode"] + "-" + c["deposit"])
@pytest.mark.parametrize("LM", [(50.0, 250), (50.0, 500), (50.0, 4096), (10.0, 64)])
def test_wrap_index_deposit_edge_cases(golden, cfg, LM):
    """np.mod(np.mod(x,L),L), floor(x/dx) and the CIC density on adversarial positions (cell edges +-1ulp, 0, -0, L,
    far outside) -- wrap and index bit-exact."""
    g = golden("deposit_edges")
    L, M = LM
    key = f"L{L:g}_M{M}"
    x = g[key + "_x"]
    N = x.shape[0]
    if cfg["mode"] == "resident" and M >= 4096:
        pytest.skip("16310 particles + a 4096-cell mesh do not fit one CTA's shared memory")
    eng = _engine(N, M, L, 0.01, **cfg)
    eng.set_state(x[None], np.zeros((1, N)))
    xs, _ = eng.get_state()
    assert np.array_equal(xs[0], g[key + "_xw"])
    il, wl, wr, _ = eng.get_cells()
    assert np.array_equal(il[0], g[key + "_cic_il"])
    assert np.abs(wl[0] - g[key + "_cic_wl"]).max() < 4e-16 * M      # 1/dx multiply vs divide: <= 1 ulp of x/dx
    assert np.abs(wr[0] - g[key + "_cic_wr"]).max() < 4e-16 * M
    n, _ = eng.get_fields()
    assert np.abs(n[0] - g[key + "_cic_n"]).max() < 1e-12 * g[key + "_cic_n"].max()
    assert eng.error_flags() == 0
    rho, k = eng.get_density_fixed()
    assert sum(int(r) for r in rho.ravel()) == N * (1 << k)           # every particle deposits exactly 2^k


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: c["mode"] + "-" + c["deposit"])
@pytest.mark.parametrize("name", ["bump_vb3", "twostream_vb3"])
def test_steps_vs_reference_golden(golden, cfg, name):
    g = golden(name)
    N, M, L, dt = int(g["N"]), int(g["N_mesh"]), float(g["L"]), float(g["dt"])
    eng = _engine(N, M, L, dt, **cfg)
    eng.set_state(g["t0_x"][None], g["t0_v"][None])
    n, E = eng.get_fields()
    assert np.abs(n[0] - g["t0_n"]).max() < 1e-12
    assert np.abs(E[0] - g["t0_E_mesh"]).max() < 1e-12 * np.abs(g["t0_E_mesh"]).max()
    d = eng.get_diag()[0]
    assert abs(d[1] * N / L - g["PE"][0]) < 1e-11 * g["PE"][0]
    assert abs(d[0] + d[1] * N / L - g["H"][0]) < 1e-12 * g["H"][0]
    done = 0
    for t in (1, 10):
        eng.step_mesh(None, t - done)
        done = t
        x, v = eng.get_state()
        assert np.abs(x[0] - g[f"t{t}_x"]).max() < 1e-12
        assert np.abs(v[0] - g[f"t{t}_v"]).max() < 1e-12
        il, wl, wr, Ep = eng.get_cells()
        assert np.array_equal(il[0], g[f"t{t}_indx_l"])
        n, E = eng.get_fields()
        assert np.abs(n[0] - g[f"t{t}_n"]).max() < 1e-12
        assert np.abs(E[0] - g[f"t{t}_E_mesh"]).max() < 1e-12 * np.abs(g[f"t{t}_E_mesh"]).max()
        assert np.abs(Ep[0] - g[f"t{t}_E"]).max() < 1e-12 * np.abs(g[f"t{t}_E"]).max()
        d = eng.get_diag()[0]
        assert abs(d[1] - g["PE_mesh"][t]) < 1e-11 * g["PE_mesh"][t]
        assert abs(d[0] + d[1] * N / L - g["H"][t]) < 1e-12 * g["H"][t]
        assert abs(d[2] - g["sum_v"][t]) < 1e-9
    assert eng.error_flags() == 0


@pytest.mark.parametrize("mode", ["resident", "streaming"])
@pytest.mark.parametrize("name,rate", [("bump_vb3", -0.001295), ("twostream_vb3", 0.021354), ("bump_vb5", 0.005568)])
def test_500_step_energy_trace_and_growth_rate(golden, mode, name, rate):
    g = golden(name)
    N, M, L, dt = int(g["N"]), int(g["N_mesh"]), float(g["L"]), float(g["dt"])
    eng = _engine(N, M, L, dt, mode=mode)
    eng.set_state(g["t0_x"][None], g["t0_v"][None])
    eng.step_mesh(None, 500)
    tr = eng.get_trace(500)[:, 0, :]
    pe = tr[:, 1]
    assert np.max(np.abs(pe - g["PE_mesh"][1:]) / g["PE_mesh"][1:]) < 1e-9
    H = tr[:, 0] + pe * N / L
    assert np.max(np.abs(H - g["H"][1:]) / g["H"][1:]) < 1e-10
    assert np.max(np.abs(tr[:, 2] - g["sum_v"][0])) < 1e-9                  # momentum conserved without control
    assert abs(O.growth_rate(pe, 50.0) - rate) < 1e-6
    x, v = eng.get_state()
    assert np.abs(x[0] - g["t500_x"]).max() < 1e-6                          # chaos-amplified ulp noise
    il, *_ = eng.get_cells(False, False)
    assert np.array_equal(il[0], g["t500_indx_l"])


@pytest.mark.parametrize("mode", ["resident", "streaming"])
@pytest.mark.parametrize("name,steps,m", [("bump_vb3_constctrl", 10, 3), ("bump_vb3_randctrl", 200, 3), ("sac_cfg", 40, 5)])
@pytest.mark.parametrize("path", ["mesh", "coeffs"])
def test_controlled_runs(golden, mode, name, steps, m, path):
    """E_external as the reference's (N_mesh,1) vector (util.py:102-103) and through the coefficient fast path."""
    g = golden(name)
    N, M, L, dt = int(g["N"]), int(g["N_mesh"]), float(g["L"]), float(g["dt"])
    eng = _engine(N, M, L, dt, mode=mode, max_mode=m)
    eng.set_state(g["t0_x"][None], g["t0_v"][None])
    pe = []
    if path == "mesh":
        for t in range(steps):
            eng.step_mesh(g["E_ext"][t][None], 1)
            pe.append(eng.get_diag()[0, 1])
    else:
        eng.set_actuator_basis(g["basis_cos"], g["basis_sin"])
        eng.step_coeffs(g["coeffs"][:steps, None, :], steps)
        pe = list(eng.get_trace(steps)[:, 0, 1])
    pe = np.array(pe)
    assert np.max(np.abs(pe - g["PE_mesh"][1:steps + 1]) / g["PE_mesh"][1:steps + 1]) < 1e-9
    x, v = eng.get_state()
    tol = 1e-12 if steps <= 10 else 1e-8
    assert np.abs(x[0] - g[f"t{steps}_x"]).max() < tol
    assert np.abs(v[0] - g[f"t{steps}_v"]).max() < tol
    il, *_ = eng.get_cells(False, False)
    assert np.array_equal(il[0], g[f"t{steps}_indx_l"])
    # reward of every transition uses the PRE-step field energy (ddpg.py:455, reward.py:71-76)
    pe_pre = np.concatenate([[g["PE_mesh"][0]], pe[:-1]])
    r = [O.reward(pe_pre[t], g["coeffs"][t], L) for t in range(steps)]
    assert np.max(np.abs(np.array(r) - g["rewards"][:steps])) < 1e-9


def test_dt_clip_case(golden):
    g = golden("clip_dt")
    N, M, L, dt = int(g["N"]), int(g["N_mesh"]), float(g["L"]), float(g["dt"])
    import pic_b200
    assert pic_b200._lib.load().pic_clip_dt(0.1, N, L) == dt
    eng = _engine(N, M, L, dt)
    assert eng.launch_info()["mode"] == "streaming"
    eng.set_state(g["t0_x"][None], g["t0_v"][None])
    eng.step_mesh(None, 5)
    x, v = eng.get_state()
    assert np.abs(x[0] - g["t5_x"]).max() < 1e-12
    assert np.abs(v[0] - g["t5_v"]).max() < 1e-12
    d = eng.get_diag()[0]
    assert abs(d[0] + d[1] * N / L - g["H"][5]) < 1e-12 * g["H"][5]


def test_bitwise_reproducibility_across_kernels(golden):
    """The integer deposit makes the density -- and with it x, v -- independent of kernel flavour, thread count and
    CTA count: resident vs streaming, cas64 vs split32, different launch shapes give IDENTICAL bits."""
    g = golden("bump_vb3")
    N, M, L, dt = int(g["N"]), int(g["N_mesh"]), float(g["L"]), float(g["dt"])
    outs = []
    variants = [dict(mode="resident", deposit="cas64"), dict(mode="resident", deposit="split32"),
                dict(mode="streaming", deposit="cas64"), dict(mode="streaming", deposit="split32"),
                dict(mode="resident", deposit="split32"), dict(mode="resident", deposit="cas64"),
                dict(mode="resident", deposit="split32")]
    # (threads, unroll [streaming only], ctas per SM)
    tunings = [None, None, (512, 1, 1), (256, 4, 2), (1024, 0, -1), (256, 0, -1), (512, 0, -1)]
    for cfg, tune in zip(variants, tunings):
        eng = _engine(N, M, L, dt, **cfg)
        if tune:
            eng.set_tuning(*tune)
        eng.set_state(g["t0_x"][None], g["t0_v"][None])
        eng.step_mesh(None, 25)
        x, v = eng.get_state()
        rho, k = eng.get_density_fixed()
        outs.append((x, v, rho))
    for o in outs[1:]:
        assert np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1]) and np.array_equal(o[2], outs[0][2])


@pytest.mark.parametrize("N,M", [(1_000_003, 4096), (300_000, 1000)])
def test_streaming_large_vs_oracle(N, M):
    """Ragged N (odd, not a multiple of any tile) against the lean oracle for two steps with an external field."""
    rng = np.random.RandomState(5)
    L = 50.0
    x = rng.uniform(0, L, N)
    v = rng.normal(0, 1, N) + 3.0 * (rng.uniform(size=N) < 0.2)
    dt = O.clip_dt(0.1, N, L)
    p = O.PicParams(N=N, N_mesh=M, n0=1.0, L=L, dt=dt)
    ext = 0.3 * np.sin(2 * np.pi * np.arange(M) / M) + 0.1
    eng = _engine(N, M, L, dt, mode="streaming")
    eng.set_state(x[None], v[None])
    xo, vo = x, v
    for _ in range(2):
        o = O.step(xo, vo, p, ext)
        xo, vo = o["x"], o["v"]
    eng.step_mesh(ext[None], 2)
    xg, vg = eng.get_state()
    assert np.abs(xg[0] - xo).max() < 1e-12
    assert np.abs(vg[0] - vo).max() < 1e-12
    il, *_ = eng.get_cells(False, False)
    assert np.array_equal(il[0], o["indx_l"])
    n, E = eng.get_fields()
    assert np.abs(E[0] - o["E_mesh"]).max() < 1e-11 * max(1.0, np.abs(o["E_mesh"]).max())
    assert eng.error_flags() == 0


def test_batched_envs_match_single_env(golden):
    """Env b of a batch is bit-identical to the same env advanced alone (same kernel, same integer deposit)."""
    from pic_b200 import BatchedPIC, Engine
    from pic_b200.dist import BumpOnTail
    B, N, M, L = 6, 5000, 250, 50.0
    bp = BatchedPIC(B, N=N, N_mesh=M, L=L, dt=0.05, max_mode=3)
    xs, vs = bp.reset_from_sampler(lambda: BumpOnTail(a=0.2, v0=3.0, sigma=1.0, n_samples=N, L=L), seed=42)
    rng = np.random.RandomState(0)
    acts = rng.uniform(-1, 1, size=(7, B, 6))
    out = bp.step(acts, n_steps=7)
    state = bp.get_state()
    bc, bs = O.actuator_basis(L, M, 3)
    for b in (0, 3, 5):
        e1 = Engine(N, M, L, 0.05, max_mode=3)
        e1.set_actuator_basis(bc, bs)
        e1.set_state(xs[b][None], vs[b][None])
        e1.step_coeffs(acts[:, b:b + 1, :], 7)
        x1, v1 = e1.get_state()
        assert np.array_equal(state[b, :N], x1[0]) and np.array_equal(state[b, N:], v1[0])
        assert np.array_equal(out["pe_mesh"][:, b], e1.get_trace(7)[:, 0, 1])
    # env 0 against the oracle (seed 42 + 0 is the runner's own sample #1; any seed works for parity)
    p = O.PicParams(N=N, N_mesh=M, n0=1.0, L=L, dt=0.05)
    xo, vo = xs[0], vs[0]
    pe_pre = O.pe_mesh(O.mesh_field_of_state(xo, p), p.dx)
    for t in range(7):
        o = O.step(xo, vo, p, O.actuator_field(bc, bs, acts[t, 0, :3], acts[t, 0, 3:]))
        assert abs(out["reward"][t, 0] - O.reward(pe_pre, acts[t, 0], L)) < 1e-9
        xo, vo = o["x"], o["v"]
        pe_pre = O.pe_mesh(o["E_mesh"], p.dx)
        assert abs(out["pe_mesh"][t, 0] - pe_pre) < 1e-10 * max(1.0, pe_pre)
    assert np.abs(state[0, :N] - xo).max() < 1e-11 and np.abs(state[0, N:] - vo).max() < 1e-11


def test_pic_class_is_a_drop_in(golden):
    """The reference's constructor sequence through our classes: seed 42, dist ctor (sample discarded), PIC()
    (second sample, perturbation, fields), then the runner loop of run_wo_oc.py:108-122."""
    from pic_b200 import PIC
    from pic_b200.dist import BumpOnTail
    g = golden("bump_vb3")
    np.random.seed(42)                                   # what importing src/env/pic.py does (pic.py:12)
    dist = BumpOnTail(a=0.2, v0=3.0, sigma=1.0, n_samples=5000, L=50.0)
    sim = PIC(N=5000, N_mesh=250, n0=1.0, L=50.0, dt=0.1, tmin=0.0, tmax=50.0, gamma=5.0, A=0.1, n_mode=2,
              interpol="CIC", init_dist=dist)
    assert np.array_equal(sim.x[:, 0], g["t0_x"]) and np.array_equal(sim.v[:, 0], g["t0_v"])
    assert sim.x.shape == (5000, 1) and sim.get_state().shape == (10000, 1)
    assert abs(sim.get_energy() - g["H"][0]) < 1e-12 * g["H"][0]
    for t in range(1, 11):
        sim.update_state(None)
        assert abs(sim.get_energy() - g["H"][t]) < 1e-12 * g["H"][t]
        assert abs(sim.get_electric_energy() - g["PE"][t]) < 1e-11 * g["PE"][t]
    assert np.abs(sim.x[:, 0] - g["t10_x"]).max() < 1e-12
    assert np.array_equal(sim.indx_l[:, 0], g["t10_indx_l"].astype(np.int64))
    assert np.abs(sim.E_mesh[:, 0] - g["t10_E_mesh"]).max() < 1e-12
    # derived attributes: phi solves the same discretisation (laplacian @ phi = n - n0) and -grad @ phi is E_mesh
    phi = sim.phi_mesh
    assert np.abs(sim.laplacian @ phi - (sim.n - sim.n0).reshape(-1, 1)).max() < 1e-9
    assert np.abs(-(sim.grad @ phi) - sim.E_mesh).max() < 1e-11
    with pytest.raises(ValueError):
        PIC(N=100, N_mesh=10, interpol="NGP", init_dist=dist)


@pytest.mark.parametrize("mode", ["resident", "streaming"])
def test_device_reward_and_spectral_modes(golden, mode):
    """Reward.compute_reward (reward.py:71-76) and the first Fourier modes of E_mesh
    (spectrum.py:17: fft/N_mesh*2) come out of the step kernels."""
    g = golden("bump_vb3_randctrl")
    N, M, L, dt, m, steps = int(g["N"]), int(g["N_mesh"]), float(g["L"]), float(g["dt"]), 3, 60
    eng = _engine(N, M, L, dt, mode=mode, max_mode=m)
    eng.set_actuator_basis(g["basis_cos"], g["basis_sin"])
    eng.enable_modes(m)
    eng.set_state(g["t0_x"][None], g["t0_v"][None])
    Ek0 = np.fft.fft(g["t0_E_mesh"]) / M * 2.0
    m0 = eng.get_modes()[0]
    assert np.abs(m0[:m] - Ek0[1:m + 1].real).max() < 1e-12 and np.abs(m0[m:] - Ek0[1:m + 1].imag).max() < 1e-12
    eng.step_coeffs(g["coeffs"][:steps, None, :], steps)
    tr = eng.get_trace(steps)[:, 0, :]
    assert np.max(np.abs(tr[:, 4] - g["rewards"][:steps])) < 1e-9                       # reward on the PRE-step state
    ie = np.sum(g["coeffs"][:steps] ** 2, axis=1) * L * 0.25
    assert np.max(np.abs(tr[:, 5] - ie)) < 1e-12
    mt = eng.get_mode_trace(steps)[:, 0, :]
    _, E = eng.get_fields()
    Ek = np.fft.fft(E[0]) / M * 2.0
    assert np.abs(mt[-1, :m] - Ek[1:m + 1].real).max() < 1e-12 and np.abs(mt[-1, m:] - Ek[1:m + 1].imag).max() < 1e-12
    assert np.array_equal(eng.get_modes()[0], mt[-1])
    # mesh-vector control carries no coefficients: input-energy term is zero, reward = r_pe + beta
    eng.step_mesh(g["E_ext"][0][None], 1)
    d = eng.get_diag()[0]
    assert d[5] == 0.0 and abs(d[4] - (max(1 - tr[-1, 1], 0) + 1.0)) < 1e-12


def test_phase_space_histogram_and_kl(golden):
    """estimate_f / estimate_KL_divergence (src/control/objective.py:8-18) on the device, checked
    against np.histogram2d + scipy rel_entr on the same state (counts bit-exact, KL to 1e-12)."""
    from scipy.special import rel_entr
    g = golden("twostream_vb3")
    N, M, L, dt = int(g["N"]), int(g["N_mesh"]), float(g["L"]), float(g["dt"])
    vmin, vmax = -25.0, 25.0
    eng = _engine(N, M, L, dt)
    eng.phase_hist_config(vmin, vmax, M)

    def ref_f(x, v):
        d, _, _ = np.histogram2d(x, v, bins=[M, M], density=False, range=np.array([[0, L], [vmin, vmax]]))
        return d

    eng.set_state(g["t0_x"][None], g["t0_v"][None])
    c0 = eng.phase_hist()[0]
    assert np.array_equal(c0, ref_f(g["t0_x"], g["t0_v"]).astype(np.uint32))
    dx, dv = L / M, (vmax - vmin) / M
    feq = ref_f(g["raw_x_init"], g["raw_v_init"]) * (1.0 / dx / dv / N)        # Reward.__init__: estimate_f(init_state)
    eng.set_feq(feq)
    eng.step_mesh(None, 40)
    x, v = eng.get_state()
    f = ref_f(x[0], v[0]) * (1.0 / dx / dv / N)
    assert np.array_equal(eng.phase_hist()[0], ref_f(x[0], v[0]).astype(np.uint32))
    assert np.array_equal(eng.estimate_f()[0], f)
    kl_ref = np.sum(rel_entr(f, feq + 1e-12)) * dx * dv
    assert abs(eng.kl_divergence()[0] - kl_ref) < 1e-12 * max(1.0, abs(kl_ref))
    # edge rules: values exactly on bin edges, on the right-most edge, and outliers
    edges_v = np.linspace(vmin, vmax, M + 1)
    xs = np.concatenate([np.linspace(0, L, M + 1)[:-1], np.nextafter(np.linspace(0, L, M + 1)[1:], 0), [0.0] * 6])
    vs = np.concatenate([edges_v[:-1], edges_v[1:], [vmax, vmin, np.nextafter(vmax, 100), np.nextafter(vmin, -100), 30.0, -30.0]])
    e2 = _engine(xs.shape[0], M, L, dt)
    e2.phase_hist_config(vmin, vmax, M)
    e2.set_state(xs[None], vs[None])
    xx, vv = e2.get_state()
    assert np.array_equal(e2.phase_hist()[0], ref_f(xx[0], vv[0]).astype(np.uint32))


@pytest.mark.parametrize("mode", ["resident", "streaming"])
@pytest.mark.parametrize("name,steps,ctrl", [("bump_vb3_tsc", 40, False), ("twostream_tsc_ctrl", 20, True)])
def test_tsc_interpolation(golden, mode, name, steps, ctrl):
    """interpol="TSC" (src/env/interpolate.py:22-44, gather util.py:110) against the reference."""
    g = golden(name)
    N, M, L, dt = int(g["N"]), int(g["N_mesh"]), float(g["L"]), float(g["dt"])
    eng = _engine(N, M, L, dt, mode=mode, max_mode=3, interpol="TSC")
    eng.set_state(g["t0_x"][None], g["t0_v"][None])
    n, E = eng.get_fields()
    assert np.abs(n[0] - g["t0_n"]).max() < 1e-12 and np.abs(E[0] - g["t0_E_mesh"]).max() < 1e-12
    if ctrl:
        eng.set_actuator_basis(g["basis_cos"], g["basis_sin"])
        eng.step_coeffs(g["coeffs"][:1, None, :], 1)
    else:
        eng.step_mesh(None, 1)
    x, v = eng.get_state()
    assert np.abs(x[0] - g["t1_x"]).max() < 1e-12 and np.abs(v[0] - g["t1_v"]).max() < 1e-12
    im, wl, wr, Ep, wm = eng.get_cells(want_wm=True)
    assert np.array_equal(im[0], g["t1_indx_m"])
    assert np.abs(wm[0] - g["t1_weight_m"]).max() < 1e-13
    assert np.abs(Ep[0] - g["t1_E"]).max() < 1e-12 * max(1.0, np.abs(g["t1_E"]).max())
    if ctrl:
        eng.step_coeffs(g["coeffs"][1:steps, None, :], steps - 1)
    else:
        eng.step_mesh(None, steps - 1)
    x, v = eng.get_state()
    assert np.abs(x[0] - g[f"t{steps}_x"]).max() < 1e-10 and np.abs(v[0] - g[f"t{steps}_v"]).max() < 1e-10
    im, *_ = eng.get_cells(False, False)
    assert np.array_equal(im[0], g[f"t{steps}_indx_m"])
    d = eng.get_diag()[0]
    assert abs(d[0] + d[1] * N / L - g["H"][steps]) < 1e-11 * g["H"][steps]
    assert abs(d[1] * N / L - g["PE"][steps]) < 1e-10 * g["PE"][steps]
    rho, k = eng.get_density_fixed()
    assert sum(int(r) for r in rho.ravel()) == N * (1 << k)
    assert eng.error_flags() == 0


def test_tsc_through_pic_class(golden):
    from pic_b200 import PIC
    from pic_b200.dist import BumpOnTail
    g = golden("bump_vb3_tsc")
    np.random.seed(42)
    dist = BumpOnTail(a=0.2, v0=3.0, sigma=1.0, n_samples=5000, L=50.0)
    sim = PIC(N=5000, N_mesh=250, n0=1.0, L=50.0, dt=0.1, tmin=0.0, tmax=50.0, gamma=5.0, A=0.1, n_mode=2,
              interpol="TSC", init_dist=dist)
    for t in range(1, 6):
        sim.update_state(None)
        assert abs(sim.get_energy() - g["H"][t]) < 1e-12 * g["H"][t]
    assert sim.indx_m.shape == (5000, 1) and sim.weight_m.shape == (5000, 1)
    assert np.array_equal(sim.indx_l, np.mod(sim.indx_m - 1, 250))


def test_zero_copy_views_for_the_policy():
    """`views()` aliases the device buffers through __cuda_array_interface__: torch sees the particle arrays in
    `get_state()` order without a copy (obs layout [env][x(0..N-1)], [env][v(0..N-1)])."""
    import torch
    from pic_b200 import BatchedPIC
    from pic_b200.dist import TwoStream
    B, N = 3, 5000
    bp = BatchedPIC(B, N=N, N_mesh=250, L=50.0, dt=0.05, max_mode=3)
    bp.reset_from_sampler(lambda: TwoStream(v0=3.0, sigma=1.0, n_samples=N, L=50.0), seed=1)
    bp.step(np.zeros((B, 6)), n_steps=2)
    vw = bp.views()
    xt = torch.as_tensor(vw["x"], device="cuda")
    vt = torch.as_tensor(vw["v"], device="cuda")
    assert xt.shape == (B, N) and xt.dtype == torch.float64 and xt.stride(0) == vw["ld"]
    st = bp.get_state()
    assert np.array_equal(xt.cpu().numpy(), st[:, :N]) and np.array_equal(vt.cpu().numpy(), st[:, N:])
    ptr0 = xt.data_ptr()
    bp.step(np.zeros((B, 6)), n_steps=1)
    assert torch.as_tensor(bp.views()["x"], device="cuda").data_ptr() == ptr0          # same buffer, updated in place
    assert np.array_equal(xt.cpu().numpy(), bp.get_state()[:, :N])
    d = torch.as_tensor(vw["diag"], device="cuda")
    assert d.shape == (B, 6) and float(d[0, 1]) == bp.engine.get_diag()[0, 1]


def test_streaming_multi_env_and_far_positions():
    """Streaming mode with several envs per handle (blockIdx.y = env) equals each env alone, including particles that
    start absurdly far outside the box (|x| up to 1e12 L: the general fmod path) -- compared with the oracle."""
    B, N, M, L, dt = 3, 20011, 512, 50.0, 0.05
    rng = np.random.RandomState(21)
    x = rng.uniform(0, L, (B, N)); v = rng.normal(size=(B, N))
    x[0, :50] = rng.uniform(-1e3, 1e3, 50) * L
    x[1, :50] = rng.uniform(-1e12, 1e12, 50) * L
    x[2, 7] = 2.0 ** 40 * (L / M) + 0.013
    ext = 0.1 * np.cos(2 * np.pi * np.arange(M) / M)
    eng = _engine(N, M, L, dt, n_envs=B, mode="streaming")
    eng.set_state(x, v)
    eng.step_mesh(np.tile(ext, (B, 1)), 2)
    xg, vg = eng.get_state()
    ilg, *_ = eng.get_cells(False, False)
    p = O.PicParams(N=N, N_mesh=M, n0=1.0, L=L, dt=dt)
    for b in range(B):
        xo, vo = O.wrap(x[b], L), v[b]          # PIC.initialize wraps the positions in place first (util.py:51)
        for _ in range(2):
            o = O.step(xo, vo, p, ext)
            xo, vo = o["x"], o["v"]
        assert np.abs(xg[b] - xo).max() < 1e-11 and np.abs(vg[b] - vo).max() < 1e-11
        assert np.array_equal(ilg[b], o["indx_l"])
    one = _engine(N, M, L, dt, mode="streaming")
    one.set_state(x[1:2], v[1:2])
    one.step_mesh(ext[None], 2)
    x1, v1 = one.get_state()
    assert np.array_equal(x1[0], xg[1]) and np.array_equal(v1[0], vg[1])
    assert eng.error_flags() == 0


def test_error_behaviour():
    """Errors are loud and typed: bad configs, calls out of order, unsupported combinations, and the sticky device
    flags for positions the reference itself cannot handle (NaN -> np.bincount raises there)."""
    import pic_b200
    from pic_b200 import Engine, PicError
    with pytest.raises(PicError) as e:
        Engine(0, 16, 10.0, 0.1)
    assert e.value.code == -1
    with pytest.raises(PicError) as e:
        Engine(1000, 16, 10.0, 0.1, precision="f32", interpol="TSC")
    assert e.value.code == -8
    with pytest.raises(PicError) as e:
        Engine(5_000_000, 16, 10.0, 0.01, mode="resident")        # does not fit the shared memory of even an 8-CTA cluster
    assert e.value.code == -8
    with pytest.raises(PicError) as e:
        Engine(1000, 2048, 10.0, 0.1, mode="resident")            # resident field solve is the small-mesh instance
    assert e.value.code == -8 and "n_mesh" in str(e.value)
    big = Engine(50_000, 16, 10.0, 0.1, mode="resident")          # one CTA is too small: spread over a cluster
    assert big.launch_info()["per_thread"] == 4
    with pytest.raises(PicError) as e:
        Engine(100_000, 20000, 10.0, 0.1, mode="streaming")       # mesh tables exceed shared memory
    assert e.value.code == -8 and "n_mesh too large" in str(e.value)
    eng = Engine(1000, 32, 10.0, 0.05, max_mode=2)
    with pytest.raises(PicError) as e:
        eng.step_mesh(None, 1)                                     # no state yet
    assert e.value.code == -5
    rng = np.random.RandomState(0)
    x = rng.uniform(0, 10, 1000); v = rng.normal(size=1000)
    eng.set_state(x[None], v[None])
    with pytest.raises(PicError) as e:
        eng.step_coeffs(np.zeros((1, 1, 4)), 1)                    # actuator basis not uploaded
    assert e.value.code == -5
    with pytest.raises(ValueError):
        eng.step_mesh(np.zeros((1, 31)), 1)                        # wrong mesh length
    assert eng.error_flags() == 0
    x[3] = np.nan
    x[4] = np.inf
    eng.set_state(x[None], v[None])
    eng.step_mesh(None, 2)
    assert eng.error_flags() & 2                                   # non-finite position flagged, no crash
    eng.clear_error_flags()
    x[3] = 1.0; x[4] = 2.0
    eng.set_state(x[None], v[None])
    eng.step_mesh(None, 2)
    assert eng.error_flags() == 0
    rho, k = eng.get_density_fixed()
    assert sum(int(r) for r in rho.ravel()) == 1000 * (1 << k)


@pytest.mark.parametrize("mode", ["resident", "streaming"])
def test_extreme_clustering_and_cell_sorted_order(mode):
    """All particles of a warp in one cell exercises the warp-aggregated deposit (REDUX + one lane's atomics); a
    cell-sorted env and an env with every particle in two cells must still match the oracle bit for bit in the index
    and to 1e-12 in the state."""
    N, M, L, dt = 4096, 128, 50.0, 0.02
    rng = np.random.RandomState(4)
    dx = L / M
    cases = {"two_cells": np.where(rng.uniform(size=N) < 0.5, 17.3 * dx, 90.9 * dx) + rng.uniform(0, 0.05 * dx, N),
             "sorted": np.sort(rng.uniform(0, L, N))}
    for name, x in cases.items():
        v = 0.3 * rng.normal(size=N)
        p = O.PicParams(N=N, N_mesh=M, n0=1.0, L=L, dt=dt)
        eng = _engine(N, M, L, dt, mode=mode)
        eng.set_state(x[None], v[None])
        xo, vo = x, v
        for _ in range(3):
            o = O.step(xo, vo, p, None)
            xo, vo = o["x"], o["v"]
        eng.step_mesh(None, 3)
        xg, vg = eng.get_state()
        scale = max(1.0, np.abs(vo).max())
        assert np.abs(xg[0] - xo).max() < 1e-11 * scale and np.abs(vg[0] - vo).max() < 1e-11 * scale, name
        il, *_ = eng.get_cells(False, False)
        assert np.array_equal(il[0], o["indx_l"]), name
        rho, k = eng.get_density_fixed()
        assert sum(int(r) for r in rho.ravel()) == N * (1 << k)
        assert eng.error_flags() == 0


@pytest.mark.parametrize("mode", ["resident", "streaming"])
def test_exact_weights_option(golden, mode):
    """exact_weights=True divides by dx exactly as interpolate.py:11-12 does: the CIC weights are then the
    reference's bits, not just within an ulp."""
    g = golden("deposit_edges")
    L, M = 50.0, 250
    key = f"L{L:g}_M{M}"
    x = g[key + "_x"]
    N = x.shape[0]
    eng = _engine(N, M, L, 0.01, mode=mode, exact_weights=True)
    eng.set_state(x[None], np.zeros((1, N)))
    il, wl, wr, _ = eng.get_cells()
    assert np.array_equal(il[0], g[key + "_cic_il"])
    assert np.array_equal(wl[0], g[key + "_cic_wl"]) and np.array_equal(wr[0], g[key + "_cic_wr"])
    gb = golden("bump_vb3")
    e2 = _engine(5000, 250, 50.0, 0.1, mode=mode, exact_weights=True)
    e2.set_state(gb["t0_x"][None], gb["t0_v"][None])
    e2.step_mesh(None, 10)
    xg, vg = e2.get_state()
    assert np.abs(xg[0] - gb["t10_x"]).max() < 1e-12 and np.abs(vg[0] - gb["t10_v"]).max() < 1e-12


@pytest.mark.parametrize("mode", ["resident", "streaming"])
@pytest.mark.parametrize("N,M", [(1, 2), (3, 5), (33, 7), (257, 31), (1025, 64)])
def test_tiny_and_ragged_configs(mode, N, M):
    """Degenerate sizes (one particle, two cells, sizes that are not multiples of any vector / warp / tile width)
    against the oracle, with an external field."""
    L, dt = 10.0, 0.05
    rng = np.random.RandomState(N * 131 + M)
    x = rng.uniform(-0.5 * L, 1.5 * L, N); v = rng.normal(size=N)
    ext = 0.2 * np.sin(2 * np.pi * (np.arange(M) + 0.5) / M)
    dtc = O.clip_dt(dt, N, L)
    p = O.PicParams(N=N, N_mesh=M, n0=1.0, L=L, dt=dtc)
    eng = _engine(N, M, L, dtc, mode=mode)
    eng.set_state(x[None], v[None])
    xo, vo = O.wrap(x, L), v
    for _ in range(4):
        o = O.step(xo, vo, p, ext)
        xo, vo = o["x"], o["v"]
    eng.step_mesh(ext[None], 4)
    xg, vg = eng.get_state()
    assert np.abs(xg[0] - xo).max() < 1e-11 and np.abs(vg[0] - vo).max() < 1e-11
    il, *_ = eng.get_cells(False, False)
    assert np.array_equal(il[0], o["indx_l"])
    n, E = eng.get_fields()
    assert np.abs(E[0] - o["E_mesh"]).max() < 1e-11 * max(1.0, np.abs(o["E_mesh"]).max())
    assert eng.error_flags() == 0


def test_tsc_batched_reward_deviation_is_flagged():
    """The reference's Reward re-deposits the state with CIC even for a TSC env (src/control/objective.py:24); the device
    reward of a TSC BatchedPIC uses the env's own TSC field energy.  The constructor says so (a warning), and the number
    returned is exactly the documented one: max(1 - PE_tsc(pre-step state), 0) * alpha + max(1 - ie / r_ie_n, 0) * beta."""
    from pic_b200 import BatchedPIC
    B, N = 3, 4000
    with pytest.warns(UserWarning, match="objective.py:24"):
        bp = BatchedPIC(B, N=N, N_mesh=250, L=50.0, dt=0.05, max_mode=3, interpol="TSC", alpha=0.7, beta=0.3)
    rng = np.random.RandomState(4)
    x = rng.uniform(0, 50.0, (B, N)); v = rng.normal(0, 1, (B, N)) + 3.0 * (rng.uniform(size=(B, N)) < 0.2)
    bp.set_state(x, v)
    pe_pre = np.concatenate([e.get_diag()[:, 1] for e in bp.engines])
    a = rng.uniform(-1, 1, (B, 6))
    out = bp.step(a, 1)
    ie = out["input_energy"][0]
    want = np.maximum(1 - pe_pre / bp.r_pe_n, 0) * 0.7 + np.maximum(1 - ie / bp.r_ie_n, 0) * 0.3
    assert np.allclose(out["reward"][0], want, rtol=1e-13, atol=0)
