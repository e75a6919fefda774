"""Device-side sampler (statistical parity with src/env/dist.py) and size-independent invariants at scale."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_device_sampler_statistics():
    from pic_b200 import Engine
    N, M, L = 2_000_000, 1024, 50.0
    eng = Engine(N, M, L, 0.01, mode="streaming")
    eng.sample_state("bump-on-tail", a=0.2, v0=3.0, sigma=1.0, A=0.0, n_mode=2, seed=1)
    x, v = eng.get_state()
    x, v = x[0], v[0]
    n1 = int(N * (1 / 1.2))
    assert x.min() >= 0 and x.max() < L and abs(x.mean() - L / 2) < 0.05
    assert abs(v[:n1].mean()) < 5e-3 and abs(v[:n1].std() - 1.0) < 5e-3          # bulk N(0,1)
    assert abs(v[n1:].mean() - 3.0) < 1e-2 and abs(v[n1:].std() - 1.0) < 1e-2    # beam N(3,1)
    assert np.abs(v).max() <= 10.0
    eng2 = Engine(N, M, L, 0.01, mode="streaming")
    eng2.sample_state("two-stream", v0=3.0, sigma=0.5, A=0.1, n_mode=2, seed=1)
    x2, v2 = eng2.get_state()
    vp = v2[0] / (1 + 0.1 * np.sin(2 * np.pi * 2 * x2[0] / L))                   # undo pic.py:68
    assert abs(vp[:N // 2].mean() - 3.0) < 5e-3 and abs(vp[N // 2:].mean() + 3.0) < 5e-3
    assert abs(vp[:N // 2].std() - 0.5) < 5e-3


def test_invariants_at_scale():
    """No oracle at this size: total deposited charge is exactly N * 2^k, sum(n) dx = n0 L, momentum is conserved
    without control, and the Hamiltonian drifts by < 1e-6 relative over 20 steps."""
    from pic_b200 import Engine
    N, M, L = 50_000_000, 4096, 50.0
    dt = 2 / np.sqrt(N / L)
    eng = Engine(N, M, L, dt, mode="streaming", deposit="split32")
    eng.set_tuning(1024, 1, 0)
    eng.sample_state("bump-on-tail", seed=3)
    d0 = eng.get_diag()[0]
    rho, k = eng.get_density_fixed()
    assert sum(int(r) for r in rho.ravel()) == N * (1 << k)
    n, E = eng.get_fields()
    assert abs(n.sum() * (L / M) - L) < 1e-9
    assert abs(E.sum()) < 1e-9 * np.abs(E).max() * M
    eng.step_mesh(None, 20)
    tr = eng.get_trace(20)[:, 0, :]
    H0 = d0[0] + d0[1] * N / L
    H = tr[:, 0] + tr[:, 1] * N / L
    assert np.max(np.abs(H - H0)) / H0 < 1e-6
    assert np.max(np.abs(tr[:, 2] - d0[2])) < 1e-7 * N ** 0.5
    assert eng.error_flags() == 0
    rho, k = eng.get_density_fixed()
    assert sum(int(r) for r in rho.ravel()) == N * (1 << k)


def test_full_size_1e9_invariants():
    """BASELINE config 5 at its full size (1e9 particles, 4096 cells) through size-independent properties: every
    particle deposits exactly 2^k (integer density sums to N 2^k), sum(n) dx = n0 L, momentum is conserved without
    control and the Hamiltonian drifts by < 1e-7 relative over 5 env steps."""
    import torch
    from pic_b200 import Engine
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs > 60 GB of device memory")
    N, M, L = 1_000_000_000, 4096, 50.0
    dt = 2 / np.sqrt(N / L)
    eng = Engine(N, M, L, dt, mode="streaming")
    eng.sample_state("bump-on-tail", seed=42)
    d0 = eng.get_diag()[0]
    rho, k = eng.get_density_fixed()
    assert k == 41 and sum(int(r) for r in rho.ravel()) == N * (1 << k)
    n, E = eng.get_fields()
    assert abs(n.sum() * (L / M) - L) < 1e-9 and abs(n.mean() - 1.0) < 1e-12
    eng.step_mesh(None, 5)
    tr = eng.get_trace(5)[:, 0, :]
    H0 = d0[0] + d0[1] * N / L
    assert np.max(np.abs(tr[:, 0] + tr[:, 1] * N / L - H0)) / H0 < 1e-7
    assert np.max(np.abs(tr[:, 2] - d0[2])) < 1e-6 * np.sqrt(N)
    rho, k = eng.get_density_fixed()
    assert sum(int(r) for r in rho.ravel()) == N * (1 << k)
    assert eng.error_flags() == 0
    eng.close()


def test_batched_reset_and_observe():
    """Device-side `reset` (PIC.reinit for every env) + zero-copy `observe`: envs get different, reproducible
    samples; an env-sharded batch holds exactly the envs of the unsharded one."""
    from pic_b200 import BatchedPIC
    B, N = 8, 5000
    whole = BatchedPIC(B, N=N, N_mesh=250, L=50.0, dt=0.05, max_mode=3)
    obs = whole.reset(seed=3)
    assert obs["x"].shape == (B, N) and obs["x"].is_cuda
    st = whole.get_state()
    assert not np.array_equal(st[0], st[1])
    assert abs(st[:, N:].mean()) < 1.0 and (st[:, :N] >= 0).all() and (st[:, :N] < 50.0).all()
    again = BatchedPIC(B, N=N, N_mesh=250, L=50.0, dt=0.05, max_mode=3)
    again.reset(seed=3)
    assert np.array_equal(again.get_state(), st)
    half = BatchedPIC(B, N=N, N_mesh=250, L=50.0, dt=0.05, max_mode=3, rank=1, world_size=2)
    half.reset(seed=3)
    assert (half.env_lo, half.env_hi) == (4, 8)
    assert np.array_equal(half.get_state(), st[4:8])          # env e draws the same sample whatever the sharding
    out_w = whole.step(np.zeros((B, 6)), n_steps=3)
    assert out_w["reward"].shape == (3, B) and np.all(out_w["reward"] <= 2.0)
    assert np.allclose(obs["diag"][:, 1].cpu().numpy(), out_w["pe_mesh"][-1])


def test_batched_env_groups_change_nothing_but_the_schedule():
    """BatchedPIC splits its envs into groups on separate streams (so that one group's launches fill the CTA slots the
    other's last wave leaves empty): every env must come out bit-identical to the single-group batch."""
    from pic_b200 import BatchedPIC
    B, N = 70, 2000
    rng = np.random.RandomState(3)
    acts = rng.uniform(-1, 1, size=(5, B, 6))
    outs = []
    for groups in (1, 2, 3):
        bp = BatchedPIC(B, N=N, N_mesh=128, L=50.0, dt=0.05, max_mode=3, groups=groups)
        assert bp.groups == groups and len(bp.engines) == groups
        bp.enable_modes()
        bp.reset("two-stream", seed=9)
        r = bp.step(acts, n_steps=5)
        r2 = bp.step(None, n_steps=2)
        outs.append((bp.get_state(), r["pe_mesh"], r["reward"], r["modes"], r2["ke"]))
        obs = bp.observe()
        if groups == 1:
            assert obs["x"].shape == (B, N) and bp.engine is bp.engines[0]
        else:
            assert [o["x"].shape[0] for o in obs] == [hi - lo for lo, hi in bp.group_bounds()]
            with pytest.raises(AttributeError):
                bp.engine
        bp.close()
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert np.array_equal(a, b)
    assert BatchedPIC(64, N=100, N_mesh=16, dt=0.01).groups == 2 and BatchedPIC(63, N=100, N_mesh=16, dt=0.01).groups == 1
