"""The reference's own runner scripts, UNCHANGED, on the B200 env (SURVEY §8 row h, BASELINE configs 1-3, T10).

`run_wo_oc.py` (both simcases, full 500 steps) and `run_ddpg.py --simcase bump-on-tail` (seeded random-init Actor saved
as ddpg_best.pt, run_ddpg.py:263) are executed with `runpy` from the unmodified reference tree (/root/reference here,
the staged copy under baseline/_ref/ on the GPU box) with `pic_b200.PIC` injected as `src.env.pic` by the launcher
(pic_b200/run.py).  Everything else the scripts import -- dist.py, reward.py, actuator.py, the DDPG Actor, the
plotting helpers -- is the reference's own code.

Pinned against
  * tests/golden/runner_*.npz: the same scripts run on the reference's CPU PIC (tests/golden/make_runner_golden.py);
  * the notebooks' published growth rates 0.02135 / 0.00557 (analysis/*.ipynb:51-53), from the runner's own output;
  * for run_ddpg.py additionally a LIVE run of the same script on the reference's CPU PIC on the same machine, because
    the Actor is float32 torch code whose last bits may depend on the host CPU (the committed golden comes from the
    build container's CPU).

Tolerances (float64; atomic-order noise grows chaotically): E, PE lists 1e-9 relative over all 500
steps, the reference Reward's electric-energy cost 1e-9, final particle positions 1e-7 absolute, KL cost 2e-3
relative (a 2-D histogram count: a particle 1e-9 from a bin edge moves one count of 5000).
"""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, reference_dir

sys.path.insert(0, GOLDEN)
import make_runner_golden as MG  # noqa: E402  (helpers only: checkpoint writer, growth-rate fit, reference runner)

REF = reference_dir()
needs_ref = pytest.mark.skipif(REF is None, reason="reference tree not present (run tools/stage_reference.py)")


def run_on_b200(script, argv, tmp_path, ddpg=False):
    from pic_b200 import run as launcher
    argv = list(argv) + ["--save_plot", str(tmp_path / "plots"), "--save_file", str(tmp_path / "data")]
    if ddpg:
        MG.write_actor_checkpoint(REF, str(tmp_path / "data" / "bump-on-tail" / "ddpg-control" / "ddpg_best.pt"))
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        return launcher.run_script(os.path.join(REF, script), argv, reference_dir=REF)
    finally:
        os.chdir(cwd)


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.abs(np.asarray(b))))


WO_OC = [("runner_wo_oc_bump", ["--simcase", "bump-on-tail"], -0.001295),
         ("runner_wo_oc_twostream", [], 0.021354),                       # published 0.02135
         ("runner_wo_oc_bump_vb5", ["--simcase", "bump-on-tail", "--vb", "5.0"], 0.005568)]   # published 0.00557


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("name,argv,rate", WO_OC, ids=[c[0] for c in WO_OC])
def test_run_wo_oc_unchanged_on_b200(name, argv, rate, tmp_path, golden):
    import pic_b200
    g = run_on_b200("run_wo_oc.py", argv, tmp_path)
    gold = golden(name)
    assert isinstance(g["sim"], pic_b200.PIC) and type(g["sim"]).__module__.startswith("pic_b200")
    assert len(g["E_list"]) == 500 and g["snapshot"].shape == (10000, 500)
    out = MG.extract(g)
    assert np.array_equal(out["x_first"] > 25.0, gold["x_first"] > 25.0)      # same particles (host RNG stream)
    assert rel(out["E"], gold["E"]) < 1e-9
    assert rel(out["PE"], gold["PE"]) < 1e-9
    assert rel(out["cost_ee"], gold["cost_ee"]) < 1e-9         # the reference's own Reward on our get_state()
    assert rel(out["cost_kl"], gold["cost_kl"]) < 2e-3
    assert np.abs(out["x_last"] - gold["x_last"]).max() < 1e-7 and np.abs(out["v_last"] - gold["v_last"]).max() < 1e-7
    got = MG.growth_rate(out["PE"], 5000, 50.0, 0.1)
    assert abs(got - rate) < 1e-6, got
    assert abs(got - MG.growth_rate(gold["PE"], 5000, 50.0, 0.1)) < 1e-9
    assert g["sim"].engine.error_flags() == 0


@pytest.mark.gpu
@needs_ref
def test_run_ddpg_unchanged_on_b200(tmp_path, golden):
    """Config 3: the evaluation episode of run_ddpg.py:260-313 -- sim.reinit(), 500 x (get_state -> Actor.get_action ->
    E_field.update_E / compute_E -> update_state(E_external) -> energies, costs)."""
    import pic_b200
    (tmp_path / "gpu").mkdir()
    (tmp_path / "cpu").mkdir()
    g = run_on_b200("run_ddpg.py", ["--simcase", "bump-on-tail"], tmp_path / "gpu", ddpg=True)
    assert isinstance(g["sim"], pic_b200.PIC)
    out = MG.extract(g, ddpg=True)
    assert out["coeff_cos"].shape == (3, 500) and out["E"].shape == (500,)
    assert np.all(np.abs(out["coeff_cos"]) <= 1.25) and np.all(np.abs(out["coeff_sin"]) <= 1.25)

    # (1) PIC parity proper: the reference's CPU PIC, built by the same constructor sequence (seed 42 at import, dist
    # ctor, PIC(), reinit(): run_ddpg.py:139-260), driven OPEN LOOP with the coefficient trajectory the GPU episode
    # produced.  Same particles, same actions => the two envs must agree to float64 step tolerance over 500 steps.
    ref = replay_on_reference_pic(out["coeff_cos"], out["coeff_sin"])
    assert rel(out["E"], ref["E"]) < 1e-9
    assert rel(out["PE"], ref["PE"]) < 1e-8
    assert rel(out["cost_ee"], ref["cost_ee"]) < 1e-8
    assert rel(out["cost_ie"], ref["cost_ie"]) < 1e-6           # the script sums float32 coefficients (reward.py:52-54)
    assert rel(out["cost_kl"], ref["cost_kl"]) < 2e-3
    # (the controlled plasma is driven hard -- PE ends 60x above the uncontrolled run -- and single trajectories
    #  diverge faster than the energies: 5e-7 measured after 500 steps)
    assert np.abs(out["x_last"] - ref["x_last"]).max() < 1e-5 and np.abs(out["v_last"] - ref["v_last"]).max() < 1e-5

    # (2) closed loop: the same unchanged script on the reference's CPU PIC on this machine, and the committed golden
    # (build container).  The Actor is float32: a 1e-13 difference in the state flips float32 roundings of its inputs,
    # its output moves by float32 ulps (~1e-6), and the loop feeds that back -- the reference does the same to itself
    # when np.bincount's summation order changes -- so these comparisons carry the policy's tolerance, not the PIC's.
    argv = ["--simcase", "bump-on-tail", "--save_plot", str(tmp_path / "cpu" / "plots"),
            "--save_file", str(tmp_path / "cpu" / "data")]
    MG.write_actor_checkpoint(REF, str(tmp_path / "cpu" / "data" / "bump-on-tail" / "ddpg-control" / "ddpg_best.pt"))
    live = MG.extract(MG.run_reference_script(REF, "run_ddpg.py", argv, str(tmp_path / "cpu")), ddpg=True)
    gold = golden("runner_ddpg_bump")
    for other in (live, gold):
        assert np.array_equal(out["coeff_cos"][:, 0], other["coeff_cos"][:, 0])       # first action: identical state
        assert np.abs(out["coeff_cos"] - other["coeff_cos"]).max() < 1e-4
        assert np.abs(out["coeff_sin"] - other["coeff_sin"]).max() < 1e-4
        assert rel(out["E"], other["E"]) < 1e-5
        assert rel(out["PE"], other["PE"]) < 1e-4
    assert g["sim"].engine.error_flags() == 0


def replay_on_reference_pic(coeff_cos, coeff_sin):
    """run_ddpg.py:139-160,260,276-313 with the reference's own classes and a GIVEN coefficient trajectory."""
    import importlib
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    try:
        pic_mod = importlib.import_module("src.env.pic")            # runs np.random.seed(42) (pic.py:12)
        dist_mod = importlib.import_module("src.env.dist")
        act_mod = importlib.import_module("src.control.actuator")
        rew_mod = importlib.import_module("src.control.rl.reward")
        dist = dist_mod.BumpOnTail(a=0.2, v0=3.0, sigma=1.0, n_samples=5000, L=50)
        sim = pic_mod.PIC(N=5000, N_mesh=250, n0=1.0, L=50, dt=0.1, tmin=0, tmax=50, gamma=5.0, A=0.1, n_mode=2,
                          interpol="CIC", init_dist=dist)
        actuator = act_mod.E_field(50, 250, 3)
        sim.reinit()
        reward = rew_mod.Reward(sim.init_dist.get_init_state(), 250, 50, -25.0, 25.0, 1.0, 0.1, 0.1, 6)
        out = {k: [] for k in ("E", "PE", "cost_kl", "cost_ee", "cost_ie")}
        for t in range(coeff_cos.shape[1]):
            coeffs = np.concatenate([coeff_cos[:, t], coeff_sin[:, t]])
            actuator.update_E(coeffs[:3], coeffs[3:])
            sim.update_state(actuator.compute_E())
            out["E"].append(sim.get_energy()); out["PE"].append(sim.get_electric_energy())
            out["cost_kl"].append(reward.compute_kl_divergence(sim.get_state()))
            out["cost_ee"].append(reward.compute_electric_energy(sim.get_state()))
            out["cost_ie"].append(reward.compute_input_energy(coeffs))
        out = {k: np.asarray(v, dtype=np.float64) for k, v in out.items()}
        out["x_last"], out["v_last"] = sim.x[:, 0].copy(), sim.v[:, 0].copy()
        return out
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]


@needs_ref
def test_runner_goldens_are_what_the_reference_produces(golden):
    """CPU: the committed run_wo_oc.py golden is reproduced by the reference here (first 20 steps; guards against a
    stale fixture) and agrees with the step-level golden generated by make_golden.py."""
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        g = MG.run_reference_script(REF, "run_wo_oc.py", ["--simcase", "bump-on-tail", "--t_max", "2",
                                                          "--save_plot", tmp + "/p", "--save_file", tmp + "/d"], tmp)
    gold = golden("runner_wo_oc_bump")
    assert np.array_equal(np.asarray(g["E"]), gold["E"][:20])
    assert np.array_equal(np.asarray(g["PE"]), gold["PE"][:20])
    assert np.array_equal(gold["E"], golden("bump_vb3")["H"][1:])
    assert abs(MG.growth_rate(golden("runner_wo_oc_twostream")["PE"], 5000, 50.0, 0.1) - 0.02135) < 5e-6
    assert abs(MG.growth_rate(golden("runner_wo_oc_bump_vb5")["PE"], 5000, 50.0, 0.1) - 0.00557) < 5e-6
