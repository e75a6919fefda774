"""Goldens of the reference's own RUNNER SCRIPTS, run unmodified on the reference's CPU `PIC` (BASELINE configs 1-3).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_runner_golden.py

  runner_wo_oc_bump.npz        run_wo_oc.py --simcase bump-on-tail            (config 1; 500 steps)
  runner_wo_oc_twostream.npz   run_wo_oc.py  (default simcase: two-stream)   (config 2)
  runner_wo_oc_bump_vb5.npz    run_wo_oc.py --simcase bump-on-tail --vb 5.0  (the published 0.00557 growth rate)
  runner_ddpg_bump.npz         run_ddpg.py --simcase bump-on-tail            (config 3; seeded random-init Actor saved
                                                                              as ddpg_best.pt, run_ddpg.py:263)

Nothing from this repository takes part in the computation: the scripts are executed with `runpy` from the
reference tree with the reference's own `src.env.pic`; the only additions are a do-nothing matplotlib (absent from
this image; `src/plot.py:2` imports it) and, for run_ddpg.py, the checkpoint file the script loads.  Stored per
runner: the script's own `E`, `PE` arrays, cost lists, the coefficient trajectory and the last snapshot column.
`tests/test_runners.py` replays the same scripts against `pic_b200.PIC` on the B200.
"""
import os
import runpy
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

REF = os.environ.get("PIC_REFERENCE", "/root/reference")
ACTOR_SEED = 0


def write_actor_checkpoint(ref, path, seed=ACTOR_SEED):
    """The file run_ddpg.py:263 loads: a randomly initialised Actor(2N=10000, 64, 6, -1.25, 1.25) from a fixed seed
    (torch's CPU generator is platform independent).  The reference's own class builds it."""
    import torch
    sys.path.insert(0, ref)
    try:
        from src.control.rl.ddpg import Actor
    finally:
        sys.path.remove(ref)
    torch.manual_seed(seed)
    net = Actor(10000, 64, 6, output_min=-1.25, output_max=1.25)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save(net.state_dict(), path)
    return path


def run_reference_script(ref, script, argv, workdir):
    """Run a reference runner unmodified with the reference's own modules.  Returns the script's globals."""
    from pic_b200.run import install_matplotlib_stub          # the stub only; no PIC injection here
    install_matplotlib_stub()
    saved_argv, saved_path, cwd = sys.argv[:], sys.path[:], os.getcwd()
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    try:
        os.chdir(workdir)
        sys.path.insert(0, ref)
        sys.argv = [os.path.join(ref, script)] + list(argv)
        return runpy.run_path(os.path.join(ref, script), run_name="__main__")
    finally:
        sys.argv, sys.path[:] = saved_argv, saved_path
        os.chdir(cwd)
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]


def extract(g, ddpg=False):
    out = {"E": np.asarray(g["E"], dtype=np.float64), "PE": np.asarray(g["PE"], dtype=np.float64),
           "cost_kl": np.asarray(g["cost_kl_list"], dtype=np.float64),
           "cost_ee": np.asarray(g["cost_ee_list"], dtype=np.float64),
           "x_last": np.asarray(g["snapshot"])[: g["snapshot"].shape[0] // 2, -1].copy(),
           "v_last": np.asarray(g["snapshot"])[g["snapshot"].shape[0] // 2:, -1].copy(),
           "x_first": np.asarray(g["snapshot"])[: g["snapshot"].shape[0] // 2, 0].copy(),
           "dt": np.float64(g["sim"].dt)}
    if ddpg:
        out["cost_ie"] = np.asarray(g["cost_ie_list"], dtype=np.float64)
        out["coeff_cos"] = np.asarray(g["coeff_cos"], dtype=np.float64)      # (3, Nt)
        out["coeff_sin"] = np.asarray(g["coeff_sin"], dtype=np.float64)
    return out


def growth_rate(PE_runner, N, L, dt, t_max=50.0):
    """The notebooks' number (analysis/*.ipynb:51-53 -> src/interpret/landau.py:44): half the slope of a least-squares
    line through log(sum E_mesh^2 dx) over time.  PE_runner = 0.5 sum E^2 dx N / L (util.py:129-130)."""
    e2 = np.asarray(PE_runner) * 2.0 * L / N
    t = np.linspace(0.0, t_max, len(e2))
    slope = np.polyfit(t, np.log(e2), 1)[0]
    return 0.5 * slope


if __name__ == "__main__":
    cases = [
        ("runner_wo_oc_bump", "run_wo_oc.py", ["--simcase", "bump-on-tail"], False),
        ("runner_wo_oc_twostream", "run_wo_oc.py", [], False),
        ("runner_wo_oc_bump_vb5", "run_wo_oc.py", ["--simcase", "bump-on-tail", "--vb", "5.0"], False),
        ("runner_ddpg_bump", "run_ddpg.py", ["--simcase", "bump-on-tail"], True),
    ]
    only = os.environ.get("GOLDEN_ONLY", "")
    for name, script, argv, ddpg in cases:
        if only and only not in name:
            continue
        with tempfile.TemporaryDirectory() as tmp:
            argv = argv + ["--save_plot", os.path.join(tmp, "plots"), "--save_file", os.path.join(tmp, "data")]
            if ddpg:
                write_actor_checkpoint(REF, os.path.join(tmp, "data", "bump-on-tail", "ddpg-control", "ddpg_best.pt"))
            g = run_reference_script(REF, script, argv, tmp)
        out = extract(g, ddpg)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "E_end %.12g PE_end %.12g growth %.6f" % (out["E"][-1], out["PE"][-1],
                                                               growth_rate(out["PE"], 5000, 50.0, 0.1)))
