"""Launcher mechanics.  On CPU the injected class is a pass-through subclass of the REFERENCE's own PIC
(only the plumbing is under test here: module injection, seed side effect, matplotlib stub, runpy) and the test is
skipped when /root/reference is absent; on the GPU box a synthetic runner that follows run_wo_oc.py's loop drives
the CUDA PIC through the same injection."""
import os
import sys
import textwrap

import numpy as np
import pytest

from conftest import reference_dir

REF = reference_dir() or "/nonexistent"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "run_wo_oc.py")), reason="reference tree not present")
def test_unchanged_runner_through_launcher_cpu_plumbing(tmp_path, golden, monkeypatch):
    from pic_b200 import run as launcher
    sys.path.insert(0, REF)
    try:
        import importlib
        sys.dont_write_bytecode = True
        ref_pic = importlib.import_module("src.env.pic")
        RefPIC = ref_pic.PIC
    finally:
        sys.path.remove(REF)
    calls = {"update": 0, "init": 0}

    class Passthrough(RefPIC):
        def __init__(self, *a, **k):
            calls["init"] += 1
            super().__init__(*a, **k)

        def update_state(self, E_external=None):
            calls["update"] += 1
            return super().update_state(E_external)

    monkeypatch.chdir(tmp_path)
    g = launcher.run_script(os.path.join(REF, "run_wo_oc.py"),
                            ["--simcase", "bump-on-tail", "--t_max", "2", "--save_plot", str(tmp_path / "plots"),
                             "--save_file", str(tmp_path / "data")], pic_class=Passthrough)
    assert calls["init"] == 1 and calls["update"] == 20
    gold = golden("bump_vb3")
    assert np.array_equal(np.asarray(g["E"]), gold["H"][1:21])        # same RNG stream, same trajectory


SYNTHETIC_RUNNER = textwrap.dedent('''
    import numpy as np
    from src.env.pic import PIC
    from src.env.dist import BumpOnTail
    dist = BumpOnTail(a=0.2, v0=3.0, sigma=1.0, n_samples=5000, L=50.0)
    sim = PIC(N=5000, N_mesh=250, n0=1.0, L=50.0, dt=0.1, tmin=0.0, tmax=50.0, gamma=5.0, A=0.1, n_mode=2,
              interpol="CIC", init_dist=dist)
    E_list, PE_list, pos = [], [], []
    for t in range(10):
        sim.update_state(None)
        E_list.append(sim.get_energy()); PE_list.append(sim.get_electric_energy())
        pos.append(sim.x.copy())
    state = sim.get_state()
''')


@pytest.mark.gpu
def test_launcher_drives_cuda_pic(tmp_path, golden):
    """A runner written like run_wo_oc.py:76-122 (imports `src.env.pic`, `src.env.dist`) gets the CUDA env and
    reproduces the reference trajectory."""
    from pic_b200 import run as launcher
    import pic_b200
    (tmp_path / "src" / "env").mkdir(parents=True)
    # the runner's `src.env.dist` import: a one-line module re-exporting our host samplers (the real runners use the
    # reference's own file, which draws the identical stream)
    (tmp_path / "src" / "env" / "dist.py").write_text("from pic_b200.dist import BumpOnTail, TwoStream\n")
    script = tmp_path / "runner.py"
    script.write_text(SYNTHETIC_RUNNER)
    g = launcher.run_script(str(script), [], reference_dir=str(tmp_path))
    gold = golden("bump_vb3")
    assert isinstance(g["sim"], pic_b200.PIC)
    assert np.max(np.abs(np.array(g["E_list"]) - gold["H"][1:11]) / gold["H"][1:11]) < 1e-12
    assert np.max(np.abs(np.array(g["PE_list"]) - gold["PE"][1:11]) / gold["PE"][1:11]) < 1e-11
    assert np.abs(g["pos"][-1][:, 0] - gold["t10_x"]).max() < 1e-12
    assert g["state"].shape == (10000, 1)
