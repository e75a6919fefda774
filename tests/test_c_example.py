"""The C ABI is usable from plain C: examples/step_from_c.c compiles with gcc against include/pic_b200.h alone,
fails loudly without a device (no CPU fallback) and conserves energy on a B200."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "optimal-control-1d-electrostatic-plasma_b200", "lib")


def _build(tmp_path):
    import pic_b200
    pic_b200._lib.load()                                    # builds the library if it is missing
    exe = str(tmp_path / "step_from_c")
    cmd = ["gcc", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "step_from_c.c"), "-o", exe, "-L" + LIBDIR, "-lpic_b200",
           "-Wl,-rpath," + LIBDIR, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_c_example_compiles_and_refuses_to_run_without_a_device(tmp_path):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_c_example_runs_on_the_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "relative energy drift" in r.stdout and "error flags 0" in r.stdout
