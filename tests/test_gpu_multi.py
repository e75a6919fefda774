"""Particle-sharded mode on >= 2 GPUs of one box: bit-identical to the single-GPU run (integer density sum)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("collective", ["fused", "nccl", "torch", "auto"])
def test_sharded_equals_single_gpu(collective, tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = tmp_path / "res.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "_multi_worker.py"), collective, str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.load(open(out))
    assert res["rho_equal"] and res["x_equal"] and res["v_equal"] and res["pe_equal"], res
    assert res["ke_rel"] < 1e-14 and res["sumv_abs"] < 1e-8, res
    assert res["sampler_shard_equal"] and res["sampler_rho_equal"], res
    assert res["flags"] == 0, res
