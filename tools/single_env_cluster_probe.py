"""Latency of ONE small env per step vs CTAs per env (cluster):  python tools/single_env_cluster_probe.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pic_b200
for (N, M) in ((5000, 250), (10000, 500), (2000, 128)):
    rng = np.random.RandomState(0)
    x = rng.uniform(0, 50.0, (1, N)); v = rng.normal(size=(1, N))
    ref = None
    for shape in ("1024x1", "512x1", "512x2", "1024x2", "256x4", "512x4", "1024x4", "256x8", "512x8", "1024x8"):
        th, cl = (int(t) for t in shape.split("x"))
        eng = pic_b200.Engine(N, M, 50.0, 0.1, n_envs=1, mode="resident")
        try:
            eng.set_tuning(th, cl, -1)
        except Exception as e:
            print("N=%d %-7s skipped %s" % (N, shape, str(e)[:60])); eng.close(); continue
        eng.set_state(x, v)
        eng.step_mesh(None, 200); eng.sync()
        xs, _ = eng.get_state()
        if ref is None: ref = xs
        t0 = time.perf_counter(); eng.step_mesh(None, 3000); eng.sync()
        us = (time.perf_counter() - t0) / 3000 * 1e6
        print("N=%6d M=%4d %-7s %7.2f us/step  identical=%s" % (N, M, shape, us, bool(np.array_equal(xs, ref))), flush=True)
        eng.close()
