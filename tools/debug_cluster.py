import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_cluster as T
rng = np.random.RandomState(8)
def cmp(tag, a, b):
    d = a["x"] != b["x"]
    print("%-40s x diff %5d  v diff %5d  rho eq %s  maxdx %.2e" % (tag, int(d.sum()), int((a["v"] != b["v"]).sum()),
          np.array_equal(a["rho"], b["rho"]), np.abs(a["x"] - b["x"]).max()))
for (N, M, m, steps) in ((5000, 250, 3, 3), (5000, 250, 3, 1), (40000, 500, 5, 1)):
    B = 2
    x = rng.uniform(0, 50.0, (B, N)); v = rng.normal(size=(B, N)) + 3.0 * (rng.uniform(size=(B, N)) < 0.17)
    for amp in (1.0, 0.0):
        coeffs = amp * rng.uniform(-1, 1, (steps, B, 2 * m))
        st = T._run(N, M, B, None, steps, coeffs, x, v, m=m, mode="streaming")
        if N <= 10000:
            res = T._run(N, M, B, (512, 1), steps, coeffs, x, v, m=m)
            cmp("N=%d steps=%d amp=%g resident vs streaming" % (N, steps, amp), res, st)
        clu = T._run(N, M, B, (1024, 4), steps, coeffs, x, v, m=m)
        cmp("N=%d steps=%d amp=%g cluster4 vs streaming" % (N, steps, amp), clu, st)
