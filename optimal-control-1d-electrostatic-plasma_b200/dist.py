"""Initial-condition samplers with the reference's interface and RNG stream (src/env/dist.py:27-194).

These stay on the host on purpose: parity with the reference needs the identical legacy numpy RNG sequence
(`np.random.seed(42)` then three uniform draws of 1000 per accept round).  Accepted points are gathered with array
masks instead of Python lists; the draws, their order and the accept rule are the reference's, so the particles are
bit-identical (tests/test_host_side.py).
"""
import numpy as np


def _maxwellian(v, vb, sigma):
    return 1 / np.sqrt(2 * np.pi) / sigma * np.exp(-0.5 * (v - vb) ** 2 / sigma ** 2)   # dist.py:66-68,147-149


class _Accumulator:
    """Grows (x, v) by accept/reject rounds drawn from the global numpy RNG (dist.py:75-81, 161-168)."""

    def __init__(self, L, batch=1000):
        self.L, self.batch = L, batch
        self.xs, self.vs, self.count = [], [], 0

    def round(self, vb, sigma):
        x = np.random.uniform(0, self.L, size=self.batch)
        v = np.random.uniform(-10, 10, size=self.batch)
        u = np.random.uniform(0, 1.0, size=self.batch)
        keep = u < _maxwellian(v, vb, sigma)
        self.xs.append(x[keep]); self.vs.append(v[keep]); self.count += int(keep.sum())

    def truncate(self, n):
        x = np.concatenate(self.xs)[:n] if self.xs else np.empty(0)
        v = np.concatenate(self.vs)[:n] if self.vs else np.empty(0)
        self.xs, self.vs, self.count = [x], [v], x.shape[0]

    def result(self, n):
        self.truncate(n)
        return self.xs[0], self.vs[0]


class _Dist:
    def initialize(self, n_samples):
        state = self.rejection_sampling(n_samples)
        self.x_init = state[:, 0]
        self.v_init = state[:, 1]

    def reinit(self):
        self.initialize(self.n_samples)

    def get_sample(self):
        return self.x_init.copy(), self.v_init.copy()

    def get_init_state(self):
        return np.concatenate([self.x_init.copy().reshape(-1, 1), self.v_init.copy().reshape(-1, 1)], axis=0)

    def update_params(self, **kwargs):
        for key in kwargs.keys():
            if hasattr(self, key) is True and kwargs[key] is not None:
                setattr(self, key, kwargs[key])


class TwoStream(_Dist):
    """Two counter-streaming Maxwellians at +-v0 (dist.py:27-102)."""

    def __init__(self, v0: float = 4.0, sigma: float = 0.5, n_samples: int = 40000, L: float = 50):
        self.v0, self.sigma, self.L, self.n_samples = v0, sigma, L, n_samples
        self.initialize(n_samples)

    def get_target_prob(self, v, vb):
        return _maxwellian(v, vb, self.sigma)

    def rejection_sampling(self, n_samples: int, batch: int = 1000):
        acc = _Accumulator(self.L, batch)
        half = n_samples // 2
        while acc.count <= half:                 # `<=`: the reference draws one round more than needed (dist.py:75)
            acc.round(self.v0, self.sigma)
        acc.truncate(half)
        while acc.count < n_samples:
            acc.round(-self.v0, self.sigma)
        x, v = acc.result(n_samples)
        out = np.zeros((n_samples, 2))
        out[:, 0] = x
        out[:, 1] = v
        return out


class BumpOnTail(_Dist):
    """Thermal bulk N(0,1) plus a beam N(v0, sigma) with density ratio a (dist.py:104-194)."""

    def __init__(self, a: float = 0.3, v0: float = 4.0, sigma: float = 0.5, n_samples: int = 40000, L: float = 10):
        self.a, self.v0, self.sigma, self.L, self.n_samples = a, v0, sigma, L, n_samples
        self.initialize(n_samples)

    def initialize(self, n_samples):
        super().initialize(n_samples)
        self.high_indx = self.inject_high_electron_indice()

    def get_target_prob(self, v, vb, sigma):
        return _maxwellian(v, vb, sigma)

    def _n_bulk(self, n_samples):
        return int(n_samples * (1 / (1 + self.a)))

    def rejection_sampling(self, n_samples: int, batch: int = 1000):
        acc = _Accumulator(self.L, batch)
        n1 = self._n_bulk(n_samples)
        while acc.count < n1:
            acc.round(0.0, 1.0)
        acc.truncate(n1)
        while acc.count < n_samples:
            acc.round(self.v0, self.sigma)
        x, v = acc.result(n_samples)
        out = np.zeros((n_samples, 2))
        out[:, 0] = x
        out[:, 1] = v
        return out

    def inject_high_electron_indice(self):
        return np.arange(self._n_bulk(self.n_samples), self.n_samples)
