"""`E_field` -- the reference's actuator (src/control/actuator.py:4-63): Fourier coefficients -> external field on
the mesh.  The tables are tiny and built once on the host exactly as the reference builds them (same linspace nodes,
same cos/sin calls) so that they can be uploaded to the device bit-for-bit; the per-step evaluation
basis_cos @ a + basis_sin @ b runs inside the step kernels when `PIC.update_state_coeffs` / `BatchedPIC.step` are
used, and `compute_E` below serves callers that want the (N_mesh, 1) vector itself."""
import numpy as np


class E_field:
    def __init__(self, L: float, N_mesh: int, max_mode: int):
        self.L = L
        self.N_mesh = N_mesh
        self.dx = L / N_mesh
        self.max_mode = max_mode
        self.reinit()

    def reinit(self):
        self.xm = np.linspace(0, self.L, self.N_mesh)              # endpoint included (actuator.py:13)
        self.coeff_cos = np.zeros((self.max_mode, 1))
        self.coeff_sin = np.zeros((self.max_mode, 1))
        self.k = np.array([2 * np.pi / self.L * n for n in range(1, self.max_mode + 1)])
        self.basis_cos = np.concatenate([np.cos(k * self.xm).reshape(-1, 1) for k in self.k], axis=1)
        self.basis_sin = np.concatenate([np.sin(k * self.xm).reshape(-1, 1) for k in self.k], axis=1)

    def update_params(self, **kwargs):
        for key in kwargs.keys():
            if hasattr(self, key) is True and kwargs[key] is not None:
                setattr(self, key, kwargs[key])

    def update_E(self, coeff_cos=None, coeff_sin=None):
        if coeff_cos is not None:
            self.coeff_cos = np.asarray(coeff_cos).copy().reshape(-1, 1)
        if coeff_sin is not None:
            self.coeff_sin = np.asarray(coeff_sin).copy().reshape(-1, 1)

    def compute_E(self, coeff_cos=None, coeff_sin=None):
        a = self.coeff_cos if coeff_cos is None else np.asarray(coeff_cos)
        b = self.coeff_sin if coeff_sin is None else np.asarray(coeff_sin)
        return self.basis_cos @ a.reshape(-1, 1) + self.basis_sin @ b.reshape(-1, 1)
