"""Poisson2D (reference: src/fluids2d/elliptic.py:71-99): same constructor and
``solve(b, x)`` contract, but the solve is the device multigrid / PCG of
``csrc/mg.cu`` instead of a SuperLU factorisation."""
import numpy as np


class Poisson2D:
    def __init__(self, mesh, location, maindiag=0):
        assert location in ["c", "v"]
        self.mesh = mesh
        self.location = location
        self.maindiag = maindiag
        if maindiag == 0:
            self._which = location
        else:
            p = mesh.param
            expected = mesh.area * p.f0 ** 2 / (p.g * p.H)
            if location != "v" or abs(maindiag - expected) > 1e-14 * abs(expected):
                raise NotImplementedError(
                    "only the vertex Helmholtz operator of meshes.py:42-47 "
                    "(maindiag = area*f0**2/(g*H)) is built on the device")
            self._which = "h"
        self.last = None       # (iterations, relative residual) of the last solve

    @property
    def G(self):
        """index map of the unknowns (elliptic.py:102-111, 198-203)"""
        m = self.mesh
        msk = (m.msk if self.location == "c" else m.mskv) * 1
        if m.param.xperiodic:
            n = m.param.halowidth
            msk[:, :n] = 0
            msk[:, -n:] = 0
        G = np.full(msk.shape, -1, dtype="i")
        G[msk == 1] = np.arange(int(np.sum(msk)))
        return G

    def solve(self, b, x, usingLU=True):
        """A x = b on the fluid points; x is the first guess on entry, masked
        entries are left untouched, then mesh.fill(x).  Returns None."""
        self.last = self.mesh.engine.solve(self._which, b, x)

    def get_rhs(self, config="basic"):
        ny, nx = self.mesh.shape
        b = np.zeros(self.mesh.shape)
        if config == "basic":
            b[2 * ny // 3, nx // 3] = 1
            b[2 * ny // 3, 2 * nx // 3] = -1
        elif config == "mixed":
            b[ny // 3 + ny // 5, nx // 5] = 1
            b[2 * ny // 3 + ny // 5, 2 * nx // 3] = -1
        return b
