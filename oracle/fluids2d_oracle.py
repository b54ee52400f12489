"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not product code.

CPU restatement (numpy + the C kernels in ``weno_oracle.c`` + scipy's SuperLU)
of the Fluids2d time-step hot path, used as the checker for the CUDA path.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the
product package ``fluids2d_b200`` never does.

What is restated, and from where (paths relative to /root/reference/):
  * stencil orders, masks, slip coefficient, halo fill
        src/fluids2d/meshes.py:70-104,135-186, src/fluids2d/noslip.py:4-36
  * WENO kernels (C)                    src/fluids2d/weno.py:21-43,75-134,167-197,255-436
  * discrete operators                  src/fluids2d/operators.py:6-211
  * model right-hand sides / diagnostics src/fluids2d/equations.py:9-155
  * Runge-Kutta drivers                 src/fluids2d/integrators.py:82-124,154-174
  * masked 5-point Laplacian + direct solve
        src/fluids2d/elliptic.py:71-87,102-203
    The factorisation itself is third-party: scipy.sparse.linalg.splu
    (SuperLU), un-pinned in the reference's pyproject.toml:24-30; this image
    carries scipy 1.18.1.  The matrix is fully defined in-repo, so the answer
    is pinned mathematically; we call the same scipy routine.

Parity pin: ``tests/test_oracle_vs_golden.py`` compares this module with
field dumps of the live reference (``tests/golden/*.npz`` written by
``tests/golden/make_golden.py``): bit-exact for the stencil path, <=1e-12
relative for anything downstream of the SuperLU solve (BLAS kernels may
differ between hosts).

The mesh set-up here is vectorised (the reference uses pure-Python loops that
take minutes at 2048^2) but produces identical integer arrays and an
identical CSC matrix.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from types import SimpleNamespace

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle_weno.so")
_SRC = os.path.join(_HERE, "weno_oracle.c")

METHODS = {"weno": 0, "upwind": 1, "centered": 2, "cweno": 3}


def build(force=False):
    """Compile the C kernels (gcc, no FMA contraction)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
             "-o", _LIB, _SRC, "-lm"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        P, I8, I64, I = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
        L.oracle_compflux.argtypes = [P, P, P, I8, I64, I64, I, I64, I64]
        L.oracle_vortexforce.argtypes = [P, P, P, I8, I64, I64, I64, I, I, I64, I64]
        L.oracle_innerproduct.argtypes = [P, P, P, I8, I64, I64, I, I64, I64]
        L.oracle_set_order.argtypes = [I8, I64, I64, I8, I]
        D = ctypes.c_double
        L.oracle_weno3z.argtypes = [D] * 3
        L.oracle_weno3z.restype = D
        L.oracle_weno5z.argtypes = [D] * 5
        L.oracle_weno5z.restype = D
        L.oracle_f1.argtypes = [I] + [D] * 3
        L.oracle_f1.restype = D
        L.oracle_f3.argtypes = [I] + [D] * 5
        L.oracle_f3.restype = D
        L.oracle_f5.argtypes = [I] + [D] * 7
        L.oracle_f5.restype = D
        _lib = L
    return _lib


def _p(a):
    assert a.flags.c_contiguous
    return a.ctypes.data


# --------------------------------------------------------------------------
# kernels  (weno.py:412-436 wrappers)
# --------------------------------------------------------------------------
def compflux(flx, U, q, o, s, method):
    assert flx.dtype == U.dtype == q.dtype == np.float64 and o.dtype == np.int8
    lib().oracle_compflux(_p(flx), _p(U), _p(q), _p(o), q.size, s, METHODS[method], 0, q.size)


def vortexforce(du, V, omega, o, s, s2, sign, method):
    assert du.dtype == V.dtype == omega.dtype == np.float64 and o.dtype == np.int8
    lib().oracle_vortexforce(_p(du), _p(V), _p(omega), _p(o), du.size, s, s2, sign,
                             METHODS[method], 0, du.size)


def innerproduct(ke, U, u, o, s, method):
    assert ke.dtype == U.dtype == u.dtype == np.float64 and o.dtype == np.int8
    lib().oracle_innerproduct(_p(ke), _p(U), _p(u), _p(o), ke.size, s, METHODS[method], 0, ke.size)


# --------------------------------------------------------------------------
# parameters / mesh
# --------------------------------------------------------------------------
_DEFAULTS = dict(  # param.py:13-59 (only what the hot path reads)
    model="euler", nx=40, ny=40, Lx=1.0, Ly=1.0, xperiodic=False, yperiodic=False,
    halowidth=3, noslip=None, f0=10.0, beta=0.0, g=1, H=1, dt=0.0, cfl=0.9, dtmax=9e99,
    integrator="rk3", compflux="weno", vortexforce="weno", innerproduct="weno",
    maxorder=6, tracer=None, RAgamma=0.1,
    # NOT in the reference (SURVEY note Y): a truly periodic y direction -- halo rows are images
    # (Mesh.fill copies rows as it copies columns) and the Laplacian wraps in y the way
    # elliptic.py:153-160 wraps it in x.  The reference has no such mode, so this extension is
    # pinned only by symmetry: tests/test_oracle_vs_golden.py runs an x-periodic channel (pinned
    # against the live reference) and its transpose as a y-periodic channel and demands the
    # transposed fields.
    ywrap=False)


def make_param(**kw):
    d = dict(_DEFAULTS)
    for k in kw:
        if k not in d:
            raise KeyError(k)
    d.update(kw)
    return SimpleNamespace(**d)


def set_order(msk, shift, maxorder):
    """meshes.py:146-186 (C loop, identical integer result)."""
    m = np.ascontiguousarray(msk, dtype=np.int8)
    o = np.zeros(m.shape, dtype=np.int8)
    lib().oracle_set_order(_p(m), m.size, shift, _p(o), maxorder)
    return o


def slipcoef(param, msk):
    """noslip.py:4-36"""
    coef = np.zeros(msk.shape)
    coef += msk
    coef[:, 1:] += msk[:, :-1]
    coef[1:, :] += msk[:-1, :]
    coef[1:, 1:] += msk[:-1, :-1]
    free = lambda x: 1 * (x == 4)
    nos = lambda x: np.minimum(x, 1)
    ns = param.noslip
    if ns is None or ns is False:
        return free(coef)
    if ns is True:
        return nos(coef)
    sc = free(coef)
    nh = param.halowidth
    if "left" in ns:
        sc[:, nh] = nos(coef[:, nh])
    if "right" in ns:
        sc[:, -nh] = nos(coef[:, -nh])
    if "bottom" in ns:
        sc[nh, :] = nos(coef[nh, :])
    if "top" in ns:
        sc[-nh, :] = nos(coef[-nh, :])
    return sc


class XY(SimpleNamespace):
    """x/y pair (states.py:63)."""


class Mesh:
    """meshes.py:7-111"""

    def __init__(self, param, msk=None):
        self.param = param
        nh = param.halowidth
        self.nx, self.ny = param.nx, param.ny
        self.shape = (param.ny + 2 * nh, param.nx + 2 * nh)
        self.dx, self.dy = param.Lx / self.nx, param.Ly / self.ny
        self.area = self.dx * self.dy
        self.xshift, self.yshift = 1, self.shape[1]
        if msk is None:
            msk = np.zeros(self.shape, dtype=np.int8)
            xs = slice(None) if param.xperiodic else slice(nh, -nh)
            ys = slice(None) if (param.yperiodic or param.ywrap) else slice(nh, -nh)
            msk[ys, xs] = 1
        self.msk = np.ascontiguousarray(msk, dtype=np.int8)
        self.hb = 0
        self.finalize()

    def finalize(self, build_solvers=True):
        p, m = self.param, self.msk
        z = lambda: np.zeros(self.shape, dtype=np.int8)
        self.mskx = z()
        self.mskx[:, 1:] = m[:, 1:] * m[:, :-1]
        self.msky = z()
        self.msky[1:, :] = m[1:, :] * m[:-1, :]
        self.mskv = z()
        self.mskv[1:, 1:] = m[:-1, 1:] * m[:-1, :-1] * m[1:, 1:] * m[1:, :-1]
        self.slipcoef = slipcoef(p, m)
        mo = p.maxorder
        self.oc = XY(x=set_order(m, self.xshift, mo), y=set_order(m, self.yshift, mo))
        mv = (self.slipcoef > 0).astype(np.int8)
        self.ov = XY(x=set_order(mv, -self.xshift, mo) * self.msky,
                     y=set_order(mv, -self.yshift, mo) * self.mskx)
        self.ok = XY(x=set_order(self.mskx, -self.xshift, mo),
                     y=set_order(self.msky, -self.yshift, mo))
        if build_solvers:
            self.poisson_centers = Poisson2D(self, "c")
            self.poisson_vertices = Poisson2D(self, "v")
            if p.model in ("qg", "qgrsw", "rsw"):
                self.qg_helmholtz = Poisson2D(self, "v", maindiag=self.area * p.f0 ** 2 / (p.g * p.H))
                self.qgcoef = p.f0 / p.H

    def fill(self, a):
        """meshes.py:135-143 -- x-periodic halo copy only."""
        if isinstance(a, XY):
            self.fill(a.x)
            self.fill(a.y)
        else:
            n = self.param.halowidth
            if self.param.xperiodic:
                a[:, :n] = a[:, -2 * n:-n]
                a[:, -n:] = a[:, n:2 * n]
            if self.param.ywrap:           # extension, see _DEFAULTS
                a[:n, :] = a[-2 * n:-n, :]
                a[-n:, :] = a[n:2 * n, :]

    def xy(self, which="c"):
        nh = self.param.halowidth
        ix = np.arange(self.nx + 2 * nh) - nh
        iy = np.arange(self.ny + 2 * nh) - nh
        sx = 0.5 if which in ("c", "y") else 0.0
        sy = 0.5 if which in ("c", "x") else 0.0
        return np.meshgrid((ix + sx) * self.dx, (iy + sy) * self.dy)


# --------------------------------------------------------------------------
# elliptic  (elliptic.py:71-203)
# --------------------------------------------------------------------------
def solver_mask(mesh, location):
    """elliptic.py:102-111"""
    msk = mesh.msk if location == "c" else mesh.mskv
    if not (mesh.param.xperiodic or mesh.param.ywrap):
        return msk
    n = mesh.param.halowidth
    m = msk * 1
    if mesh.param.xperiodic:
        m[:, :n] = 0
        m[:, -n:] = 0
    if mesh.param.ywrap:                   # extension: halo rows are images, not unknowns
        m[:n, :] = 0
        m[-n:, :] = 0
    return m


def laplacian(mesh, location, maindiag=0.0):
    """Vectorised assembly of elliptic.py:114-195.  Returns (A_csc, G)."""
    from scipy import sparse
    msk = solver_mask(mesh, location)
    G = np.full(msk.shape, -1, dtype=np.int32)
    G[msk == 1] = np.arange(int(np.sum(msk)), dtype=np.int32)
    ny, nx = G.shape
    N = int(np.sum(G > -1))
    dx2, dy2 = mesh.dy / mesh.dx, mesh.dx / mesh.dy
    xper = mesh.param.xperiodic
    n1 = mesh.param.halowidth if xper else 0
    neg = np.full((ny, nx), -1, dtype=np.int32)
    west, east, south, north = neg.copy(), neg.copy(), neg.copy(), neg.copy()
    west[:, n1 + 1:] = G[:, n1:-1]
    east[:, :nx - 1 - n1] = G[:, 1:nx - n1]
    if xper:
        west[:, :n1 + 1] = G[:, -n1 - 1][:, None]
        east[:, nx - 1 - n1:] = G[:, n1][:, None]
    south[1:, :] = G[:-1, :]
    north[:-1, :] = G[1:, :]
    if mesh.param.ywrap:                   # extension: rows wrap the way elliptic.py:153-160 wraps columns
        n2 = mesh.param.halowidth
        south[:n2 + 1, :] = G[-n2 - 1, :][None, :]
        north[ny - 1 - n2:, :] = G[n2, :][None, :]
    fluid = G > -1
    rows, cols, vals = [], [], []
    offsum = np.zeros((ny, nx))
    # same accumulation order as the reference: W, E, S, N
    for nb, c in ((west, dx2), (east, dx2), (south, dy2), (north, dy2)):
        k = fluid & (nb > -1)
        rows.append(G[k]); cols.append(nb[k]); vals.append(np.full(int(k.sum()), c))
        offsum[k] += c
    rows.append(G[fluid]); cols.append(G[fluid])
    if location == "v":
        vals.append(np.full(N, -2 * (dx2 + dy2) - maindiag))
    else:
        vals.append(-offsum[fluid] - maindiag)
    A = sparse.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(N, N)).tocsc()
    A.sort_indices()
    return A, G


class Poisson2D:
    """elliptic.py:71-87.  The reference assembles and factorises in __init__; the
    oracle does both on first use (same matrix, same SuperLU call), so that a test
    only pays for the solvers its model actually calls."""

    # test knob (not in the reference): steps of iterative refinement after the direct
    # solve, x += LU^-1 (b - A x).  The reference's answer carries the forward error of its
    # LU, cond(A) * eps; comparing a refined and an unrefined run measures how far that
    # alone moves a simulation -- the floor under any parity tolerance.
    refine = 0

    def __init__(self, mesh, location, maindiag=0.0):
        self.mesh, self.location, self.maindiag = mesh, location, maindiag
        self._A = self._G = self._LU = None

    def _assemble(self):
        if self._A is None:
            self._A, self._G = laplacian(self.mesh, self.location, self.maindiag)
            self._k = self._G > -1

    @property
    def A(self):
        self._assemble()
        return self._A

    @property
    def G(self):
        self._assemble()
        return self._G

    @property
    def A_LU(self):
        if self._LU is None:
            import scipy.sparse.linalg as spl
            self._LU = spl.splu(self.A)
        return self._LU

    def solve(self, b, x):
        lu = self.A_LU
        bk = b[self._k]
        xk = lu.solve(bk)
        for _ in range(self.refine):
            xk = xk + lu.solve(bk - self._A @ xk)
        x[self._k] = xk
        self.mesh.fill(x)


# --------------------------------------------------------------------------
# operators  (operators.py)
# --------------------------------------------------------------------------
def addvortexforce(param, mesh, U, omega, du):          # :6-13
    vortexforce(du.x, U.y, omega, mesh.ov.y, mesh.yshift, mesh.xshift, +1, param.vortexforce)
    vortexforce(du.y, U.x, omega, mesh.ov.x, mesh.xshift, mesh.yshift, -1, param.vortexforce)


def div(mesh, U, delta):                                # :104-107
    delta[:, :-1] = -(U.x[:, 1:] - U.x[:, :-1])
    delta[:-1, :] -= U.y[1:, :] - U.y[:-1, :]
    delta *= mesh.msk


def divflux(param, mesh, flx, q, U, dq):                # :21-29
    compflux(flx.x, U.x, q, mesh.oc.x, mesh.xshift, param.compflux)
    compflux(flx.y, U.y, q, mesh.oc.y, mesh.yshift, param.compflux)
    div(mesh, flx, dq)


def addcoriolis(param, mesh, U, du):                    # :32-39
    f = param.f0 * mesh.area * 0.25
    du.x[:-1, 1:-1] += f * (U.y[:-1, :-2] + U.y[1:, :-2] + U.y[:-1, 1:-1] + U.y[1:, 1:-1])
    du.y[1:-1, :-1] -= f * (U.x[:-2, :-1] + U.x[:-2, 1:] + U.x[1:-1, :-1] + U.x[1:-1, 1:])


def addgrad(mesh, phi, du):                             # :49-53
    du.x[:, 1:] -= (phi[:, 1:] - phi[:, :-1]) * mesh.mskx[:, 1:]
    du.y[1:, :] -= (phi[1:, :] - phi[:-1, :]) * mesh.msky[1:, :]


def sharp(mesh, u, U):                                  # :59-64
    U.x[:] = u.x * (1 / mesh.dx ** 2)
    U.y[:] = u.y * (1 / mesh.dy ** 2)


def compute_vorticity(mesh, u, omega):                  # :67-77
    omega[1:, :] = -(u.x[1:, :] - u.x[:-1, :])
    omega[:, 1:] += u.y[:, 1:] - u.y[:, :-1]
    omega *= mesh.slipcoef


def compute_kinetic_energy(param, mesh, u, U, ke):      # :80-93
    ke[:] = 0.0
    m = param.innerproduct
    if m == "classic":
        ke[:, :-1] = +u.x[:, 1:] * U.x[:, 1:] + u.x[:, :-1] * U.x[:, :-1]
        ke[:-1, :] += u.y[1:, :] * U.y[1:, :] + u.y[:-1, :] * U.y[:-1, :]
        ke *= mesh.msk * 0.25
    else:
        innerproduct(ke, U.x, u.x, mesh.ok.x, mesh.xshift, m)
        innerproduct(ke, U.y, u.y, mesh.ok.y, mesh.yshift, m)
        ke *= mesh.msk * 0.5


def compute_pressure(param, mesh, h, p):                # :110-111
    p[:] = (param.g / mesh.area) * (h + mesh.hb)


def pressure_projection(mesh, U, delta, p, u):          # :114-119
    sharp(mesh, u, U)
    div(mesh, U, delta)
    mesh.poisson_centers.solve(-delta * mesh.area, p)
    addgrad(mesh, p, u)
    mesh.fill(u)


def centerstovertices(mesh, a, v, addto=False):         # :126-133
    t = 0.25 * (a[:-1, :-1] + a[1:, :-1] + a[:-1, 1:] + a[1:, 1:])
    if addto:
        v[1:, 1:] += t
    else:
        v[1:, 1:] = t
    v *= mesh.mskv


def verticestocenters(mesh, vh, h):                     # :136-141
    m = mesh.mskv
    coef = m[:-1, :-1] + m[1:, :-1] + m[:-1, 1:] + m[1:, 1:]
    with np.errstate(divide="ignore", invalid="ignore"):
        h[:-1, :-1] = (1 / coef) * (vh[:-1, :-1] + vh[1:, :-1] + vh[:-1, 1:] + vh[1:, 1:])
    h *= mesh.msk


def perpgrad(mesh, psi, u, contravariant=False):        # :144-149
    u.x[:-1, :] = -(psi[1:, :] - psi[:-1, :]) * mesh.mskx[:-1, :]
    u.y[:, :-1] = (psi[:, 1:] - psi[:, :-1]) * mesh.msky[:, :-1]
    if contravariant:
        u.x[:] *= (1 / mesh.dy ** 2)
        u.y[:] *= (1 / mesh.dx ** 2)


def addbuoyancy(mesh, b, du):                           # :152-153
    du.y[1:, :] += (0.5 * mesh.dy) * (b[1:, :] + b[:-1, :]) * mesh.msky[1:, :]


def add_stretching(mesh, pv, h, anomaly):               # :194-201
    f0, H = mesh.param.f0, mesh.param.H
    if anomaly:
        centerstovertices(mesh, h * (-f0 / H), pv, addto=True)
    else:
        h0 = H * mesh.area - mesh.hb
        centerstovertices(mesh, (h - h0) * (-f0 / H), pv, addto=True)


def thickness_from_psi(mesh, psi, h, anomaly):          # :204-211
    f0, H, g = mesh.param.f0, mesh.param.H, mesh.param.g
    verticestocenters(mesh, psi * (f0 * mesh.area / g), h)
    if not anomaly:
        h += (H * mesh.area) - mesh.hb
    h *= mesh.msk


def qg_projection(mesh, u, h, pv, psi, anomaly=False):  # :176-183
    compute_vorticity(mesh, u, pv)
    add_stretching(mesh, pv, h, anomaly)
    mesh.qg_helmholtz.solve(pv, psi)
    thickness_from_psi(mesh, psi, h, anomaly)
    perpgrad(mesh, psi, u)


def set_uv_from_omega(mesh, omega, u, contravariant=False):   # tools.py:6-26
    psi = omega * 0
    mesh.poisson_vertices.solve(omega, psi)
    perpgrad(mesh, psi, u, contravariant=contravariant)
    return psi


# --------------------------------------------------------------------------
# states (states.py:7-17) / equations (equations.py) / integrators
# --------------------------------------------------------------------------
SPECS = {
    "euler": (("u", "U", "omega", "ke", "p", "div", "flx"), ("u",)),
    "boussinesq": (("b", "u", "U", "omega", "ke", "p", "div", "flx"), ("b", "u")),
    "rsw": (("u", "h", "U", "omega", "ke", "p", "flx", "pv"), ("u", "h")),
    "qgrsw": (("u", "h", "U", "omega", "ke", "p", "flx", "pv", "psi"), ("u", "h")),
    "eulerpsi": (("omega", "U", "psi", "vomega", "flx"), ("omega",)),
    "qg": (("pv", "U", "h", "flx", "work", "psi"), ("pv",)),
    "advection": (("q", "U", "flx"), ("q",)),
    "vectoradv": (("v", "U", "omega", "q"), ("v",)),
}
VECTORS = ("u", "U", "flx", "v")


def _alloc(names, shape):
    return SimpleNamespace(**{n: (XY(x=np.zeros(shape), y=np.zeros(shape)) if n in VECTORS
                                  else np.zeros(shape)) for n in names})


def rhs_and_diag(param, mesh):
    model = param.model
    fill = mesh.fill

    if model == "euler":                                 # equations.py:9-24
        def rhs(s, ds):
            addvortexforce(param, mesh, s.U, s.omega, ds.u)
            addgrad(mesh, s.ke, ds.u)
            fill(ds.u)

        def diag(s):
            pressure_projection(mesh, s.U, s.div, s.p, s.u)
            sharp(mesh, s.u, s.U)
            compute_vorticity(mesh, s.u, s.omega)
            compute_kinetic_energy(param, mesh, s.u, s.U, s.ke)
            fill(s.omega); fill(s.ke)

    elif model == "boussinesq":                          # equations.py:27-45
        def rhs(s, ds):
            addvortexforce(param, mesh, s.U, s.omega, ds.u)
            addgrad(mesh, s.ke, ds.u)
            addbuoyancy(mesh, s.b, ds.u)
            divflux(param, mesh, s.flx, s.b, s.U, ds.b)
            fill(ds.u); fill(ds.b)

        def diag(s):
            pressure_projection(mesh, s.U, s.div, s.p, s.u)
            sharp(mesh, s.u, s.U)
            compute_vorticity(mesh, s.u, s.omega)
            compute_kinetic_energy(param, mesh, s.u, s.U, s.ke)
            fill(s.omega); fill(s.ke)

    elif model == "rsw":                                 # equations.py:139-157
        def rhs(s, ds):
            addvortexforce(param, mesh, s.U, s.omega, ds.u)
            addcoriolis(param, mesh, s.U, ds.u)
            addgrad(mesh, s.ke, ds.u)
            addgrad(mesh, s.p, ds.u)
            divflux(param, mesh, s.flx, s.h, s.U, ds.h)
            fill(ds.u); fill(ds.h)

        def diag(s):
            sharp(mesh, s.u, s.U)
            compute_vorticity(mesh, s.u, s.omega)
            compute_kinetic_energy(param, mesh, s.u, s.U, s.ke)
            compute_pressure(param, mesh, s.h, s.p)
            fill(s.omega); fill(s.ke)

    elif model == "qgrsw":                               # equations.py:106-136
        def rhs(s, ds):
            addvortexforce(param, mesh, s.U, s.omega, ds.u)
            addcoriolis(param, mesh, s.U, ds.u)
            divflux(param, mesh, s.flx, s.h, s.U, ds.h)
            qg_projection(mesh, ds.u, ds.h, s.pv, s.psi, anomaly=True)
            fill(ds.u); fill(ds.h)

        def diag(s):
            sharp(mesh, s.u, s.U)
            compute_vorticity(mesh, s.u, s.omega)
            fill(s.omega)
    elif model == "eulerpsi":                            # equations.py:73-86
        def rhs(s, ds):
            divflux(param, mesh, s.flx, s.omega, s.U, ds.omega)
            fill(ds.omega)

        def diag(s):
            centerstovertices(mesh, s.omega, s.vomega)
            mesh.poisson_vertices.solve(s.vomega, s.psi)
            perpgrad(mesh, s.psi, s.U, contravariant=True)
            fill(s.U)

    elif model == "qg":                                  # equations.py:89-103, operators.py:186-191
        def rhs(s, ds):
            divflux(param, mesh, s.flx, s.pv, s.U, ds.pv)
            fill(ds.pv)

        def diag(s):
            pvback = mesh.hb * mesh.qgcoef
            centerstovertices(mesh, s.pv - pvback, s.work)
            if param.beta != 0:
                raise NotImplementedError("qg with beta needs the user-set mesh.f")
            mesh.qg_helmholtz.solve(s.work, s.psi)
            perpgrad(mesh, s.psi, s.U, contravariant=True)

    elif model == "advection":                           # equations.py:160-170
        def rhs(s, ds):
            divflux(param, mesh, s.flx, s.q, s.U, ds.q)
            fill(ds.q)

        def diag(s):
            pass

    elif model == "vectoradv":                           # equations.py:173-187
        def rhs(s, ds):
            addvortexforce(param, mesh, s.U, s.omega, ds.v)
            addgrad(mesh, s.q, ds.v)
            fill(ds.v)

        def diag(s):
            compute_vorticity(mesh, s.v, s.omega)
            compute_kinetic_energy(param, mesh, s.v, s.U, s.q)
            s.q[:] *= 2
            fill(s.omega); fill(s.q)
    else:
        raise NotImplementedError(model)

    if param.tracer and param.tracer != "None":          # equations.py:206-208, 217-226
        model_rhs, name = rhs, param.tracer

        def rhs(s, ds):
            model_rhs(s, ds)
            divflux(param, mesh, s.flx, getattr(s, name), s.U, getattr(ds, name))
    return rhs, diag


def specs_of(param):                                     # states.py:22-34
    names, prog = SPECS[param.model]
    if param.tracer and param.tracer != "None":
        newprog = prog + (param.tracer,)
        names = newprog + names[len(prog):]
        if "flx" not in names:
            names = names + ("flx",)
        prog = newprog
    return names, prog


def _leaves(ns, names):
    out = []
    for n in names:
        v = getattr(ns, n)
        out += [v.x, v.y] if isinstance(v, XY) else [v]
    return out


RK_COEFS = {  # integrators.py:82-124 (incremental form)
    "ef": lambda dt: [(dt,)],
    "rk3": lambda dt: [(dt,), (-3 * dt / 4, dt / 4), (-dt / 12, -dt / 12, 2 * dt / 3)],
    "enrk3": lambda dt: [(dt / 3,), (-dt / 3 - 5 * dt / 48, 15 * dt / 16),
                         (5 * dt / 48 + dt / 10, -7 * dt / 16, 2 * dt / 5)],
}


class Model:
    """Minimal driver: Mesh + State + RK integrator (model.py:14-27,67-87)."""

    def __init__(self, param, msk=None):
        self.param = param
        self.mesh = Mesh(param, msk)
        self._alloc_state()

    def _alloc_state(self):
        names, prog = specs_of(self.param)
        self.prognostic = prog
        self.state = _alloc(names, self.mesh.shape)
        nst = 1 if self.param.integrator == "ef" else 3     # LFRA: scratch = [sb, sa, ds]
        self.scratch = [_alloc(prog, self.mesh.shape) for _ in range(nst)]
        self.rhs, self.diag = rhs_and_diag(self.param, self.mesh)
        self.t, self.ite = 0.0, 0

    def refinalize(self):
        self.mesh.finalize()
        self.rhs, self.diag = rhs_and_diag(self.param, self.mesh)

    def compute_dt(self):                                # model.py:71-87
        p = self.param
        if p.dt > 0:
            return p.dt
        if p.model == "rsw":
            c = (p.g * p.H) ** 0.5
            maxU = c / self.mesh.dx + c / self.mesh.dy
        else:
            U = self.state.U
            maxU = np.max(np.abs(U.x)) + np.max(np.abs(U.y)) + 1e-99
        return min(p.cfl / maxU, p.dtmax)

    def step(self, dt=None):
        """integrators.py:76-79,92-107 + addto_list :154-174"""
        dt = self.compute_dt() if dt is None else dt
        s = self.state
        ys = _leaves(s, self.prognostic)
        if self.param.integrator == "LFRA":
            self._step_lfra(dt, ys)
            self.t += dt
            self.ite += 1
            return dt
        for k, coefs in enumerate(RK_COEFS[self.param.integrator](dt)):
            self.rhs(s, self.scratch[k])
            xs = [_leaves(self.scratch[i], self.prognostic) for i in range(k + 1)]
            for f, y in enumerate(ys):
                # y[:] += sum(c*x ...): Python's sum starts from int 0
                acc = 0
                for c, x in zip(coefs, xs):
                    acc = acc + c * x[f]
                y[:] += acc
            self.diag(s)
        self.t += dt
        self.ite += 1
        return dt

    def _step_lfra(self, dt, ys):
        """integrators.py:39-53 with copyto (:140-151), rightpermute (:133-137)"""
        s, gamma = self.state, self.param.RAgamma
        sb, sa, ds = (_leaves(x, self.prognostic) for x in self.scratch)
        self.rhs(s, self.scratch[2])
        if self.ite == 0:
            for f, y in enumerate(ys):
                sb[f][:] = y
                sa[f][:] = y
                y[:] += 0 + dt * ds[f]
        else:
            for f, y in enumerate(ys):
                sa[f][:] += 0 + (2 * dt) * ds[f]
                y[:] += ((0 + gamma * sa[f]) + gamma * sb[f]) + (-2 * gamma) * y
                sb[f][:] = y           # rightpermute(sa, s, sb)
                y[:] = sa[f]
                sa[f][:] = sb[f]
        self.diag(s)
