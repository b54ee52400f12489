"""Host-side discrete operators that experiment scripts call while building
initial conditions (reference: src/fluids2d/operators.py).  They act on the
caller's numpy arrays, once, before the run; inside the time loop the same
operators are fused CUDA kernels (csrc/step.cu) and the elliptic solves always
run on the device (``mesh.poisson_*.solve``)."""
import numpy as np


def sharp(mesh, u, U):                                   # operators.py:59-64
    U.x[:] = u.x * (1 / mesh.dx ** 2)
    U.y[:] = u.y * (1 / mesh.dy ** 2)


def perpgrad(mesh, psi, u, contravariant=False):         # operators.py:144-149
    u.x[:-1, :] = -(psi[1:, :] - psi[:-1, :]) * mesh.mskx[:-1, :]
    u.y[:, :-1] = (psi[:, 1:] - psi[:, :-1]) * mesh.msky[:, :-1]
    if contravariant:
        u.x[:] *= (1 / mesh.dy ** 2)
        u.y[:] *= (1 / mesh.dx ** 2)


def compute_vorticity(mesh, u, omega):                   # operators.py:67-77
    omega[1:, :] = -(u.x[1:, :] - u.x[:-1, :])
    omega[:, 1:] += u.y[:, 1:] - u.y[:, :-1]
    omega *= mesh.slipcoef


def centerstovertices(mesh, a, v, addto=False):          # operators.py:126-133
    t = 0.25 * (a[:-1, :-1] + a[1:, :-1] + a[:-1, 1:] + a[1:, 1:])
    if addto:
        v[1:, 1:] += t
    else:
        v[1:, 1:] = t
    v *= mesh.mskv


def verticestocenters(mesh, vh, h):                      # operators.py:136-141
    m = mesh.mskv
    coef = m[:-1, :-1] + m[1:, :-1] + m[:-1, 1:] + m[1:, 1:]
    with np.errstate(divide="ignore", invalid="ignore"):
        h[:-1, :-1] = (1 / coef) * (vh[:-1, :-1] + vh[1:, :-1] + vh[:-1, 1:] + vh[1:, 1:])
    h *= mesh.msk


def compute_pv(param, mesh, omega, h, pv):               # operators.py:42-46
    f = param.f0 * mesh.area
    centerstovertices(mesh, h, pv)
    k = pv > 0
    pv[k] = (f + omega[k]) / pv[k]
    pv *= mesh.area


def add_stretching(mesh, pv, h, anomaly):                # operators.py:194-201
    f0, H = mesh.param.f0, mesh.param.H
    if anomaly:
        centerstovertices(mesh, h * (-f0 / H), pv, addto=True)
    else:
        h0 = H * mesh.area - mesh.hb
        centerstovertices(mesh, (h - h0) * (-f0 / H), pv, addto=True)


def thickness_from_psi(mesh, psi, h, anomaly):           # operators.py:204-211
    f0, H, g = mesh.param.f0, mesh.param.H, mesh.param.g
    verticestocenters(mesh, psi * (f0 * mesh.area / g), h)
    if not anomaly:
        h += (H * mesh.area) - mesh.hb
    h *= mesh.msk


def qg_projection(mesh, u, h, pv, psi, anomaly=False):   # operators.py:176-183
    compute_vorticity(mesh, u, pv)
    add_stretching(mesh, pv, h, anomaly)
    mesh.qg_helmholtz.solve(pv, psi)                     # device Helmholtz solve
    thickness_from_psi(mesh, psi, h, anomaly)
    perpgrad(mesh, psi, u)


def compute_streamfunction(mesh, vomega, psi):           # operators.py:122-123
    mesh.poisson_vertices.solve(vomega, psi)
