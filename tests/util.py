"""Shared helpers for the parity tests: golden loading, oracle drivers, metrics."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

EULER_CASES = ["euler40", "vortex", "vortex_triangle", "disc_island",
               # two connected fluid components: one Neumann null-space constant each
               "euler_lake", "euler_two_basins", "xper_noslip",
               "euler_enrk3_upwind", "euler_centered_ef", "euler_cweno"]
ALL_CASES = EULER_CASES + ["rsw", "rsw_islands", "qgrsw_topo", "qgrsw_islands",
                           "warm_bubble", "lock_exchange",
                           # the remaining models on the same kernels
                           "advection", "advection_disc_upwind", "eulerpsi", "qg", "vectoradv",
                           # leap-frog + Robert-Asselin filter
                           "euler_lfra", "rsw_lfra",
                           # param.tracer: the extra advected scalar (equations.py:217-226)
                           "euler_tracer", "lock_exchange_tracer", "rsw_tracer"]


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN, f"case_{name}.npz"))
        self.meta = json.loads(str(self.z["meta"]))
        self.param = self.meta["param"]
        self.dts = self.meta["dts"]
        self.nsteps = self.meta["nsteps"]
        self.msk = self.z["msk"]
        self.hb = self.z["hb"] if "hb" in self.z.files else None

    def fields(self, tag):
        pre = tag + "/"
        return {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}


def set_state(state, fields):
    """Copy a {name or name.x: array} dict into a namespace/namedtuple state in place."""
    for k, v in fields.items():
        if "." in k:
            n, c = k.split(".")
            getattr(getattr(state, n), c)[...] = v
        else:
            getattr(state, k)[...] = v


def get_state(state, names):
    out = {}
    for k in names:
        if "." in k:
            n, c = k.split(".")
            out[k] = getattr(getattr(state, n), c)
        else:
            out[k] = getattr(state, k)
    return out


def oracle_model(orc, g):
    p = orc.make_param(**g.param)
    m = orc.Model(p, msk=g.msk.copy())
    if g.hb is not None:
        m.mesh.hb = g.hb.copy()
    set_state(m.state, g.fields("init"))
    return m


def rel_l2(a, b, w=None):
    """relative L2 of a-b against b over the cells where w != 0 (finite entries only)."""
    if w is None:
        w = np.ones(a.shape, dtype=bool)
    k = (np.asarray(w) != 0) & np.isfinite(b)
    d = a[k] - b[k]
    nb = np.sqrt(np.sum(b[k] ** 2))
    if nb == 0:
        return float(np.sqrt(np.sum(d ** 2)))
    return float(np.sqrt(np.sum(d ** 2)) / nb)


def remove_component_means(p, msk):
    """p minus its mean over each connected fluid component (Neumann null space)."""
    from scipy import ndimage
    lab, n = ndimage.label(msk != 0)
    q = p.copy()
    for c in range(1, n + 1):
        k = lab == c
        q[k] -= q[k].mean()
    return q


def field_mask(mesh, name):
    """which cells count for the parity metric of a given field."""
    base = name.split(".")[0]
    if name.endswith(".x"):
        return mesh.mskx
    if name.endswith(".y"):
        return mesh.msky
    if base in ("omega", "pv", "psi", "vomega", "work"):
        return mesh.mskv
    return mesh.msk


# --------------------------------------------------------------------------
# CUDA-side drivers
# --------------------------------------------------------------------------
_PARAM_DEFAULTS = dict(
    model="euler", nx=40, ny=40, Lx=1.0, Ly=1.0, xperiodic=False, yperiodic=False,
    halowidth=3, noslip=None, f0=10.0, beta=0.0, g=1, H=1, dt=0.0, cfl=0.9, dtmax=9e99,
    integrator="rk3", compflux="weno", vortexforce="weno", innerproduct="weno",
    maxorder=6, tracer=None, RAgamma=0.1)


def plain_param(**kw):
    from types import SimpleNamespace
    d = dict(_PARAM_DEFAULTS)
    d.update(kw)
    return SimpleNamespace(**d)


def engine_for(g, exact=False, **solver_kw):
    """Engine loaded with a golden case's mask, topography and initial state."""
    from fluids2d_b200._cabi import Engine
    e = Engine(plain_param(**g.param), exact=exact, **solver_kw)
    e.set_mask(g.msk)
    if g.hb is not None:
        e.set_topography(g.hb)
    for k, v in g.fields("init").items():
        if k.startswith("flx") and g.param["model"] == "euler":
            continue
        e.upload(k, v)
    return e


def mesh_masks(e):
    from types import SimpleNamespace
    return SimpleNamespace(msk=e.mesh_array("msk"), mskx=e.mesh_array("mskx"),
                           msky=e.mesh_array("msky"), mskv=e.mesh_array("mskv"))
