"""Generate the golden vectors in this directory from the LIVE reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Each ``case_*.npz`` holds: ``meta`` (JSON: the Param attributes that differ
from the defaults, number of steps, dt per step), ``msk`` (int8 mask given to
``mesh.finalize()``), ``init/<field>`` the full state before the first step,
``s1/<field>`` after one step and ``final/<field>`` after ``nsteps`` steps
(float64, whole haloed arrays).  ``ops_*.npz`` hold per-kernel input/output
vectors, ``solve_*.npz`` direct-solve vectors.  The reference is
bit-deterministic run to run, so re-running this script reproduces the files.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import refshim  # noqa: E402

f2d = refshim.load()
from fluids2d import weno as rweno  # noqa: E402
from fluids2d.operators import qg_projection  # noqa: E402

NSTEPS = 10


def gaussian(x, y, x0, y0, r):
    return np.exp(-((x - x0) ** 2 + (y - y0) ** 2) / (2 * r ** 2))


def snapshot(state, tag, out):
    for name in state._fields:
        v = getattr(state, name)
        if hasattr(v, "_fields"):
            out[f"{tag}/{name}.x"] = v.x.copy()
            out[f"{tag}/{name}.y"] = v.y.copy()
        else:
            out[f"{tag}/{name}"] = v.copy()


def make_param(**kw):
    p = f2d.Param()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def dipole_ic(model, x0, y0, r0, d):
    """vortex.py:8-42 (euler branch)"""
    x, y = model.mesh.xy("v")
    om = model.state.omega
    om[:, :] = gaussian(x, y, x0 + d, y0, r0) - gaussian(x, y, x0 - d, y0, r0)
    om *= model.mesh.mskv * model.mesh.area
    f2d.tools.set_uv_from_omega(model, om, model.state.u)
    model.integrator.diag(model.state)


def run_case(name, pkw, mask_fn=None, ic=None, nsteps=NSTEPS, extra_mesh=None):
    p = make_param(**pkw)
    model = f2d.Model(p)
    if mask_fn is not None:
        mask_fn(model)
        model.mesh.finalize()
    if extra_mesh is not None:
        extra_mesh(model.mesh)
    ic(model)
    out = {"msk": model.mesh.msk.copy()}
    hb = getattr(model.mesh, "hb", 0)
    if isinstance(hb, np.ndarray):
        out["hb"] = hb.copy()
    snapshot(model.state, "init", out)
    dts = []
    for k in range(nsteps):
        model.set_dt()
        dts.append(model.time.dt)
        model.step(1)
        if k == 0:
            snapshot(model.state, "s1", out)
    snapshot(model.state, "final", out)
    meta = dict(param=pkw, nsteps=nsteps, dts=dts)
    out["meta"] = np.array(json.dumps(meta))
    path = os.path.join(HERE, f"case_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: shape={model.mesh.shape} dts[0]={dts[0]:.4g} -> {os.path.getsize(path)/1e3:.0f} kB")
    return model


# ---------------------------------------------------------------- cases ---
def case_euler40():
    # tests/test_models.py:9-15 set-up; adaptive dt (cfl 0.9), first 10 steps
    run_case("euler40", dict(model="euler"), ic=lambda m: dipole_ic(m, 0.5, 0.5, 0.05, 0.05))


def case_vortex():
    # src/experiments/vortex.py:59-71, fixed dt = 0.2
    kw = dict(model="euler", Lx=2.0, ny=100, nx=200, dt=0.2, f0=0.0, noslip=False)
    run_case("vortex", kw, ic=lambda m: dipole_ic(m, 1.0, 0.5, 0.05, 0.05))


def case_vortex_triangle():
    # vortex.py:45-51 mask (triangle at the bottom), half resolution
    kw = dict(model="euler", Lx=2.0, ny=50, nx=100, dt=0.4, f0=0.0, noslip=False)

    def mask(model):
        x, y = model.mesh.xy()
        model.mesh.msk[y < 0.2 - 0.5 * np.abs(x - model.param.Lx / 2)] = 0

    run_case("vortex_triangle", kw, mask_fn=mask, ic=lambda m: dipole_ic(m, 1.0, 0.5, 0.05, 0.05))


def case_disc_island():
    kw = dict(model="euler", nx=64, ny=64, noslip=True)

    def mask(model):
        x, y = model.mesh.xy()
        msk = model.mesh.msk
        msk[(x - 0.5) ** 2 + (y - 0.5) ** 2 > 0.48 ** 2] = 0
        msk[(x - 0.6) ** 2 + (y - 0.35) ** 2 < 0.07 ** 2] = 0

    run_case("disc_island", kw, mask_fn=mask, ic=lambda m: dipole_ic(m, 0.5, 0.6, 0.06, 0.06))


def lake_mask(model):
    """closed box with a wall ring that encloses a lake: two connected fluid
    components, each with its own Neumann null-space constant (elliptic.py:186-190)"""
    x, y = model.mesh.xy()
    msk = model.mesh.msk
    r = np.sqrt((x - 0.62) ** 2 + (y - 0.55) ** 2)
    msk[(r > 0.17) & (r < 0.22)] = 0


def case_euler_lake():
    kw = dict(model="euler", nx=72, ny=64, noslip=False)

    def ic(model):
        x, y = model.mesh.xy("v")
        om = model.state.omega
        # a dipole in the basin and a single vortex inside the lake
        om[:, :] = (gaussian(x, y, 0.25, 0.3, 0.05) - gaussian(x, y, 0.25, 0.42, 0.05)
                    + 0.8 * gaussian(x, y, 0.6, 0.57, 0.04))
        om *= model.mesh.mskv * model.mesh.area
        f2d.tools.set_uv_from_omega(model, om, model.state.u)
        model.integrator.diag(model.state)
    run_case("euler_lake", kw, mask_fn=lake_mask, ic=ic)


def case_euler_two_basins():
    # a wall three cells thick splits the box into two disconnected basins
    kw = dict(model="euler", nx=80, ny=48, Lx=1.6, noslip=True)

    def mask(model):
        x, y = model.mesh.xy()
        model.mesh.msk[np.abs(x - 0.7) < 0.03] = 0

    def ic(model):
        x, y = model.mesh.xy("v")
        om = model.state.omega
        om[:, :] = (gaussian(x, y, 0.35, 0.55, 0.06) - gaussian(x, y, 0.35, 0.4, 0.06)
                    - gaussian(x, y, 1.2, 0.5, 0.07))
        om *= model.mesh.mskv * model.mesh.area
        f2d.tools.set_uv_from_omega(model, om, model.state.u)
        model.integrator.diag(model.state)
    run_case("euler_two_basins", kw, mask_fn=mask, ic=ic)


def case_xper_noslip():
    # non-square cells (dx != dy), odd-ish sizes, per-wall no-slip
    kw = dict(model="euler", nx=72, ny=40, Lx=1.5, Ly=1.0, xperiodic=True, noslip=["bottom"])
    run_case("xper_noslip", kw, ic=lambda m: dipole_ic(m, 0.1, 0.4, 0.07, 0.07))


def case_euler_enrk3_upwind():
    kw = dict(model="euler", nx=48, ny=48, integrator="enrk3", vortexforce="upwind",
              innerproduct="classic", maxorder=4)
    run_case("euler_enrk3_upwind", kw, ic=lambda m: dipole_ic(m, 0.5, 0.5, 0.06, 0.06))


def case_euler_centered_ef():
    kw = dict(model="euler", nx=48, ny=40, integrator="ef", vortexforce="centered",
              innerproduct="centered", cfl=0.2)
    run_case("euler_centered_ef", kw, ic=lambda m: dipole_ic(m, 0.5, 0.5, 0.06, 0.06))


def case_euler_cweno():
    kw = dict(model="euler", nx=48, ny=40, vortexforce="cweno", innerproduct="cweno")
    run_case("euler_cweno", kw, ic=lambda m: dipole_ic(m, 0.5, 0.5, 0.06, 0.06))


def case_euler_lfra():
    # the LFRA settings of geos_adj.py:77-81 / tracer_advection.py:91-94 on the Euler dipole
    kw = dict(model="euler", nx=48, ny=40, integrator="LFRA", cfl=0.5, compflux="centered",
              vortexforce="centered", innerproduct="classic")
    run_case("euler_lfra", kw, ic=lambda m: dipole_ic(m, 0.5, 0.5, 0.06, 0.06))


def case_rsw_lfra():
    kw = dict(model="rsw", nx=48, ny=48, dtmax=1, f0=10.0, integrator="LFRA", cfl=0.5, compflux="centered",
              vortexforce="centered", innerproduct="classic", RAgamma=0.05)
    run_case("rsw_lfra", kw, ic=lambda m: rsw_ic(m))


def rsw_ic(model, amp=0.2, r0=0.1, y0=0.5, sub_hb=False):
    """geos_adj.py:12-49 (flow='dipole'); rsw_with_topo.py:12-53 when sub_hb"""
    mesh = model.mesh
    x, y = mesh.xy("c")
    x0, d = 0.5, r0
    h = model.state.h
    H = model.param.H
    h[:] = H + amp * (gaussian(x, y, x0 + d, y0, r0) - gaussian(x, y, x0 - d, y0, r0))
    h *= (mesh.msk * mesh.area)
    if sub_hb:
        h -= mesh.hb
    s = model.state
    if model.param.model == "qgrsw":
        qg_projection(mesh, s.u, s.h, s.pv, s.psi)
    model.integrator.diag(model.state)


def island_mask(model):
    x, y = model.mesh.xy()
    msk = model.mesh.msk
    msk[(x - 0.25) ** 2 + (y - 0.3) ** 2 < 0.06 ** 2] = 0
    msk[(x - 0.7) ** 2 + (y - 0.75) ** 2 < 0.08 ** 2] = 0
    msk[(x - 0.8) ** 2 + (y - 0.2) ** 2 < 0.05 ** 2] = 0


def case_rsw():
    kw = dict(model="rsw", nx=64, ny=64, dtmax=1, f0=10.0)
    run_case("rsw", kw, ic=lambda m: rsw_ic(m))


def case_rsw_islands():
    kw = dict(model="rsw", nx=64, ny=56, dtmax=1, f0=10.0, noslip=True)
    run_case("rsw_islands", kw, mask_fn=island_mask, ic=lambda m: rsw_ic(m))


def set_topo(mesh):
    # rsw_with_topo.py:96-99
    x, y = mesh.xy()
    mesh.hb = 0.2 * gaussian(x, y, 0.3, 0.7, 0.05) * mesh.area * mesh.msk


def case_qgrsw_topo():
    kw = dict(model="qgrsw", nx=64, ny=64, dtmax=1, f0=10.0)
    run_case("qgrsw_topo", kw, extra_mesh=set_topo,
             ic=lambda m: rsw_ic(m, r0=0.08, y0=0.3, sub_hb=True))


def case_qgrsw_islands():
    kw = dict(model="qgrsw", nx=64, ny=56, dtmax=1, f0=10.0)
    run_case("qgrsw_islands", kw, mask_fn=island_mask, extra_mesh=set_topo,
             ic=lambda m: rsw_ic(m, r0=0.08, y0=0.5, sub_hb=True))


def bubble_ic(model):
    # warm_bubble.py:14-20
    mesh = model.mesh
    x, y = mesh.xy()
    b = model.state.b
    b[:, :] = y + 0.1 * gaussian(x, y, 1, 0.25, 0.02 * 4)
    b *= mesh.msk
    model.integrator.diag(model.state)


def case_warm_bubble():
    kw = dict(model="boussinesq", nx=96, Lx=2.0, ny=48, cfl=0.9, dtmax=1e-1)
    run_case("warm_bubble", kw, ic=bubble_ic)


def lock_ic(model):
    # lockexchange.py:9-17
    x, y = model.mesh.xy("c")
    b = model.state.b
    Lx, dx = model.param.Lx, model.mesh.dx
    st = lambda xx, x0: np.tanh((xx - x0) / (3 * dx))
    b[:, :] = (-st(x, Lx / 4) + st(x, Lx / 2))
    b *= model.mesh.msk
    model.mesh.fill(b)
    model.integrator.diag(model.state)


def case_lock_exchange():
    kw = dict(model="boussinesq", nx=100, Lx=5.0, ny=20, cfl=0.9, dtmax=1.0, xperiodic=True)
    run_case("lock_exchange", kw, ic=lock_ic)


# ---- param.tracer: one more advected scalar (states.py:23-34, equations.py:217-226) ----
def checker_tracer(model, name):
    # tracer_advection.py:44-48
    x, y = model.mesh.xy("c")
    q = getattr(model.state, name)
    q[:, :] = (np.round(x * 8) % 2 + np.round(y * 8) % 2) / 2
    q *= model.mesh.msk


def case_euler_tracer():
    kw = dict(model="euler", nx=64, ny=64, noslip=True, tracer="dye")

    def mask(model):
        x, y = model.mesh.xy()
        msk = model.mesh.msk
        msk[(x - 0.5) ** 2 + (y - 0.5) ** 2 > 0.48 ** 2] = 0
        msk[(x - 0.6) ** 2 + (y - 0.35) ** 2 < 0.07 ** 2] = 0

    def ic(model):
        checker_tracer(model, "dye")
        dipole_ic(model, 0.5, 0.6, 0.06, 0.06)

    run_case("euler_tracer", kw, mask_fn=mask, ic=ic)


def case_lock_exchange_tracer():
    # x-periodic: the tracer tendency is NOT filled by the reference, its halo columns
    # evolve on their own (equations.py:217-222 sits outside the model's fill)
    kw = dict(model="boussinesq", nx=100, Lx=5.0, ny=20, cfl=0.9, dtmax=1.0, xperiodic=True, tracer="c")

    def ic(model):
        checker_tracer(model, "c")
        lock_ic(model)

    run_case("lock_exchange_tracer", kw, ic=ic)


def case_rsw_tracer():
    kw = dict(model="rsw", nx=64, ny=56, dtmax=1, f0=10.0, tracer="age")

    def ic(model):
        checker_tracer(model, "age")
        rsw_ic(model)

    run_case("rsw_tracer", kw, mask_fn=island_mask, ic=ic)


# ---- the scalar-transport / stream-function models ----------
def case_advection():
    # tests/test_models.py:17-26 set-up (body rotation + checkerboard tracer)
    def ic(model):
        mesh = model.mesh
        omega = (4 * np.pi / 100) * np.ones(mesh.shape) * mesh.mskv
        f2d.tools.set_uv_from_omega(model, omega, model.state.U)
        x, y = mesh.xy("c")
        q = model.state.q
        q[:, :] = (np.round(x * 8) % 2 + np.round(y * 8) % 2) / 2
        q *= mesh.msk
    run_case("advection", dict(model="advection"), ic=ic)


def case_advection_disc_upwind():
    def mask(model):
        x, y = model.mesh.xy()
        model.mesh.msk[(x - 0.5) ** 2 + (y - 0.5) ** 2 > 0.47 ** 2] = 0

    def ic(model):
        mesh = model.mesh
        x, y = mesh.xy("v")
        omega = (gaussian(x, y, 0.4, 0.5, 0.08) - gaussian(x, y, 0.6, 0.5, 0.08)) * mesh.mskv
        f2d.tools.set_uv_from_omega(model, omega, model.state.U)
        xc, yc = mesh.xy("c")
        model.state.q[:, :] = gaussian(xc, yc, 0.5, 0.65, 0.1) * mesh.msk
    run_case("advection_disc_upwind", dict(model="advection", nx=56, ny=48, compflux="upwind", maxorder=4),
             mask_fn=mask, ic=ic)


def case_eulerpsi():
    # vortex.py:22-25 (eulerpsi branch): vorticity at cell centres
    def ic(model):
        x, y = model.mesh.xy("c")
        om = model.state.omega
        om[:, :] = gaussian(x, y, 1.05, 0.5, 0.05) - gaussian(x, y, 0.95, 0.5, 0.05)
        om *= model.mesh.msk * model.mesh.area
        model.integrator.diag(model.state)
    run_case("eulerpsi", dict(model="eulerpsi", Lx=2.0, ny=50, nx=100, dt=0.4), ic=ic)


def case_qg():
    # geos_adj.py:12-49 with model = "qg"
    def ic(model):
        mesh = model.mesh
        x, y = mesh.xy("c")
        h = model.state.h
        h[:] = model.param.H + 0.2 * (gaussian(x, y, 0.6, 0.5, 0.1) - gaussian(x, y, 0.4, 0.5, 0.1))
        h *= (mesh.msk * mesh.area)
        s = model.state
        qg_projection(mesh, s.U, s.h, s.pv, s.psi)
        model.integrator.diag(model.state)
    run_case("qg", dict(model="qg", nx=64, ny=56, dtmax=1, f0=10.0), mask_fn=island_mask, ic=ic)


def case_vectoradv():
    # src/experiments/vector_advection.py:10-30, 66-76 (disc domain, body rotation)
    def mask(model):
        x, y = model.mesh.xy()
        model.mesh.msk[(x - 0.5) ** 2 + (y - 0.5) ** 2 > 0.5 ** 2] = 0

    def ic(model):
        mesh = model.mesh
        omega = (4 * np.pi / 100) * np.ones(mesh.shape) * mesh.mskv
        f2d.tools.set_uv_from_omega(model, omega, model.state.U)
        x, y = mesh.xy("x")
        vx = model.state.v.x
        vx[:, :] = gaussian(x, y, 0.7, 0.5, 0.05)
        vx *= mesh.mskx
        model.integrator.diag(model.state)
    run_case("vectoradv", dict(model="vectoradv", nx=50, ny=50), mask_fn=mask, ic=ic)


def run_to_tend():
    """tests/test_models.py:9-15: default 40x40 Euler dipole, model.run() to
    tend = 10 with the adaptive CFL step.  The reference's own assertion
    (ite == 25) is stale against its current default cfl = 0.9; the live code
    gives the count stored here."""
    p = make_param(model="euler", tend=10)
    model = f2d.Model(p)
    dipole_ic(model, 0.5, 0.5, 0.05, 0.05)
    out = {}
    snapshot(model.state, "init", out)
    model.run()
    snapshot(model.state, "final", out)
    out["meta"] = np.array(json.dumps(dict(param=dict(model="euler", tend=10), ite=model.time.ite,
                                           t=model.time.t, dt_last=model.time.dt)))
    path = os.path.join(HERE, "run_euler40.npz")
    np.savez_compressed(path, **out)
    print(f"run_euler40: ite={model.time.ite} t={model.time.t} -> {os.path.getsize(path)/1e3:.0f} kB")


# ---------------------------------------------------- per-kernel vectors ---
def ops_vectors():
    """weno.VortexForce / InnerProduct / CompFlux for all 4 methods on a masked,
    x-periodic-free mesh whose order arrays contain 0, 2, 4 and 6."""
    rng = np.random.default_rng(1234)
    out = {}
    for tag, maxorder in (("o6", 6), ("o4", 4), ("o2", 2)):
        p = make_param(model="euler", nx=44, ny=36, maxorder=maxorder)
        model = f2d.Model(p)
        x, y = model.mesh.xy()
        model.mesh.msk[(x - 0.4) ** 2 + (y - 0.6) ** 2 < 0.12 ** 2] = 0
        model.mesh.msk[y < 0.15 - 0.4 * np.abs(x - 0.7)] = 0
        model.mesh.finalize()
        mesh = model.mesh
        shp = mesh.shape
        q = rng.standard_normal(shp)
        q[:, ::7] *= 1e-3      # smooth-ish and rough regions both present
        Ux, Uy = rng.standard_normal(shp), rng.standard_normal(shp)
        Ux[5:9, :] = 0.0       # exercise the U == 0 (not > 0) branch
        ux, uy = rng.standard_normal(shp), rng.standard_normal(shp)
        ke0 = rng.standard_normal(shp)
        for k, v in dict(q=q, Ux=Ux, Uy=Uy, ux=ux, uy=uy, ke0=ke0, msk=mesh.msk,
                         ocx=mesh.oc.x, ocy=mesh.oc.y, ovx=mesh.ov.x, ovy=mesh.ov.y,
                         okx=mesh.ok.x, oky=mesh.ok.y, mskx=mesh.mskx, msky=mesh.msky,
                         mskv=mesh.mskv, slipcoef=mesh.slipcoef.astype(np.float64)).items():
            out[f"{tag}/{k}"] = v.copy()
        xs, ys = mesh.xshift, mesh.yshift
        for m in ("weno", "upwind", "centered", "cweno"):
            fx, fy = np.full(shp, 7.0), np.full(shp, 7.0)
            rweno.compflux(fx, Ux, q, mesh.oc.x, xs, m)
            rweno.compflux(fy, Uy, q, mesh.oc.y, ys, m)
            dux, duy = np.full(shp, 7.0), np.full(shp, 7.0)
            rweno.vortexforce(dux, Uy, q, mesh.ov.y, ys, xs, +1, m)
            rweno.vortexforce(duy, Ux, q, mesh.ov.x, xs, ys, -1, m)
            ke = ke0.copy()
            rweno.innerproduct(ke, Ux, ux, mesh.ok.x, xs, m)
            rweno.innerproduct(ke, Uy, uy, mesh.ok.y, ys, m)
            for k, v in dict(flx_x=fx, flx_y=fy, du_x=dux, du_y=duy, ke=ke).items():
                out[f"{tag}/{m}/{k}"] = v
    # scalar reconstructions
    a = rng.standard_normal((6, 400))
    a[:, :50] *= 1e-9
    a[:, 50:100] = np.round(a[:, 50:100])      # exact ties / zero smoothness
    out["scalar/args"] = a
    out["scalar/weno5z"] = np.array([rweno.weno5z(*a[:5, i]) for i in range(a.shape[1])])
    out["scalar/weno3z"] = np.array([rweno.weno3z(*a[:3, i]) for i in range(a.shape[1])])
    U = rng.standard_normal(a.shape[1])
    U[::5] = 0.0
    out["scalar/U"] = U
    for m, (f1, f3, f5) in rweno._fluxes.items():
        out[f"scalar/{m}/f5"] = np.array([f5(U[i], *a[:6, i]) for i in range(a.shape[1])])
        out[f"scalar/{m}/f3"] = np.array([f3(U[i], *a[:4, i]) for i in range(a.shape[1])])
        out[f"scalar/{m}/f1"] = np.array([f1(U[i], *a[:2, i]) for i in range(a.shape[1])])
    path = os.path.join(HERE, "ops_weno.npz")
    np.savez_compressed(path, **out)
    print(f"ops_weno -> {os.path.getsize(path)/1e3:.0f} kB")


def solve_vectors():
    """Poisson2D.solve for 'c', 'v' and Helmholtz on four masks."""
    rng = np.random.default_rng(99)
    out = {}
    cases = {
        "closed": (dict(model="rsw", nx=48, ny=40, Lx=1.2), None),
        "xper": (dict(model="rsw", nx=50, ny=36, xperiodic=True), None),
        "islands": (dict(model="rsw", nx=64, ny=56), island_mask),
        "triangle": (dict(model="rsw", nx=60, ny=30, Lx=2.0),
                     lambda m: m.mesh.msk.__setitem__(
                         m.mesh.xy()[1] < 0.2 - 0.5 * np.abs(m.mesh.xy()[0] - 1.0), 0)),
    }
    for name, (kw, mask_fn) in cases.items():
        p = make_param(**kw)
        model = f2d.Model(p)
        if mask_fn:
            mask_fn(model)
            model.mesh.finalize()
        mesh = model.mesh
        out[f"{name}/meta"] = np.array(json.dumps(kw))
        out[f"{name}/msk"] = mesh.msk.copy()
        for loc, solver in (("c", mesh.poisson_centers), ("v", mesh.poisson_vertices),
                            ("h", mesh.qg_helmholtz)):
            fluid = solver.G > -1
            b = rng.standard_normal(mesh.shape) * fluid
            if loc == "c":
                # compatible RHS: zero sum (the Neumann operator is singular)
                b[fluid] -= b[fluid].mean()
            x = np.zeros(mesh.shape)
            solver.solve(b, x)
            out[f"{name}/{loc}/b"] = b
            out[f"{name}/{loc}/x"] = x
            out[f"{name}/{loc}/G"] = solver.G.copy()
            # the operator itself, as applied to a random vector (pins the matrix)
            v = rng.standard_normal(mesh.shape) * fluid
            Av = np.zeros(mesh.shape)
            Av[fluid] = solver.A @ v[fluid]
            out[f"{name}/{loc}/v"] = v
            out[f"{name}/{loc}/Av"] = Av
    path = os.path.join(HERE, "solve_poisson.npz")
    np.savez_compressed(path, **out)
    print(f"solve_poisson -> {os.path.getsize(path)/1e3:.0f} kB")


def mesh_vectors():
    """Order / mask arrays of Mesh.finalize for several domains."""
    out = {}
    cases = {
        "closed": (dict(nx=30, ny=22), None),
        "xper": (dict(nx=30, ny=22, xperiodic=True), None),
        "yper_quirk": (dict(nx=24, ny=20, yperiodic=True), None),
        "noslip_all": (dict(nx=30, ny=22, noslip=True), island_mask),
        "noslip_walls": (dict(nx=30, ny=22, noslip=["left", "top"], xperiodic=False), island_mask),
        "maxorder4": (dict(nx=30, ny=22, maxorder=4), island_mask),
    }
    for name, (kw, mask_fn) in cases.items():
        p = make_param(**kw)
        model = f2d.Model(p)
        if mask_fn:
            mask_fn(model)
            model.mesh.finalize()
        m = model.mesh
        out[f"{name}/meta"] = np.array(json.dumps(kw))
        for k, v in dict(msk=m.msk, mskx=m.mskx, msky=m.msky, mskv=m.mskv,
                         slipcoef=np.asarray(m.slipcoef, dtype=np.float64),
                         ocx=m.oc.x, ocy=m.oc.y, ovx=m.ov.x, ovy=m.ov.y,
                         okx=m.ok.x, oky=m.ok.y).items():
            out[f"{name}/{k}"] = v.copy()
    path = os.path.join(HERE, "mesh_orders.npz")
    np.savez_compressed(path, **out)
    print(f"mesh_orders -> {os.path.getsize(path)/1e3:.0f} kB")


if __name__ == "__main__":
    which = sys.argv[1:] or None
    todo = [case_euler40, case_vortex, case_vortex_triangle, case_disc_island, case_euler_lake, case_euler_two_basins, case_xper_noslip,
            case_euler_enrk3_upwind, case_euler_centered_ef, case_euler_cweno,
            case_rsw, case_rsw_islands, case_qgrsw_topo, case_qgrsw_islands,
            case_warm_bubble, case_lock_exchange, case_advection, case_advection_disc_upwind, case_eulerpsi,
            case_qg, case_vectoradv, case_euler_lfra, case_rsw_lfra,
            case_euler_tracer, case_lock_exchange_tracer, case_rsw_tracer, run_to_tend, ops_vectors, solve_vectors, mesh_vectors]
    for fn in todo:
        if which is None or fn.__name__ in which:
            fn()
