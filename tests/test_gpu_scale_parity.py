"""Parity of the CUDA path against the oracle on BASELINE.json's configurations at
the largest sizes the CPU oracle reaches in seconds to minutes (512^2 - 1024^2).

At these sizes the path the benchmark times is the one under test: 6-7 multigrid
levels, almost every tile on the open (register) path, CUDA-graph iterations that
run without a host check, the cubic first guess in model time (solver_guess = 4,
the default), fixed AND adaptive time steps.  Ten steps, every field, relative L2
on fluid cells <= 1e-10 (BASELINE.json north_star); `p` after removing its mean
per connected fluid component.

  config 2   euler, x-periodic channel, band-limited random vorticity (bench.py's
             own initial condition), 512^2 and 1024^2
  config 3   rsw, closed basin with four disc islands and a thin peninsula, 512^2
  config 4   qgrsw, same basin + Gaussian topography (Thiry projection: one vertex
             Helmholtz solve per RK stage), 512^2
  config 5   boussinesq, x-periodic, b = y + 0.1 gaussian (warm_bubble.py:14-20), 1024 x 512
             + an enclosed lake at 512^2 (two null-space constants)
The oracle (oracle/fluids2d_oracle.py, numpy + C + SuperLU) is the checker only.
"""
import os
import sys
import time

import numpy as np
import pytest

from util import rel_l2, remove_component_means

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu
NH = 3
TOL = 1e-10


def gaussian(x, y, x0, y0, r):
    return np.exp(-((x - x0) ** 2 + (y - y0) ** 2) / (2 * r ** 2))


@pytest.fixture(scope="module")
def f2d():
    import fluids2d_b200 as f2d
    f2d.Param._quiet = True
    return f2d


def basin_mask(shape, nh=NH):
    """closed basin with four disc islands and a thin peninsula"""
    n2, n1 = shape
    y, x = np.ogrid[0:n2, 0:n1]
    ny, nx = n2 - 2 * nh, n1 - 2 * nh
    msk = np.zeros(shape, np.int8)
    msk[nh:-nh, nh:-nh] = 1
    for (cx, cy, r) in ((0.25, 0.3, 0.06), (0.7, 0.75, 0.08), (0.8, 0.2, 0.05), (0.4, 0.65, 0.03)):
        msk[(x - nh - cx * nx) ** 2 + (y - nh - cy * ny) ** 2 < (r * nx) ** 2] = 0
    msk[nh + ny // 2:nh + ny // 2 + 5, nh:nh + nx // 5] = 0
    return msk


def make_pair(f2d, orc, kw, msk_fn=None, hb_fn=None):
    """the same model on the device (public API) and in the oracle"""
    p = f2d.Param()
    for k, v in kw.items():
        setattr(p, k, v)
    model = f2d.Model(p)
    if msk_fn is not None:
        model.mesh.msk[...] = msk_fn(model.mesh.shape)
        model.mesh.finalize()
    op = orc.make_param(**{k: v for k, v in kw.items() if not k.startswith("solver_")})   # solver knobs are ours only
    om = orc.Model(op, msk=model.mesh.msk.copy())
    if hb_fn is not None:
        hb = hb_fn(model.mesh)
        model.mesh.hb = hb
        om.mesh.hb = hb.copy()
    return model, om


def copy_state(src, dst):
    for name in src._fields:
        a = getattr(src, name)
        b = getattr(dst, name, None)
        if b is None:
            continue
        if hasattr(a, "_fields"):
            b.x[...] = a.x
            b.y[...] = a.y
        else:
            b[...] = a


def field_pair(mesh, model_name, sa, sb, name):
    n, c = (name.split(".") + [None])[:2]
    a, b = getattr(sa, n), getattr(sb, n)
    if c:
        a, b = getattr(a, c), getattr(b, c)
        w = {"x": mesh.mskx, "y": mesh.msky}[c]
    else:
        w = mesh.mskv if n in ("omega", "pv", "psi") else mesh.msk
    if n == "p" and model_name in ("euler", "boussinesq"):
        a, b = remove_component_means(a, mesh.msk), remove_component_means(b, mesh.msk)
    return a, b, w


def run_and_compare(model, om, nsteps, fields, adaptive, label, om_refined=None):
    """10 steps on both sides with identical dt (the device decides an adaptive one);
    returns the worst relative L2 per field.  om_refined: a second oracle whose direct
    solves get one step of iterative refinement -- its distance to the plain oracle is the
    reference's own forward-error floor, returned as `floor`."""
    mesh, s, o = model.mesh, model.state, om.state
    copy_state(s, o)
    t0 = time.time()
    eng = mesh.engine
    eng.solver_stats()
    dts, per_step = [], []
    for _ in range(nsteps):
        if adaptive:
            model.set_dt()
        dts.append(model.time.dt)
        model.step(1)
        per_step.append(eng.solver_stats())
    st = dict(nsolves=sum(x["nsolves"] for x in per_step), niters=sum(x["niters"] for x in per_step),
              max_relres=max(x["max_relres"] for x in per_step),
              last3=sum(x["niters"] for x in per_step[-3:]) / max(sum(x["nsolves"] for x in per_step[-3:]), 1))
    t1 = time.time()
    for dt in dts:
        om.step(dt)
    t2 = time.time()
    worst = {}
    for name in fields:
        a, b, w = field_pair(mesh, model.param.model, s, o, name)
        assert np.all(np.isfinite(a[np.asarray(w) != 0])), (label, name)
        worst[name] = rel_l2(a, b, w)
    print(f"{label}: device {t1 - t0:.1f}s oracle {t2 - t1:.1f}s dt {dts[0]:.3e}..{dts[-1]:.3e} solver {st} "
          + " ".join(f"{k}={v:.1e}" for k, v in worst.items()))
    if om_refined is not None:
        for dt in dts:
            om_refined.step(dt)
        floor = {}
        for name in fields:
            a, b, w = field_pair(mesh, model.param.model, om_refined.state, o, name)
            floor[name] = rel_l2(a, b, w)
        print(f"{label}: reference's own forward-error floor (LU vs LU + 1 refinement step) "
              + " ".join(f"{k}={v:.1e}" for k, v in floor.items()))
        st["floor"] = floor
    return worst, st, dts


def euler_turbulence_ic(f2d, model):
    import bench
    mesh, s = model.mesh, model.state
    s.omega[...] = bench.turbulence_vorticity(mesh.x("v"), mesh.y("v"), mesh.area) * mesh.mskv
    f2d.tools.set_uv_from_omega(model, s.omega, s.u)
    umax = max(np.abs(s.u.x).max() / mesh.dx, np.abs(s.u.y).max() / mesh.dy)
    s.u.x[...] *= 1.0 / umax          # physical speed ~ 1 -> CFL dt ~ 1 / n, as bench.py does
    s.u.y[...] *= 1.0 / umax
    model.integrator.diag(s)


@pytest.mark.parametrize("n,adaptive", [(512, False), (512, True), (1024, False)])
def test_config2_euler_channel(f2d, oracle, n, adaptive):
    """the headline workload of bench.py at oracle-reachable sizes"""
    import bench
    kw = dict(model="euler", nx=n, ny=n, xperiodic=True, cfl=0.9, maxorder=6,
              integrator="rk3", vortexforce="weno", innerproduct="weno", compflux="weno")
    model, om = make_pair(f2d, oracle, kw)
    euler_turbulence_ic(f2d, model)
    if not adaptive:
        model.set_dt()
        model.param.dt = model.time.dt
    info = model.mesh.engine.solver_info("c")
    assert info["levels"] >= 7 and info["components"] == 1, info
    worst, st, dts = run_and_compare(model, om, 10, ["u.x", "u.y", "omega", "ke", "p", "U.x", "U.y"], adaptive,
                                     f"config2 euler {n}^2 {'adaptive' if adaptive else 'fixed'} dt")
    if adaptive:
        assert len(set(dts)) > 1            # the step really changed from step to step
    for k, v in worst.items():
        assert v <= TOL, (k, v)
    assert st["max_relres"] <= 1e-12
    # from a cold start the first solves take ~12 iterations; once the first-guess history of each
    # RK stage is four steps deep they take 4-5: measured 7.7 / 7.2 per solve over
    # the ten steps at 512^2 / 1024^2 and 7.0 / 6.2 over the last three (dt ~ 1/n is larger on these
    # grids than at 4096^2, where the extrapolated guess leaves 4.0)
    assert st["niters"] <= 8.5 * st["nsolves"] and st["last3"] <= 7.5, st
    model.mesh.engine.close()


def rsw_dipole_ic(model, orc_model=None, sub_hb=False):
    """geos_adj.py:12-49 (flow='dipole'); rsw_with_topo.py:12-53 when sub_hb"""
    mesh, s, p = model.mesh, model.state, model.param
    x, y = mesh.xy("c")
    s.h[...] = p.H + 0.2 * (gaussian(x, y, 0.6, 0.5, 0.1) - gaussian(x, y, 0.4, 0.5, 0.1))
    s.h[...] *= mesh.msk * mesh.area
    if sub_hb:
        s.h[...] -= mesh.hb


def test_config3_rsw_basin_with_islands(f2d, oracle):
    n = 512
    kw = dict(model="rsw", nx=n, ny=n, f0=10.0, dtmax=1.0, noslip=True)
    model, om = make_pair(f2d, oracle, kw, msk_fn=basin_mask)
    rsw_dipole_ic(model)
    model.integrator.diag(model.state)
    worst, st, dts = run_and_compare(model, om, 10, ["u.x", "u.y", "h", "omega", "ke", "p"], True, "config3 rsw 512^2 islands")
    for k, v in worst.items():
        assert v <= TOL, (k, v)
    model.mesh.engine.close()


def test_config4_qgrsw_basin_topography(f2d, oracle):
    n = 512
    kw = dict(model="qgrsw", nx=n, ny=n, f0=10.0, dtmax=1.0)

    def hb(mesh):   # rsw_with_topo.py:96-99
        x, y = mesh.xy()
        return 0.2 * gaussian(x, y, 0.3, 0.7, 0.05) * mesh.area * mesh.msk

    model, om = make_pair(f2d, oracle, kw, msk_fn=basin_mask, hb_fn=hb)
    rsw_dipole_ic(model, sub_hb=True)
    s = model.state
    f2d.operators.qg_projection(model.mesh, s.u, s.h, s.pv, s.psi)     # rsw_with_topo.py:49-50
    model.integrator.diag(s)
    info = model.mesh.engine.solver_info("h")
    assert info["levels"] >= 6, info
    worst, st, dts = run_and_compare(model, om, 10, ["u.x", "u.y", "h", "omega", "pv", "psi"], True,
                                     "config4 qgrsw 512^2 islands + topography")
    for k, v in worst.items():
        assert v <= TOL, (k, v)
    assert st["max_relres"] <= 1e-12 and st["nsolves"] == 30
    model.mesh.engine.close()


def test_config5_boussinesq_channel(f2d, oracle):
    """A stratified fluid at rest: the pressure has to cancel the O(1) buoyancy b = y while the
    flow the warm bubble drives is small, and the smooth pressure modes that do the cancelling
    are the ones a Poisson solve determines worst (forward error cond(A) eps ~ 1e6 x 1e-16).
    The reference's own answer moves by MORE than 1e-10 in omega when its direct solve is given
    one step of iterative refinement (a backward-stable perturbation), so for this configuration
    the tolerance is the north_star's 1e-10 or 3x that floor, whichever is larger; the floor is
    measured in the test (oracle twice).  Tightening the device solver (rtol 1e-12 -> 1e-14)
    does not change the distance: it is not the device's residual."""
    kw = dict(model="boussinesq", nx=1024, ny=512, Lx=2.0, Ly=1.0, xperiodic=True, cfl=0.9, dtmax=1e-1)
    model, om = make_pair(f2d, oracle, kw)
    om2 = oracle.Model(oracle.make_param(**kw), msk=model.mesh.msk.copy())
    om2.mesh.poisson_centers.refine = 1
    mesh, s = model.mesh, model.state
    x, y = mesh.xy()
    s.b[...] = (y + 0.1 * gaussian(x, y, 1.0, 0.25, 0.08)) * mesh.msk      # warm_bubble.py:14-20
    model.integrator.diag(s)
    copy_state(s, om2.state)
    worst, st, dts = run_and_compare(model, om, 10, ["b", "u.x", "u.y", "omega", "ke", "p"], True,
                                     "config5 boussinesq 1024x512", om_refined=om2)
    for k, v in worst.items():
        assert v <= max(TOL, 3 * st["floor"][k]), (k, v, st["floor"][k])
    assert st["floor"]["omega"] > TOL          # documents why omega cannot be pinned tighter here
    assert worst["b"] <= TOL and worst["u.x"] <= TOL and worst["u.y"] <= TOL
    assert st["max_relres"] <= 1e-12
    model.mesh.engine.close()


def lake_mask(shape, nh=NH):
    """closed box, a wall ring enclosing a lake, plus a land bar that cuts off the north-east corner"""
    n2, n1 = shape
    y, x = np.ogrid[0:n2, 0:n1]
    ny, nx = n2 - 2 * nh, n1 - 2 * nh
    msk = np.zeros(shape, np.int8)
    msk[nh:-nh, nh:-nh] = 1
    r = np.sqrt((x - nh - 0.62 * nx) ** 2 + (y - nh - 0.5 * ny) ** 2)
    msk[(r > 0.15 * nx) & (r < 0.18 * nx)] = 0
    msk[(x - nh) + (y - nh) > 1.72 * nx] = 0
    msk[(np.abs((x - nh) + (y - nh) - 1.6 * nx) < 0.02 * nx) & (msk == 1)] = 0
    return msk


def test_euler_enclosed_lake_512(f2d, oracle):
    """several connected fluid components: the reference's direct solve gives each its own
    null-space constant (elliptic.py:186-190); the PCG projects per component"""
    from scipy import ndimage
    n = 512
    kw = dict(model="euler", nx=n, ny=n, noslip=False)
    model, om = make_pair(f2d, oracle, kw, msk_fn=lake_mask)
    mesh, s = model.mesh, model.state
    ncomp = ndimage.label(mesh.msk != 0)[1]
    info = mesh.engine.solver_info("c")
    assert ncomp >= 2 and info["components"] == ncomp, (ncomp, info)
    xv, yv = mesh.xy("v")
    s.omega[...] = (gaussian(xv, yv, 0.25, 0.3, 0.05) - gaussian(xv, yv, 0.25, 0.42, 0.05)
                    + 0.8 * gaussian(xv, yv, 0.6, 0.52, 0.04)) * mesh.mskv * mesh.area
    f2d.tools.set_uv_from_omega(model, s.omega, s.u)
    model.integrator.diag(s)       # projects a velocity that is divergence-free already: the rhs is rounding noise
    mesh.engine.solver_info("c")   # ... so start the compatibility record here
    worst, st, dts = run_and_compare(model, om, 10, ["u.x", "u.y", "omega", "ke", "p"], True, "euler 512^2 lake")
    for k, v in worst.items():
        assert v <= TOL, (k, v)
    assert st["max_relres"] <= 1e-12 and st["niters"] <= 12 * st["nsolves"], st
    assert mesh.engine.solver_info("c")["rhs_incompat"] < 1e-10      # div(U) sums to zero on every closed component
    mesh.engine.close()
