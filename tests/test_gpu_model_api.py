"""The reference's Python surface, driven the way its experiment scripts and its
own tests do (tests/test_models.py, src/experiments/vortex.py), on the GPU."""
import json
import os

import numpy as np
import pytest

from util import GOLDEN, Golden, rel_l2, remove_component_means, set_state

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def f2d():
    import fluids2d_b200 as f2d
    f2d.Param._quiet = True
    return f2d


def gaussian(x, y, x0, y0, r):
    return np.exp(-((x - x0) ** 2 + (y - y0) ** 2) / (2 * r ** 2))


def set_initial_dipole(f2d, model, x0=None, y0=0.5, r0=0.05, d=0.05):
    """what src/experiments/vortex.py:8-42 does for the euler model"""
    x0 = model.param.Lx / 2 if x0 is None else x0
    x, y = model.mesh.xy("v")
    omega, u = model.state.omega, model.state.u
    omega[:, :] = gaussian(x, y, x0 + d, y0, r0) - gaussian(x, y, x0 - d, y0, r0)
    omega *= model.mesh.mskv * model.mesh.area
    f2d.tools.set_uv_from_omega(model, omega, u)
    model.integrator.diag(model.state)


def test_reference_euler_test_runs_as_is(f2d, capsys):
    """tests/test_models.py:9-15 through Model.run(): same iteration count and
    final fields as the live reference (22 iterations with cfl = 0.9)."""
    z = np.load(os.path.join(GOLDEN, "run_euler40.npz"))
    meta = json.loads(str(z["meta"]))
    p = f2d.Param()
    p.animation = False
    p.tend = 10
    model = f2d.Model(p)
    set_initial_dipole(f2d, model)
    # the initial condition itself (device Poisson solve + host perpgrad) matches
    for k in ("u.x", "u.y", "omega", "ke"):
        n, c = (k.split(".") + [None])[:2]
        a = getattr(model.state, n)
        a = getattr(a, c) if c else a
        assert rel_l2(a, z[f"init/{k}"]) < 1e-11, k
    model.run()
    assert model.time.ite == meta["ite"] == 22
    assert abs(model.time.t - meta["t"]) < 1e-9
    for k, w in (("u.x", model.mesh.mskx), ("u.y", model.mesh.msky), ("omega", model.mesh.mskv)):
        n, c = (k.split(".") + [None])[:2]
        a = getattr(model.state, n)
        a = getattr(a, c) if c else a
        assert rel_l2(a, z[f"final/{k}"], w) < 1e-10, k
    assert "s/dof" in capsys.readouterr().out       # print_perf, model.py:89-93


def test_vortex_script_flow_with_mask_and_finalize(f2d):
    """vortex.py with set_mask(): edit mesh.msk in place, finalize(), dipole, steps"""
    g = Golden("vortex_triangle")
    p = f2d.Param()
    for k, v in g.param.items():
        setattr(p, k, v)
    model = f2d.Model(p)
    x, y = model.mesh.xy()
    model.mesh.msk[y < 0.2 - 0.5 * np.abs(x - p.Lx / 2)] = 0
    model.mesh.finalize()
    assert np.array_equal(model.mesh.msk, g.msk)
    set_initial_dipole(f2d, model, x0=1.0)
    u0 = model.state.u.x            # scripts keep references across run()
    for _ in range(g.nsteps):
        model.set_dt()
        model.step(1)
    assert model.state.u.x is u0
    fin = g.fields("final")
    assert rel_l2(model.state.u.x, fin["u.x"], model.mesh.mskx) < 1e-10
    assert rel_l2(model.state.u.y, fin["u.y"], model.mesh.msky) < 1e-10
    assert rel_l2(model.state.omega, fin["omega"], model.mesh.mskv) < 1e-10
    pm = remove_component_means(model.state.p, model.mesh.msk)
    pr = remove_component_means(fin["p"], model.mesh.msk)
    assert rel_l2(pm, pr, model.mesh.msk) < 1e-9


def test_multi_step_resident_equals_single_steps(f2d):
    g = Golden("vortex")
    def fresh():
        p = f2d.Param()
        for k, v in g.param.items():
            setattr(p, k, v)
        m = f2d.Model(p)
        set_state(m.state, {k: v for k, v in g.fields("init").items()})
        return m
    a, b = fresh(), fresh()
    a.step(4)                       # resident: one upload, 4 fused steps, one download
    for _ in range(4):
        b.set_dt()
        b.step(1)                   # per-step host round trip
    assert a.time.ite == b.time.ite == 4
    for f in ("x", "y"):
        assert np.array_equal(getattr(a.state.u, f), getattr(b.state.u, f))
    assert np.array_equal(a.state.omega, b.state.omega)


def test_forcing_callback_sees_host_state(f2d):
    """model.add_forcing (model.py:121-123): the callback mutates ds on the host
    every stage; a zero forcing must reproduce the unforced run (to solver
    tolerance: the stage-by-stage path has no first-guess history), a non-zero
    one must change it."""
    g = Golden("warm_bubble")
    def fresh():
        p = f2d.Param()
        for k, v in g.param.items():
            setattr(p, k, v)
        p.dt = g.dts[0]
        m = f2d.Model(p)
        set_state(m.state, g.fields("init"))
        return m
    calls = []
    def zero_forcing(param, mesh, s, ds):
        calls.append(float(np.abs(s.b).max()))
        ds.b[0] += 0.0
    def heat(param, mesh, s, ds):
        ds.b[mesh.msk == 1] += 1e-3
    a, b, c = fresh(), fresh(), fresh()
    b.add_forcing(zero_forcing)
    c.add_forcing(heat)
    for m in (a, b, c):
        m.set_dt()
        m.step(2)
    assert len(calls) == 6 and calls[0] > 0
    assert rel_l2(b.state.b, a.state.b) < 1e-11 and rel_l2(b.state.u.x, a.state.u.x) < 1e-10
    assert rel_l2(c.state.b, a.state.b) > 1e-6


def test_integrator_callables_and_scratch(f2d):
    """integrator.rhs / .diag / .scratch are usable from scripts (tracer_advection.py:54,
    vortex.py:42)"""
    g = Golden("euler40")
    p = f2d.Param()
    m = f2d.Model(p)
    set_state(m.state, g.fields("init"))
    ds = m.integrator.scratch[0]
    m.integrator.rhs(m.state, ds)
    assert np.abs(ds.u.x).max() > 0
    from oracle import fluids2d_oracle as orc
    om = orc.Model(orc.make_param(), msk=m.mesh.msk.copy())
    set_state(om.state, g.fields("init"))
    om.rhs(om.state, om.scratch[0])
    assert rel_l2(ds.u.x, om.scratch[0].u.x) < 1e-12
    assert rel_l2(ds.u.y, om.scratch[0].u.y) < 1e-12


def test_poisson_objects_on_mesh(f2d):
    p = f2d.Param()
    p.model = "qgrsw"
    p.nx, p.ny = 48, 40
    m = f2d.Model(p)
    for name in ("poisson_centers", "poisson_vertices", "qg_helmholtz"):
        S = getattr(m.mesh, name)
        b = S.get_rhs("basic")
        x = np.zeros(m.mesh.shape)
        assert S.solve(b, x) is None          # elliptic.py:87 returns mesh.fill(x) == None
        assert np.abs(x).max() > 0 and S.last[1] <= 1e-12
        assert S.G.max() + 1 == int((S.G > -1).sum())
    assert m.mesh.hb == 0 and m.mesh.qgcoef == p.f0 / p.H


def test_install_as_fluids2d_alias(f2d):
    f2d.install_as_fluids2d()
    import fluids2d
    from fluids2d.integrators import copyto        # tracer_advection.py:4
    from fluids2d.equations import fill            # lockexchange.py:3
    from fluids2d.operators import compute_pv, qg_projection, perpgrad   # geos_adj.py:3
    assert fluids2d.Model is f2d.Model and callable(copyto) and callable(fill)
    import sys
    for k in [k for k in sys.modules if k == "fluids2d" or k.startswith("fluids2d.")]:
        del sys.modules[k]


@pytest.mark.parametrize("case", ["euler_tracer", "rsw_tracer"])
def test_param_tracer_through_model(f2d, case):
    """param.tracer = "<name>" adds a prognostic scalar of that name to the state and
    the integrator scratch (states.py:23-34) advected by s.U (equations.py:217-226);
    stepped both per step through host buffers and resident on the device."""
    g = Golden(case)
    name = g.param["tracer"]
    for resident in (False, True):
        p = f2d.Param()
        for k, v in g.param.items():
            setattr(p, k, v)
        model = f2d.Model(p)
        model.mesh.msk[:] = g.msk
        model.mesh.finalize()
        assert model.state._fields[:len(model.integrator.scratch[0]._fields)] == model.integrator.scratch[0]._fields
        assert name in model.integrator.scratch[0]._fields
        set_state(model.state, g.fields("init"))
        if resident:
            model.integrator.upload(model.state)
            for dt in g.dts:
                model.integrator.step_resident(dt, 1)
            model.integrator.download(model.state)
        else:
            for dt in g.dts:
                model.set_dt()
                assert abs(model.time.dt - dt) <= 1e-12 * dt
                model.step(1)
        ref = g.fields("final")[name]
        assert rel_l2(getattr(model.state, name), ref, model.mesh.msk) <= 1e-10, (case, resident)
