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


def test_device_observers_history_and_bulk(f2d, tmp_path, monkeypatch):
    """Model.run() with nhis > 0 and diagnostics.Bulk: the history records are
    float32 conversions done on the device + copy stream (f2d_download_f32) and
    the bulk averages come from one device reduction (f2d_bulk_sums); both must
    equal what the reference's host expressions give on the same states."""
    monkeypatch.chdir(tmp_path)
    from fluids2d_b200 import _nc
    from fluids2d_b200.diagnostics import Bulk

    def make(nhis):
        p = f2d.Param()
        p.nx, p.ny, p.dt, p.tend, p.maxite = 64, 48, 0.05, 1e9, 9
        p.nhis, p.var_to_store = nhis, ["u", "omega", "p"]
        m = f2d.Model(p)
        x, y = m.mesh.xy()
        m.mesh.msk[(x - 0.7) ** 2 + (y - 0.3) ** 2 < 0.1 ** 2] = 0
        m.mesh.finalize()
        set_initial_dipole(f2d, m, x0=0.4)
        return m

    # device path
    m = make(3)
    bulk = Bulk(m, ncfile="bulk_dev.nc")
    m.diags.append(bulk)
    assert not m._observation_due()                 # nothing here needs the host state
    m.run()
    assert m.time.ite == 9 and m.io.kt == 4          # records at ite 0, 3, 6, 9
    dev = bulk.read()
    # host path: same steps one by one through host buffers, observers fed host arrays
    h = make(0)
    hb = Bulk(h, ncfile="bulk_host.nc")
    recs, stamps = [], []
    for it in range(10):
        if it % 3 == 0:
            recs.append({k: np.array(a, dtype=np.float32) for k, a in
                         (("ux", h.state.u.x), ("uy", h.state.u.y), ("omega", h.state.omega), ("p", h.state.p))})
        if it < 9:
            h.set_dt()
            h.step(1)
            hb()                                     # compute_diags follows each step (model.py:50)
    hb.finalize()
    host = hb.read()
    with _nc.Dataset(m.param.outputfile, "r") as nc:
        assert list(np.asarray(nc.variables["ite"][:])) == [0, 3, 6, 9]
        for k in ("ux", "uy", "omega"):
            got = np.asarray(nc.variables[k][:])
            assert got.dtype.itemsize == 4 and got.shape == (4,) + m.mesh.shape
            for r in range(4):
                ref = recs[r][k]
                assert np.abs(got[r] - ref).max() <= 2e-7 * max(np.abs(ref).max(), 1e-30), (k, r)
    assert len(dev.time) == len(host.time) == 3
    for name in ("ke", "ens", "vort", "angular"):
        a, b = np.asarray(getattr(dev, name), float), np.asarray(getattr(host, name), float)
        assert np.allclose(a, b, rtol=1e-6, atol=1e-7 * np.abs(b).max() + 1e-30), name   # float32 file
    # the sums themselves, in double
    m.integrator.upload(m.state)
    s = m.state
    sums = m.mesh.engine.bulk_sums()
    xv, yu = m.mesh.xy("y")[0], m.mesh.xy("x")[1]
    ref = [s.ke.sum(), (s.omega ** 2).sum(), s.omega.sum(), (s.U.y * xv).sum(), (s.U.x * yu).sum(), m.mesh.msk.sum()]
    for k in range(6):
        assert abs(sums[k] - ref[k]) <= 1e-12 * max(abs(ref[k]), np.abs(s.ke).sum()), k


def test_device_forcing_equals_host_callback_and_oracle(f2d):
    """equations.DeviceForcing (ds.<field> += amplitude(t) * pattern, the shape of
    forced_convection.py:9-24) keeps the fused resident step; the same object used
    as a HOST callback (reference granularity) and in the CPU oracle must give the
    same fields.  Covers a scalar (b), the momentum (u.x) and a time-dependent
    amplitude."""
    from types import SimpleNamespace
    from oracle import fluids2d_oracle as orc
    from fluids2d_b200.equations import DeviceForcing
    g = Golden("warm_bubble")
    nsteps, dt = 4, g.dts[0]

    def make_forcing(mesh):
        x, y = mesh.xy()
        Fb = 0.3 * np.exp(-((x - 1.0) ** 2 + (y - 0.2) ** 2) / 0.02) * mesh.msk
        Fu = 0.05 * np.sin(2 * np.pi * y) * mesh.mskx
        return DeviceForcing({"b": Fb, "u": (Fu, np.zeros_like(Fu))}, amplitude=lambda t: 1.0 + 4.0 * t)

    def fresh():
        p = f2d.Param()
        for k, v in g.param.items():
            setattr(p, k, v)
        p.dt = dt
        m = f2d.Model(p)
        set_state(m.state, g.fields("init"))
        return m

    dev, host, free = fresh(), fresh(), fresh()
    dev.add_forcing(make_forcing(dev.mesh))
    assert dev.integrator.rhs is dev.integrator._device_rhs        # still the fused path
    host.mesh.time = host.time                                     # forced_convection.py:6
    fh = make_forcing(host.mesh)
    host.add_forcing(lambda param, mesh, s, ds: fh(param, mesh, s, ds))
    assert host.integrator.rhs is not host.integrator._device_rhs
    # device forcing, resident steps
    dev.integrator.upload(dev.state)
    for k in range(nsteps):
        dev.integrator.step_resident(dt, 1)
        dev.time.pushforward()
    dev.integrator.download(dev.state)
    for m in (host, free):
        m.set_dt()
        m.step(nsteps)
    # oracle with the same callback
    om = orc.Model(orc.make_param(**dict(g.param, dt=dt)), msk=g.msk.copy())
    set_state(om.state, g.fields("init"))
    clock = SimpleNamespace(t=0.0)
    om.mesh.time = clock
    fo = make_forcing(SimpleNamespace(xy=dev.mesh.xy, msk=g.msk, mskx=dev.mesh.mskx))
    plain_rhs = om.rhs

    def forced(s, ds):
        plain_rhs(s, ds)
        fo(om.param, om.mesh, s, ds)
    om.rhs = forced
    for k in range(nsteps):
        clock.t = k * dt
        om.step(dt)
    for name, w in (("b", dev.mesh.msk), ("u.x", dev.mesh.mskx), ("u.y", dev.mesh.msky), ("omega", dev.mesh.mskv)):
        n, c = (name.split(".") + [None])[:2]
        get = lambda s: getattr(getattr(s, n), c) if c else getattr(s, n)
        a, b, o, f = get(dev.state), get(host.state), get(om.state), get(free.state)
        assert rel_l2(a, o, w) <= 1e-10, ("device vs oracle", name, rel_l2(a, o, w))
        assert rel_l2(b, o, w) <= 1e-10, ("host callback vs oracle", name)
        assert rel_l2(a, f, w) > 1e-4, ("the forcing must matter", name)
