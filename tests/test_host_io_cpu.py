"""History / bulk files written by the host side (reference: io.py:5-93,
diagnostics.py:8-107): layout, float32 storage, record appends.  CPU only: the
host-array branch, no device calls."""
import os
from types import SimpleNamespace

import numpy as np
import pytest


@pytest.fixture
def env(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    import fluids2d_b200 as f2d
    import fluids2d_b200.states as st
    monkeypatch.setattr(st, "_pinned", False)
    f2d.Param._quiet = True
    p = f2d.Param()
    shape = (p.ny + 6, p.nx + 6)
    x1, y1 = (np.arange(shape[1]) - 2.5) / p.nx, (np.arange(shape[0]) - 2.5) / p.ny
    mesh = SimpleNamespace(shape=shape, slab=None, msk=np.ones(shape, "i1"), dx=1 / p.nx, dy=1 / p.ny,
                           area=1 / (p.nx * p.ny), xy=lambda which="c": np.meshgrid(x1, y1))
    return f2d, p, mesh, st.State(p, shape)


def test_history_file_layout_and_records(env):
    f2d, p, mesh, s = env
    from fluids2d_b200 import _nc
    from fluids2d_b200.io import IO
    p.nhis, p.var_to_store = 2, ["u", "omega"]
    rng = np.random.default_rng(0)
    s.u.x[:] = rng.standard_normal(mesh.shape)
    s.omega[:] = rng.standard_normal(mesh.shape)
    t = SimpleNamespace(t=0.5, ite=3, dt=0.1)
    io = IO(p, mesh, s, t)
    io.write(s, t)
    t.t, t.ite = 0.75, 5
    s.omega[:] *= 2
    io.write(s, t)
    assert io.kt == 2 and _nc.number_of_records(p.outputfile) == 2
    with _nc.Dataset(p.outputfile, "r") as nc:
        v = nc.variables
        assert set(["xc", "yc", "t", "ite", "dt", "ux", "uy", "omega"]) <= set(v.keys())   # io.py:63-88
        assert v["ux"][:].shape == (2,) + mesh.shape
        assert np.dtype(v["omega"][:].dtype).itemsize == 4                               # dtype "f", io.py:65
        assert np.array_equal(np.asarray(v["omega"][1]), s.omega.astype(np.float32))
        assert np.array_equal(np.asarray(v["omega"][0]), (s.omega / 2).astype(np.float32))
        assert list(np.asarray(v["ite"][:])) == [3, 5]
        assert np.allclose(np.asarray(v["t"][:]), [0.5, 0.75])
        assert np.array_equal(np.asarray(v["xc"][:]), mesh.xy()[0].astype(np.float32))


def test_bulk_matches_reference_expressions(env):
    f2d, p, mesh, s = env
    from fluids2d_b200.diagnostics import Bulk
    rng = np.random.default_rng(1)
    for a in (s.ke, s.omega, s.U.x, s.U.y):
        a[:] = rng.standard_normal(mesh.shape)
    t = SimpleNamespace(t=0.0, ite=0, dt=0.1)
    model = SimpleNamespace(param=p, mesh=mesh, state=s, time=t)
    b = Bulk(model)
    for ite in range(7):
        t.ite, t.t = ite, 0.1 * ite
        b()
    assert b.kt == 3                                  # every third iteration, diagnostics.py:40-41
    # diagnostics.py:52-56 written out
    n = mesh.msk.sum()
    xv, yu = mesh.xy("y")[0], mesh.xy("x")[1]
    assert b.data.ke[0] == s.ke.sum() / n
    assert b.data.ens[0] == 0.5 * ((s.omega ** 2).sum() / n) / mesh.area ** 2
    assert b.data.vort[0] == (s.omega.sum() / n) / mesh.area
    assert b.data.angular[0] == ((s.U.y * xv).sum() / n) * mesh.dx - ((s.U.x * yu).sum() / n) * mesh.dy
    b.finalize()
    assert b.kt == 0 and b.k0 == 3
    got = b.read()
    assert np.allclose(got.time, [0.0, 0.3, 0.6]) and len(got.ke) == 3
    # a second Bulk on the same file appends (diagnostics.py:31-34)
    b2 = Bulk(model)
    assert b2.k0 == 3
