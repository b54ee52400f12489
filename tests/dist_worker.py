"""Worker of the multi-GPU parity test (launched by torch.distributed.run, one
process per GPU).  Every rank steps its y-slab of ONE grid; rank 0 also steps
the whole grid on its own GPU; the gathered slab result must match it.

    python -m torch.distributed.run --nproc-per-node 2 tests/dist_worker.py <case>
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def gaussian(x, y, x0, y0, r):
    return np.exp(-((x - x0) ** 2 + (y - y0) ** 2) / (2 * r ** 2))


def build(f2d, case, rank, nranks, device):
    p = f2d.Param()
    p.model = case.get("model", "euler")
    p.nx, p.ny = case["nx"], case["ny"]
    p.Lx, p.Ly = case.get("Lx", 1.0), case.get("Ly", 1.0)
    p.xperiodic = case.get("xperiodic", False)
    p.noslip = case.get("noslip", None)
    p.dt = case["dt"]
    if "f0" in case:
        p.f0 = case["f0"]
    if "dtmax" in case:
        p.dtmax = case["dtmax"]
    p.device = device
    p.rank, p.nranks = rank, nranks
    m = f2d.Model(p)
    x, y = m.mesh.xy()
    if case.get("islands"):
        m.mesh.msk[(x - 0.3 * p.Lx) ** 2 + (y - 0.52 * p.Ly) ** 2 < 0.07 ** 2] = 0   # straddles the slab interface
        m.mesh.msk[(x - 0.75 * p.Lx) ** 2 + (y - 0.2 * p.Ly) ** 2 < 0.05 ** 2] = 0
        m.mesh.finalize()
    xv, yv = m.mesh.xy("v")
    s = m.state
    if p.model == "qgrsw":
        # rsw_with_topo.py:12-53, 96-101: thickness dipole over a bump, balanced by the QG projection
        m.mesh.hb = 0.1 * gaussian(x, y, 0.35 * p.Lx, 0.7 * p.Ly, 0.06) * m.mesh.area * m.mesh.msk
        s.h[...] = (p.H + 0.2 * (gaussian(x, y, 0.6 * p.Lx, 0.5 * p.Ly, 0.1) - gaussian(x, y, 0.4 * p.Lx, 0.5 * p.Ly, 0.1))) \
            * m.mesh.msk * m.mesh.area - m.mesh.hb
        f2d.operators.qg_projection(m.mesh, s.u, s.h, s.pv, s.psi)
        m.integrator.diag(s)
        return m
    if p.model == "rsw":
        # geostrophic adjustment of a thickness dipole over a Gaussian bump (geos_adj.py:12-49,
        # rsw_with_topo.py:96-99): the flow starts at rest
        m.mesh.hb = 0.1 * gaussian(x, y, 0.35 * p.Lx, 0.7 * p.Ly, 0.06) * m.mesh.area * m.mesh.msk
        s.h[...] = (p.H + 0.2 * (gaussian(x, y, 0.6 * p.Lx, 0.5 * p.Ly, 0.1) - gaussian(x, y, 0.4 * p.Lx, 0.5 * p.Ly, 0.1))) \
            * m.mesh.msk * m.mesh.area - m.mesh.hb
        m.integrator.diag(s)
        return m
    if case.get("turbulence"):
        import bench
        s.omega[...] = bench.turbulence_vorticity(m.mesh.x("v"), m.mesh.y("v"), m.mesh.area) * m.mesh.mskv
        f2d.tools.set_uv_from_omega(m, s.omega, s.u)
        s.u.x[...] *= 20.0      # physical speed ~ 1 (bench.py normalises by the measured maximum; a constant
        s.u.y[...] *= 20.0      # keeps every rank and the single-GPU reference identical): CFL dt ~ 0.5 / nx
    else:
        s.omega[...] = (gaussian(xv, yv, 0.55 * p.Lx, 0.5 * p.Ly, 0.06) - gaussian(xv, yv, 0.45 * p.Lx, 0.5 * p.Ly, 0.06)) \
            * m.mesh.mskv * m.mesh.area
        f2d.tools.set_uv_from_omega(m, s.omega, s.u)
    if p.model == "boussinesq":
        s.b[...] = (y + 0.1 * gaussian(x, y, 0.5 * p.Lx, 0.45 * p.Ly, 0.08)) * m.mesh.msk
    m.integrator.diag(s)
    return m


def main():
    import torch
    import torch.distributed as dist
    case = json.loads(sys.argv[1])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    import fluids2d_b200 as f2d
    from fluids2d_b200 import slabs
    f2d.Param._quiet = True
    slabs.init_from_torch_distributed()

    m = build(f2d, case, rank, world, local)
    for _ in range(case["steps"]):
        m.set_dt()
        m.step(1)
    shape_g = (case["ny"] + 6, case["nx"] + 6)
    names = {"qgrsw": ["u.x", "u.y", "omega", "h", "psi"]}.get(case.get("model"), ["u.x", "u.y", "omega", "ke", "p"]) \
        + {"boussinesq": ["b"], "rsw": ["h"]}.get(case.get("model"), [])
    got = {}
    for n in names:
        a = getattr(m.state, n.split(".")[0])
        a = getattr(a, n.split(".")[1]) if "." in n else a
        got[n] = slabs.gather_global(m.mesh.slab, a, shape_g)
    msk_g = slabs.gather_global(m.mesh.slab, m.mesh.msk, shape_g)
    stats = m.mesh.engine.solver_stats()
    # the six all-reduced sums behind diagnostics.Bulk (owned rows of every slab)
    has_bulk = case.get("model") != "qgrsw"         # (its diagnostics do not form ke)
    if has_bulk:
        m.integrator.upload(m.state, ["ke", "omega", "U.x", "U.y", "u.x", "u.y"])
        bulk = m.mesh.engine.bulk_sums()
    out = {"rank": rank, "exchanges": m.mesh.engine.exchange_count(), "solver": stats}
    if rank == 0:
        from util import rel_l2, remove_component_means
        ref = build(f2d, case, 0, 1, local)         # the same grid on ONE GPU
        for _ in range(case["steps"]):
            ref.set_dt()
            ref.step(1)
        assert np.array_equal(msk_g, ref.mesh.msk)
        errs = {}
        for n in names:
            a = getattr(ref.state, n.split(".")[0])
            a = getattr(a, n.split(".")[1]) if "." in n else a
            g = got[n]
            if n == "p" and case.get("model") != "rsw":       # (the rsw pressure is g (h + hb), not a solve)
                g, a = remove_component_means(g, ref.mesh.msk), remove_component_means(a, ref.mesh.msk)
            w = {"u.x": ref.mesh.mskx, "u.y": ref.mesh.msky, "omega": ref.mesh.mskv, "psi": ref.mesh.mskv}.get(n, ref.mesh.msk)
            errs[n] = rel_l2(g, a, w)
        if not has_bulk:
            out["errors"] = errs
            out["ref_solver"] = ref.mesh.engine.solver_stats()
            print("DIST_RESULT " + json.dumps(out), flush=True)
            dist.barrier()
            dist.destroy_process_group()
            return
        ref.integrator.upload(ref.state, ["ke", "omega", "U.x", "U.y", "u.x", "u.y"])
        bulk_ref = ref.mesh.engine.bulk_sums()
        # [sum ke, sum omega^2, sum omega, sum U.y xv, sum U.x yu, sum msk]: the signed sums may cancel to ~0,
        # so they are measured against the scale of their terms
        d = np.abs(bulk - bulk_ref)
        scale = np.array([abs(bulk_ref[0]), abs(bulk_ref[1]), np.sqrt(bulk_ref[1] * bulk_ref[5]),
                          max(abs(bulk_ref[3]), abs(bulk_ref[4])) + 1e3 * np.sqrt(bulk_ref[0] * bulk_ref[5]),
                          max(abs(bulk_ref[3]), abs(bulk_ref[4])) + 1e3 * np.sqrt(bulk_ref[0] * bulk_ref[5]),
                          bulk_ref[5]]) + 1e-300
        errs["bulk_sums"] = float(np.max(d / scale))
        out["errors"] = errs
        out["ref_solver"] = ref.mesh.engine.solver_stats()
        print("DIST_RESULT " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
