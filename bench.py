#!/usr/bin/env python
"""Headline benchmark: grid-point updates/s incl. elliptic solve, 4096^2 Euler.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config euler4096|euler4096dp|vortex|rsw8192|qgrsw8192|bouss16384] [--grid n]

One "step" = one full RK3 step of the model on the configuration's grid: for the
headline (BASELINE config 2, `euler4096`) 3 stages = 3 WENO advection kernels + 3
RK updates + 3 pressure projections, each with a multigrid-PCG elliptic solve +
vorticity/kinetic-energy diagnostics, on a 4096^2 x-periodic channel (the
reference-supported variant of "doubly periodic", SURVEY note Y), fp64, WENO5-Z +
SSP-RK3, fixed dt from CFL 0.9.  The other --config values are BASELINE configs
1, 3, 4, 5; the driver runs the default.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the fields.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grid-point updates/s incl. elliptic solve, 4096^2 Euler"
UNIT = "grid-point updates/s"


# ------------------------------------------------------------------ workload --
def turbulence_vorticity(x, y, area, seed=0, kpeak=8.0, kwidth=3.0, nmodes=96, yperiod=None):
    """band-limited random vorticity on the vertex grid x (1-D, columns) x y (1-D,
    rows): a sum of `nmodes` Fourier modes with numpy default_rng(seed) wave
    vectors (|k| ~ N(kpeak, kwidth), integer kx: x-periodic), phases and
    amplitudes, times the cell area as the reference's `omega` carries it
    (vortex.py:19-20).  Point-wise in (x, y), so every slab of a decomposed grid
    evaluates its own rows; separable, so it is two small matrix products."""
    rng = np.random.default_rng(seed)
    kmag = np.abs(rng.normal(kpeak, kwidth, nmodes))
    theta = rng.uniform(0, 2 * np.pi, nmodes)
    kx = np.rint(kmag * np.cos(theta))
    ky = kmag * np.sin(theta)
    if yperiod:                     # doubly periodic box: integer wave numbers in y too
        ky = np.rint(ky * yperiod) / yperiod
    phase = rng.uniform(0, 2 * np.pi, nmodes)
    amp = rng.normal(0, 1, nmodes) * (area / np.sqrt(nmodes))
    ax = 2 * np.pi * np.outer(x, kx) + phase          # (n1, M)
    by = 2 * np.pi * np.outer(y, ky)                  # (n2, M)
    # cos(a + b) = cos a cos b - sin a sin b
    return (np.cos(by) * amp) @ np.cos(ax).T - (np.sin(by) * amp) @ np.sin(ax).T


def gaussian(x, y, x0, y0, r):
    return np.exp(-((x - x0) ** 2 + (y - y0) ** 2) / (2 * r ** 2))


def basin_mask(shape, nh=3, row0=0, ny=None):
    """closed basin with four disc islands and a thin peninsula;
    `row0` / `ny`: the rows of a slab inside a global grid of `ny` interior rows"""
    n2, n1 = shape
    nx = n1 - 2 * nh
    ny = n2 - 2 * nh if ny is None else ny
    y, x = np.ogrid[row0:row0 + n2, 0:n1]
    msk = ((y >= nh) & (y < nh + ny) & (x >= nh) & (x < nh + nx)).astype(np.int8)
    for (cx, cy, r) in ((0.25, 0.3, 0.06), (0.7, 0.75, 0.08), (0.8, 0.2, 0.05), (0.4, 0.65, 0.03)):
        msk[(x - nh - cx * nx) ** 2 + (y - nh - cy * ny) ** 2 < (r * nx) ** 2] = 0
    msk[(y >= nh + ny // 2) & (y < nh + ny // 2 + 5) & (x >= nh) & (x < nh + nx // 5)] = 0
    return msk


# BASELINE.json configs.  `bytes_pt`: SURVEY 8d's algorithmic bytes per grid point per
# step outside the elliptic solves; `solves`: elliptic solves per step.
CONFIGS = {
    "euler4096": dict(baseline_config=2, model="euler", n=4096, xperiodic=True, bytes_pt=497, solves=3,
                      what="Euler {nx}x{ny} fp64, x-periodic channel, WENO5-Z + SSP-RK3, band-limited random "
                           "vorticity (rng 0), fixed dt = CFL 0.9"),
    "euler4096dp": dict(baseline_config=2, model="euler", n=4096, xperiodic=True, ywrap=True, bytes_pt=497, solves=3,
                        what="Euler {nx}x{ny} fp64, DOUBLY PERIODIC (param.ywrap: a new feature, the reference has no "
                             "periodic y -- SURVEY note Y; oracle pinned by transposition symmetry), WENO5-Z + SSP-RK3, "
                             "band-limited random vorticity (rng 0, made y-periodic), fixed dt = CFL 0.9"),
    "vortex": dict(baseline_config=1, model="euler", n=None, bytes_pt=497, solves=3,
                   what="experiments/vortex.py as shipped: Euler 200x100, Lx = 2, closed free-slip box, dipole, "
                        "fixed dt = 0.2 (vortex.py:59-71)"),
    "rsw8192": dict(baseline_config=3, model="rsw", n=8192, bytes_pt=440, solves=0,
                    what="rotating shallow water {nx}x{ny} fp64, closed basin with 4 disc islands + peninsula, "
                         "geos_adj dipole, f0 = 10, WENO5-Z + SSP-RK3"),
    "qgrsw8192": dict(baseline_config=4, model="qgrsw", n=8192, bytes_pt=340, solves=3,
                      what="QG projection method (Thiry et al 2024) {nx}x{ny} fp64, same masked basin + Gaussian "
                           "topography, one vertex Helmholtz solve per RK stage"),
    "bouss16384": dict(baseline_config=5, model="boussinesq", n=16384, xperiodic=True, bytes_pt=590, solves=3,
                       what="Boussinesq vertical plane {nx}x{ny} fp64, x-periodic, b = y + 0.1 gaussian "
                            "(warm_bubble.py:14-20)"),
}


def param_for(cfg, n, Param, ny_factor=1):
    p = Param()
    p.model = cfg["model"]
    if cfg["n"] is None:                 # vortex.py:59-71
        p.nx, p.ny, p.Lx, p.Ly = 200, 100, 2.0, 1.0
        p.dt, p.f0, p.noslip = 0.2, 0.0, False
    else:
        p.nx = n
        p.ny = n * ny_factor
        p.Lx, p.Ly = 1.0, 1.0 * ny_factor
    p.xperiodic = bool(cfg.get("xperiodic", False))
    p.ywrap = bool(cfg.get("ywrap", False))
    p.integrator = "rk3"
    p.vortexforce = p.innerproduct = p.compflux = "weno"
    p.maxorder = 6
    p.cfl = 0.9
    if cfg["model"] in ("rsw", "qgrsw"):
        p.f0, p.dtmax = 10.0, 1.0
    if cfg["model"] == "boussinesq":
        p.dtmax = 1e-1
    return p


def initial_condition(name, cfg, model, f2d, allreduce_max):
    """the configuration's initial state through the public API; returns nothing"""
    mesh, s, p = model.mesh, model.state, model.param
    slab = mesh.slab
    if name == "vortex":                  # vortex.py:8-42
        x, y = mesh.xy("v")
        s.omega[...] = gaussian(x, y, 1.05, 0.5, 0.05) - gaussian(x, y, 0.95, 0.5, 0.05)
        s.omega[...] *= mesh.mskv * mesh.area
        f2d.tools.set_uv_from_omega(model, s.omega, s.u)
    elif cfg["model"] == "euler":
        s.omega[...] = turbulence_vorticity(mesh.x("v"), mesh.y("v"), mesh.area, yperiod=p.Ly if p.ywrap else None)
        s.omega[...] *= mesh.mskv
        if p.ywrap:                   # the vertex Poisson problem of a domain without walls needs a zero-mean right-hand side
            nh = p.halowidth
            s.omega[...] -= s.omega[nh:-nh, nh:-nh].mean()
        f2d.tools.set_uv_from_omega(model, s.omega, s.u)
        umax = allreduce_max(max(np.abs(s.u.x).max() / mesh.dx, np.abs(s.u.y).max() / mesh.dy))
        s.u.x[...] *= 1.0 / umax      # physical speed ~1 -> CFL dt ~ 1 / n
        s.u.y[...] *= 1.0 / umax
    elif cfg["model"] in ("rsw", "qgrsw"):
        mesh.msk[...] = basin_mask(mesh.shape, p.halowidth, slab.row0 if slab.nranks > 1 else 0, p.ny)
        mesh.finalize()
        x, y = mesh.x("c")[None, :], mesh.y("c")[:, None]
        if cfg["model"] == "qgrsw":       # rsw_with_topo.py:96-99
            mesh.hb = 0.2 * gaussian(x, y, 0.3, 0.7, 0.05) * mesh.area * mesh.msk
        s.h[...] = p.H + 0.2 * (gaussian(x, y, 0.6, 0.5, 0.1) - gaussian(x, y, 0.4, 0.5, 0.1))     # geos_adj.py:12-49
        s.h[...] *= mesh.msk * mesh.area
        if cfg["model"] == "qgrsw":
            s.h[...] -= mesh.hb
            f2d.operators.qg_projection(mesh, s.u, s.h, s.pv, s.psi)
    elif cfg["model"] == "boussinesq":    # warm_bubble.py:14-20, scaled to the unit box
        x, y = mesh.x("c")[None, :], mesh.y("c")[:, None]
        s.b[...] = (y + 0.1 * gaussian(x, y, 0.5 * p.Lx, 0.25, 0.08)) * mesh.msk
    model.integrator.diag(s)


def clock_sampler(device_index, stop_evt, out):
    """nvidia-smi clocks / throttle reasons while the timed region runs"""
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
        pr = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                               "-lms", "100", "-i", str(device_index)],
                              stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return
    stop_evt.wait()
    time.sleep(0.15)
    pr.terminate()
    try:
        txt = pr.communicate(timeout=5)[0]
    except Exception:
        txt = ""
    for line in txt.strip().splitlines():
        f = [x.strip() for x in line.split(",")]
        if len(f) >= 7:
            out.append(f)


def summarize_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = sorted(float(s[0]) for s in samples)
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for k, n in enumerate(names) if any(s[3 + k].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(samples[0][1]), "reasons": reasons,
            "power_w_max": max(float(s[2]) for s in samples), "samples": len(samples)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------- CPU baseline ---
def cpu_port_run(model_name, n, steps, warmup):
    """The reference's algorithm on the host cores: the oracle port (numpy + C WENO
    kernels + SuperLU).  Returns (updates/s, seconds per step, set-up seconds)."""
    from oracle import fluids2d_oracle as orc
    orc.build()
    kw = dict(model=model_name, nx=n, ny=n)
    if model_name in ("euler", "boussinesq"):
        kw["xperiodic"] = True
    if model_name in ("rsw", "qgrsw"):
        kw.update(f0=10.0, dtmax=1.0)
    p = orc.make_param(**kw)
    t0 = time.time()
    m = orc.Model(p, msk=basin_mask((n + 6, n + 6)) if model_name in ("rsw", "qgrsw") else None)
    s, mesh = m.state, m.mesh
    xv, yv = mesh.xy("v")
    x, y = mesh.xy("c")
    if model_name == "euler":
        s.omega[...] = turbulence_vorticity(xv[0], yv[:, 0], mesh.area) * mesh.mskv
        orc.set_uv_from_omega(mesh, s.omega, s.u)
        mesh.poisson_centers.A_LU              # factorise inside the set-up time, as Mesh.finalize does
    elif model_name == "boussinesq":
        s.b[...] = (y + 0.1 * gaussian(x, y, 0.5, 0.25, 0.08)) * mesh.msk
        mesh.poisson_centers.A_LU
    else:
        if model_name == "qgrsw":
            mesh.hb = 0.2 * gaussian(x, y, 0.3, 0.7, 0.05) * mesh.area * mesh.msk
        s.h[...] = (p.H + 0.2 * (gaussian(x, y, 0.6, 0.5, 0.1) - gaussian(x, y, 0.4, 0.5, 0.1))) * mesh.msk * mesh.area
        if model_name == "qgrsw":
            s.h[...] -= mesh.hb
            orc.qg_projection(mesh, s.u, s.h, s.pv, s.psi)
    setup = time.time() - t0
    m.diag(s)
    dt = m.compute_dt()
    for _ in range(warmup):
        m.step(dt)
    t0 = time.time()
    for _ in range(steps):
        m.step(dt)
    el = time.time() - t0
    return n * n * steps / el, el / steps, setup


def cpu_reference_itself(n, steps, warmup):
    """The UNMODIFIED reference (baseline/_ref, installed by oracle/install_ref.py:
    numba kernels + numpy + scipy SuperLU) on the headline workload at a bounded
    size, driven through its own public API: Param, Model, tools.set_uv_from_omega,
    integrator.step.  Returns (updates/s, s/step, set-up s) or None if it cannot run here."""
    try:
        from oracle import refshim
        if not refshim.vendored_available():
            return None
        t0 = time.time()
        ref = refshim.load(refshim.VENDORED)
        p = ref.Param()
        p.model, p.nx, p.ny, p.xperiodic = "euler", n, n, True
        p.integrator, p.maxorder, p.cfl = "rk3", 6, 0.9
        p.vortexforce = p.innerproduct = p.compflux = "weno"
        p.animation, p.nhis = False, 0
        model = ref.Model(p)
        mesh, s = model.mesh, model.state
        xv, yv = mesh.xy("v")
        s.omega[...] = turbulence_vorticity(xv[0], yv[:, 0], mesh.area) * mesh.mskv
        ref.tools.set_uv_from_omega(model, s.omega, s.u)
        umax = max(np.abs(s.u.x).max() / mesh.dx, np.abs(s.u.y).max() / mesh.dy)
        s.u.x[...] *= 1.0 / umax
        s.u.y[...] *= 1.0 / umax
        model.integrator.diag(s)
        setup = time.time() - t0
        model.set_dt()
        for _ in range(warmup):
            model.integrator.step(s, model.time)
        t0 = time.time()
        for _ in range(steps):
            model.integrator.step(s, model.time)
        el = time.time() - t0
        if not np.isfinite(s.u.x).all():
            return None
        return n * n * steps / el, el / steps, setup
    except Exception as e:           # numba or the package missing on this box
        print(f"[bench] reference itself unavailable ({type(e).__name__}: {e}); using the oracle port", file=sys.stderr)
        return None


def cpu_baseline(model_name, n, steps, warmup):
    """bounded CPU sample of the workload: the reference itself for the Euler headline
    when it is installed, else (and for the other models) the oracle port"""
    r = cpu_reference_itself(n, steps, warmup) if model_name == "euler" else None
    kind = "reference"
    if r is None:
        r = cpu_port_run(model_name, n, steps, warmup)
        kind = "port"
    v, sps, setup = r
    what = ("the unmodified reference (baseline/_ref: numba kernels + numpy + scipy SuperLU, its own Param/Model/"
            "integrator.step)" if kind == "reference" else
            "oracle port of the reference CPU path (numpy + C WENO kernels + scipy SuperLU)")
    sample = (f"{what}, {model_name} {n}^2 as a bounded sample of the workload (the reference's LU cannot reach "
              f"4096^2, SURVEY section 0 fact 5); {steps} steps after {warmup} warm-up, {sps:.3f} s/step; "
              f"set-up (mesh + LU{' + numba JIT' if kind == 'reference' else ''}) {setup:.1f} s not counted; "
              "1 core: the reference's only threaded kernel (compflux, param.nthreads) is not on the Euler path "
              "and SuperLU's triangular solves are serial")
    return {"value": v, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
            "host_cores_available": os.cpu_count()}, sps


# ------------------------------------------------------------------- ours -----
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import fluids2d_b200 as f2d
    f2d.Param._quiet = True
    cfg = CONFIGS[args.config]
    n = args.n or cfg["n"]
    # N > 1, default: WEAK scaling -- every GPU owns an n x n slab of one grid
    # that is N times taller (n x nN cells, Ly = N); --strong keeps the n x n
    # grid and splits it; --replicas runs N independent n x n grids.
    weak = world > 1 and not args.strong and not args.replicas
    p = param_for(cfg, n, f2d.Param, world if weak else 1)
    p.device = local
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from fluids2d_b200 import slabs
        slabs.init_from_torch_distributed()
        if args.replicas:
            world_slabs = 1          # N independent copies of the whole grid
        else:
            p.rank, p.nranks = rank, world      # ONE grid, y-slabs over the GPUs
            world_slabs = world
    else:
        world_slabs = 1
    p.solver_nu = args.nu
    p.solver_rtol = args.rtol
    p.solver_guess = args.guess
    model = f2d.Model(p)
    mesh, s, eng, integ = model.mesh, model.state, model.mesh.engine, model.integrator

    def allreduce_max(v):
        if world_slabs > 1:
            t = torch.tensor([v], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

    initial_condition(args.config, cfg, model, f2d, allreduce_max)
    model.set_dt()
    dt = model.time.dt
    if cfg["model"] == "boussinesq":
        # the fluid starts at rest (the CFL step is dtmax), and the bubble then accelerates: a step
        # fixed for the whole run has to respect the CFL limit of the speeds it reaches, O(1)
        dt = min(dt, 0.5 * min(mesh.dx, mesh.dy))
    p.dt = dt                     # identical steps from here on
    model.time.dt = dt

    def barrier():
        eng.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) ------------------------------
    integ.upload(s)
    for _ in range(args.warmup):
        eng.step(dt, 1)
    eng.solver_stats()
    stop_evt, samples = threading.Event(), []
    th = threading.Thread(target=clock_sampler, args=(local, stop_evt, samples))
    if rank == 0:
        th.start()
    barrier()
    l0 = eng.launch_count()
    x0 = eng.exchange_count()
    eng.timer_start()
    for _ in range(args.steps):
        eng.step(dt, 1)
    ms = eng.timer_stop()
    barrier()
    launches = eng.launch_count() - l0
    nexch = eng.exchange_count() - x0
    stats = eng.solver_stats()
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    copies = world // world_slabs          # 1 when the grid is decomposed
    npoints = p.nx * p.ny                     # points of ONE grid
    value = copies * npoints * args.steps / (ms * 1e-3)
    iters_per_solve = stats["niters"] / max(stats["nsolves"], 1)

    # ---- end to end through the public per-step call, host buffers -----------
    integ.download(s)
    for _ in range(2):
        integ.step(s, model.time)
    nfields = len(integ._step_outputs(s))     # what comes back every step ...
    nin = len(integ._step_inputs(s))          # ... the ones a (not the first) step reads go in: counted after the
                                              # warm-up steps, which seed the solutions of the step's own solves
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        integ.step(s, model.time)     # H2D state, fused step, D2H state
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    if rank == 0:
        stop_evt.set()
        th.join()
    e2e_value = copies * npoints * args.steps / e2e_s
    field_bytes = mesh.shape[0] * mesh.shape[1] * 8 * world    # all ranks together
    probe = s.u.x if hasattr(s, "u") else s.omega
    ok = bool(np.isfinite(probe).all() and np.isfinite(s.omega).all())

    # ---- roofline of the individual kernels, timed alone on the same stream ---
    peak, peak_src = measured_peak()
    kernels = {}
    for name in ([] if args.no_kernels else eng.bench_kernel_names()):
        kms, kbytes = eng.bench_kernel(name, 20)
        per_step = eng.launches_per_step(name, iters_per_solve)
        kernels[name] = {"ms": round(kms, 4), "alg_bytes": kbytes,
                         "gbs": round(kbytes / (kms * 1e-3) / 1e9, 1) if kbytes else None,
                         "frac": round(kbytes / (kms * 1e-3) / 1e9 / peak, 3) if kbytes else None,
                         "launches_per_step": round(per_step, 2), "ms_per_step": round(kms * per_step, 4)}
    roof = None
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tpath) and args.config == "euler4096" and n == 4096:
        with open(tpath) as f:
            traffic = json.load(f)      # DRAM bytes per launch from the committed ncu --set full capture
    for k, v in kernels.items():
        v["traffic"] = traffic.get(k)
    if kernels:
        # the dominant kernel = the bandwidth-bound kernel with the largest share of a step,
        # from the times measured in this run x its launches per step
        cand = {k: v for k, v in kernels.items() if v["alg_bytes"]}
        dom = max(cand, key=lambda k: cand[k]["ms_per_step"])
        alg_step = (cfg["bytes_pt"] + cfg["solves"] * iters_per_solve * 183.0) * npoints / max(world_slabs, 1)
        roof = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak,
                "unit": "GB/s", "frac": kernels[dom]["frac"], "traffic": traffic.get(dom), "peak_source": peak_src,
                "whole_step": {"alg_bytes_per_gpu": alg_step,
                               "formula": f"({cfg['bytes_pt']} + {cfg['solves']} solves x {iters_per_solve:.2f} "
                                          "iterations x 183) B/pt (SURVEY 8d)",
                               "frac": round(alg_step / (ms / args.steps * 1e-3) / 1e9 / peak, 3)},
                "kernels": kernels}

    out = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            cpu, _ = cpu_baseline(cfg["model"], args.cpu_n, 3, 1)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["what"].format(nx=p.nx, ny=p.ny)
                                   + (f" ({n}^2 per GPU, weak scaling)" if weak else ""),
                       "name": args.config, "baseline_config": cfg["baseline_config"],
                       "grid": [p.ny, p.nx], "dt": dt, "per_gpu": ("one independent replica per GPU" if copies > 1 else
                                   (f"y-slab of {p.ny // world} rows + 8 ghost rows per interface, NCCL halo exchange"
                                    if world > 1 else "whole grid")),
                       "exchanges_per_step": nexch / max(args.steps, 1),
                       "l2_note": f"every field is {mesh.shape[0] * mesh.shape[1] * 8 / 1e6:.1f} MB against 126 MB of L2"
                                  + ("; no flush needed" if mesh.shape[0] * mesh.shape[1] * 8 > 126e6 else
                                     ": the working set is L2-resident by nature of this small configuration"),
                       "solver": {"kind": p.solver, "rtol": p.solver_rtol,
                                  "iters_per_solve": iters_per_solve,
                                  "max_relres": stats["max_relres"]}},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_s / args.steps * 1e3,
                    "h2d_bytes_per_step": nin * field_bytes, "d2h_bytes_per_step": nfields * field_bytes,
                    "call": "integrator.step(state, time) with pinned numpy state"},
            "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
            "clocks": summarize_clocks(samples), "finite": ok,
        }
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    n = args.cpu_n
    cpu, sps = cpu_baseline(cfg["model"], n, args.steps, min(args.warmup, 1))
    v = cpu["value"]
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT,
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": min(args.warmup, 1),
        "ms_per_step": sps * 1e3, "higher_is_better": True, "scaling": "strong" if args.strong else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{cfg['model']} {n}^2 fp64, WENO5-Z + SSP-RK3 (bounded sample of: "
                               + cfg["what"].format(nx=cfg["n"] or 200, ny=cfg["n"] or 100) + ")",
                   "name": args.config, "baseline_config": cfg["baseline_config"]},
        "cpu_baseline": cpu,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


_json_out = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL
    prints its version banner on stdout when NCCL_DEBUG is WARN or VERSION), so
    keep a private handle on the real stdout for the JSON line and point fd 1 --
    and with it every other writer in the process -- at stderr."""
    global _json_out
    sys.stdout.flush()
    _json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(obj):
    out = _json_out if _json_out is not None else sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


if __name__ == "__main__":
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=6,
                    help="untimed steps; the first-guess history of the elliptic solves (4 steps deep) fills during them")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="euler4096", choices=sorted(CONFIGS),
                    help="BASELINE.json configuration (default: the headline, config 2)")
    ap.add_argument("--grid", dest="n", type=int, default=0, help="grid size n (default: the configuration's)")
    ap.add_argument("--cpu-n", type=int, default=512, help="grid of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--strong", action="store_true", help="N > 1: split the n x n grid instead of growing it with N")
    ap.add_argument("--replicas", action="store_true", help="N > 1: independent copies instead of one decomposed grid")
    ap.add_argument("--sweeps", dest="nu", type=int, default=2, help="smoothing sweeps per V-cycle leg")
    ap.add_argument("--guess", type=int, default=4, help="first-guess extrapolation order (0 off)")
    ap.add_argument("--rtol", type=float, default=1e-12, help="elliptic solver tolerance")
    ap.add_argument("--no-kernels", action="store_true", help="skip the per-kernel roofline timings")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
