#!/usr/bin/env python
"""Headline benchmark: grid-point updates/s incl. elliptic solve, 4096^2 Euler.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one full RK3 step of the Euler model (3 stages = 3 WENO advection
kernels + 3 RK updates + 3 pressure projections, each with a multigrid-PCG
elliptic solve + vorticity/kinetic-energy diagnostics) on a 4096^2 grid,
x-periodic channel (the reference-supported variant of BASELINE config 2, see
SURVEY note Y), fp64, WENO5-Z + SSP-RK3, fixed dt from CFL 0.9.

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
def turbulence_vorticity(x, y, area, seed=0, kpeak=8.0, kwidth=3.0, nmodes=96):
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
    phase = rng.uniform(0, 2 * np.pi, nmodes)
    amp = rng.normal(0, 1, nmodes) * (area / np.sqrt(nmodes))
    ax = 2 * np.pi * np.outer(x, kx) + phase          # (n1, M)
    by = 2 * np.pi * np.outer(y, ky)                  # (n2, M)
    # cos(a + b) = cos a cos b - sin a sin b
    return (np.cos(by) * amp) @ np.cos(ax).T - (np.sin(by) * amp) @ np.sin(ax).T


def param_for(n, Param, ny_factor=1):
    p = Param()
    p.model = "euler"
    p.nx = n
    p.ny = n * ny_factor
    p.Lx, p.Ly = 1.0, 1.0 * ny_factor
    p.xperiodic = True
    p.integrator = "rk3"
    p.vortexforce = p.innerproduct = p.compflux = "weno"
    p.maxorder = 6
    p.cfl = 0.9
    return p


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
def cpu_reference_run(n, steps, warmup):
    """The reference's algorithm on the host cores: the oracle port (numpy + C
    WENO kernels + SuperLU), the only CPU implementation that can travel to the
    GPU box (the reference itself is Python under /root/reference).  Returns
    (updates/s, seconds per step, set-up seconds)."""
    from oracle import fluids2d_oracle as orc
    orc.build()
    p = orc.make_param(model="euler", nx=n, ny=n, xperiodic=True)
    t0 = time.time()
    m = orc.Model(p)
    setup = time.time() - t0
    xv, yv = m.mesh.xy("v")
    m.state.omega[...] = turbulence_vorticity(xv[0], yv[:, 0], m.mesh.area) * m.mesh.mskv
    orc.set_uv_from_omega(m.mesh, m.state.omega, m.state.u)
    m.diag(m.state)
    dt = m.compute_dt()
    for _ in range(warmup):
        m.step(dt)
    t0 = time.time()
    for _ in range(steps):
        m.step(dt)
    el = time.time() - t0
    return n * n * steps / el, el / steps, setup


# ------------------------------------------------------------------- ours -----
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import fluids2d_b200 as f2d
    f2d.Param._quiet = True
    n = args.n
    # N > 1, default: WEAK scaling -- every GPU owns an n x n slab of one grid
    # that is N times taller (n x nN cells, Ly = N); --strong keeps the n x n
    # grid and splits it; --replicas runs N independent n x n grids.
    weak = world > 1 and not args.strong and not args.replicas
    p = param_for(n, f2d.Param, world if weak else 1)
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

    # initial condition through the public API (device Poisson solve for psi)
    s.omega[...] = turbulence_vorticity(mesh.x("v"), mesh.y("v"), mesh.area)
    s.omega[...] *= mesh.mskv
    f2d.tools.set_uv_from_omega(model, s.omega, s.u)
    umax = max(np.abs(s.u.x).max() / mesh.dx, np.abs(s.u.y).max() / mesh.dy)
    if world_slabs > 1:
        tmax = torch.tensor([umax], device="cuda", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        umax = float(tmax.item())
    s.u.x[...] *= 1.0 / umax      # physical speed ~1 -> CFL dt ~ 1e-4
    s.u.y[...] *= 1.0 / umax
    integ.diag(s)
    model.set_dt()
    dt = model.time.dt
    p.dt = dt                     # identical steps from here on

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

    # ---- end to end through the public per-step call, host buffers -----------
    integ.download(s)
    nfields = len(integ._names(s))            # every field comes back ...
    nin = len(integ._step_inputs(s))          # ... the ones the step reads go in
    for _ in range(2):
        integ.step(s, model.time)
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
    ok = bool(np.isfinite(s.u.x).all() and np.isfinite(s.omega).all())

    # ---- roofline of the individual kernels, timed alone on the same stream ---
    peak, peak_src = measured_peak()
    kernels = {}
    for name in ([] if args.no_kernels else eng.bench_kernel_names()):
        kms, kbytes = eng.bench_kernel(name, 20)
        kernels[name] = {"ms": round(kms, 4), "alg_bytes": kbytes,
                         "gbs": round(kbytes / (kms * 1e-3) / 1e9, 1) if kbytes else None,
                         "frac": round(kbytes / (kms * 1e-3) / 1e9 / peak, 3) if kbytes else None}
    dom = eng.dominant_kernel()
    roof = None
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath) and n == 4096:
        with open(tpath) as f:
            traffic = json.load(f)      # DRAM bytes per launch from the committed ncu --set full capture
    for k, v in kernels.items():
        v["traffic"] = traffic.get(k)
    if kernels:
        roof = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak,
                "unit": "GB/s", "frac": kernels[dom]["frac"], "traffic": traffic.get(dom), "peak_source": peak_src,
                "kernels": kernels}

    out = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            v, sps, setup = cpu_reference_run(args.cpu_n, 3, 1)
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"oracle port (numpy + C WENO kernels + SuperLU direct solve), Euler "
                             f"{args.cpu_n}^2 x-periodic, 3 steps after 1 warm-up, {sps:.3f} s/step; "
                             f"LU set-up {setup:.1f} s not counted; the reference cannot reach 4096^2 "
                             f"(SURVEY section 0 fact 5)", "host_cores_available": os.cpu_count()}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak" if (copies > 1 or weak or world == 1) else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"Euler {p.nx}x{p.ny} fp64, x-periodic channel, WENO5-Z + SSP-RK3, "
                                   "band-limited random vorticity (rng 0), fixed dt = CFL 0.9"
                                   + (f" ({n}^2 per GPU, weak scaling)" if weak else ""),
                       "grid": [p.ny, p.nx], "dt": dt, "per_gpu": ("one independent replica per GPU" if copies > 1 else
                                   (f"y-slab of {p.ny // world} rows + 8 ghost rows per interface, NCCL halo exchange"
                                    if world > 1 else "whole grid")),
                       "exchanges_per_step": nexch / max(args.steps, 1),
                       "l2_note": "every field is 134.6 MB > 126 MB L2; no flush needed",
                       "solver": {"kind": p.solver, "rtol": p.solver_rtol,
                                  "iters_per_solve": stats["niters"] / max(stats["nsolves"], 1),
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
    n = args.cpu_n
    v, sps, setup = cpu_reference_run(n, args.steps, min(args.warmup, 1))
    sample = (f"oracle port of the reference CPU path (numpy + C WENO kernels + scipy SuperLU), Euler "
              f"{n}^2 x-periodic as a bounded sample of the 4096^2 workload (the reference's LU cannot "
              f"reach 4096^2); {sps:.3f} s/step; LU set-up {setup:.1f} s not counted")
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT,
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": min(args.warmup, 1),
        "ms_per_step": sps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"Euler {n}^2 fp64 x-periodic channel, WENO5-Z + SSP-RK3 (bounded sample of the 4096^2 workload)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "host_cores_available": os.cpu_count()},
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
    ap.add_argument("--grid", dest="n", type=int, default=4096, help="grid size n (default: the headline 4096)")
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
