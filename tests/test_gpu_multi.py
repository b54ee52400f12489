"""Two-GPU parity: one grid split into y-slabs (NCCL ghost-row exchange, CG
scalars all-reduced) must reproduce the single-GPU run.  Skipped with < 2 GPUs."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


CASES = {
    "closed_islands": dict(nx=256, ny=256, dt=0.05, steps=4, islands=True, noslip=True),
    "xper_channel": dict(nx=256, ny=192, Lx=2.0, dt=0.05, steps=4, xperiodic=True, noslip=["bottom"]),
    "boussinesq": dict(model="boussinesq", nx=192, ny=128, Lx=1.5, dt=0.02, steps=3, islands=True),
    # rotating shallow water (no solve in the step): u*, h* exchanged once per RK stage, the diagnostics
    # computed on the ghost rows; islands straddling the interfaces, topography, adaptive dt
    "rsw_islands": dict(model="rsw", nx=256, ny=192, Lx=1.0, dt=0.0, dtmax=1.0, f0=10.0, steps=8, islands=True, noslip=True),
    # the QG-projected shallow water model: one vertex Helmholtz solve per stage, no message besides the solver's
    "qgrsw_islands": dict(model="qgrsw", nx=256, ny=192, Lx=1.0, dt=0.0, dtmax=1.0, f0=10.0, steps=5, islands=True),
    # BASELINE config 2 (bench.py's workload) at 1024^2: four tile levels per slab even on 8 ranks,
    # open-tile kernels, CUDA-graph iterations, the cubic first guess -- ten steps
    "config2_1024": dict(nx=1024, ny=1024, dt=0.0, steps=10, xperiodic=True, turbulence=True),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_slabs_match_single_gpu(name, nproc):
    if ngpus() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    case = CASES[name]
    if case["ny"] // nproc < 32:
        pytest.skip("too few rows per rank")
    if nproc == 8 and name != "config2_1024":
        pytest.skip("the small cases have too few rows for 8 slabs")
    if nproc > 2 and name in ("rsw_islands", "qgrsw_islands"):
        pytest.skip("rsw / qgrsw slabs were added late in round 2 and verified on 2 GPUs (parity) and, rsw, on "
                    "2 / 4 / 8 GPUs through bench.py only")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(29611 + nproc), os.path.join(ROOT, "tests", "dist_worker.py"),
           json.dumps(case)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    line = [l for l in r.stdout.splitlines() if l.startswith("DIST_RESULT ")]
    assert line, r.stdout[-2000:] + r.stderr[-4000:]
    out = json.loads(line[0][len("DIST_RESULT "):])
    print(name, nproc, out)
    assert out["exchanges"] > 0
    for k, v in out["errors"].items():
        assert v <= 1e-10, (k, v)
    # the decomposition does not change the convergence of the solver
    assert out["solver"]["niters"] <= out["ref_solver"]["niters"] + out["ref_solver"]["nsolves"]
    if case.get("model") == "rsw":
        assert out["solver"]["nsolves"] == 0
