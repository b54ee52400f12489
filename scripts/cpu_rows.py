"""CPU side-by-side rows of SURVEY 8d: the reference itself (baseline/_ref) and the oracle
port on the host cores, full Euler path (x-periodic channel, bench.py's initial condition),
at the grid sizes given on the command line (default 512 1024).  One JSON line per row.

    python scripts/cpu_rows.py [n ...]      # run on the GPU box's host while the GPU is busy elsewhere
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

for n in [int(a) for a in sys.argv[1:]] or [512, 1024]:
    steps = 3 if n <= 512 else 2
    for kind, fn in (("reference", lambda: bench.cpu_reference_itself(n, steps, 1)),
                     ("port", lambda: bench.cpu_port_run("euler", n, steps, 1))):
        r = fn()
        if r is None:
            print(json.dumps({"kind": kind, "n": n, "unavailable": True}), flush=True)
            continue
        v, sps, setup = r
        print(json.dumps({"kind": kind, "n": n, "updates_per_s": v, "s_per_step": sps, "setup_s": setup, "cores": 1,
                          "host_cores_available": os.cpu_count(), "steps": steps}), flush=True)
