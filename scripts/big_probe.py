"""python scripts/big_probe.py model n [steps]: a few resident steps of a model at n^2 through the Engine
(no host state beyond the uploaded fields), printing solver statistics -- for F2D_DEBUG=1 runs."""
import sys
import time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from util import plain_param
from fluids2d_b200._cabi import Engine

model, n = sys.argv[1], int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
p = plain_param(model=model, nx=n, ny=n, xperiodic=True)
t0 = time.time()
e = Engine(p)
e.set_mask(None)
print("set_mask %.1fs" % (time.time() - t0), e.solver_info("c"), flush=True)
n2, n1 = e.shape
dx = 1.0 / n
y = ((np.arange(n2) - 3 + 0.5) * dx)[:, None]
x = ((np.arange(n1) - 3 + 0.5) * dx)[None, :]
msk = e.mesh_array("msk")
if model == "boussinesq":
    b = (y + 0.1 * np.exp(-((x - 0.5) ** 2 + (y - 0.25) ** 2) / (2 * 0.08 ** 2))) * msk
    e.upload("b", b)
    del b
    dt = 1e-2
else:
    k = 2 * np.pi * 8
    ux = (np.sin(k * y) * np.cos(k * x)) * dx * e.mesh_array("mskx")
    uy = (-np.cos(k * y) * np.sin(k * x)) * dx * e.mesh_array("msky")
    e.upload("u.x", ux); e.upload("u.y", uy)
    del ux, uy
    dt = 0.4 * dx
e.diag()
print("diag", e.solver_stats(), flush=True)
for s in range(steps):
    e.step(dt, 1)
    print("step", s, e.solver_stats(), flush=True)
print("ok", float(np.abs(e.download("p")).max()))
