"""Where does the one-kernel scalar transport differ from the flux / divergence / update kernels?
Runs a golden case one step under F2D_STAGE=tma and =point in subprocesses and prints, per field,
the largest difference and the rows / columns it sits in."""
import os
import subprocess
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SNIP = """
import sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + '/tests')
import numpy as np
from util import Golden, engine_for
g = Golden({case!r})
e = engine_for(g)
for dt in g.dts[:{nsteps}]:
    e.step(dt, 1)
out = dict()
for f in {fields!r}:
    out[f] = e.download(f)
e.close()
np.savez(sys.argv[1], **out)
"""
case = sys.argv[1] if len(sys.argv) > 1 else "lock_exchange"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
fields = ("b", "ds0.b", "ds1.b", "u.x", "u.y") if "rsw" not in case else ("h", "ds0.h", "ds1.h", "u.x", "u.y")
res = {}
for tag, env in (("tma", {}), ("point", {"F2D_STAGE": "point"})):
    path = f"/tmp/td_{tag}.npz"
    subprocess.run([sys.executable, "-c", SNIP.format(root=ROOT, case=case, fields=fields, nsteps=nsteps), path], check=True,
                   env={**os.environ, **env})
    res[tag] = np.load(path)
for f in fields:
    a, b = res["tma"][f], res["point"][f]
    d = np.abs(a - b)
    jj, ii = np.nonzero(d)
    print(f, "max diff", d.max(), "of", np.abs(b).max(), "count", len(jj),
          "rows", sorted(set(jj.tolist()))[:30], "cols", sorted(set(ii.tolist()))[:110])
