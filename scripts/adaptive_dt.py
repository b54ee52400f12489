"""Iterations per elliptic solve with the reference's default adaptive time step
(cfl-based dt, model.py:71-87), Euler n^2 x-periodic channel, `steps` steps."""
import sys
import numpy as np
sys.path.insert(0, ".")
import bench
import fluids2d_b200 as f2d

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
f2d.Param._quiet = True
p = bench.param_for(n, f2d.Param)
p.dt = 0.0
model = f2d.Model(p)
s, mesh = model.state, model.mesh
s.omega[:] = bench.turbulence_vorticity(mesh.x("v"), mesh.y("v"), mesh.area) * mesh.mskv
f2d.tools.set_uv_from_omega(model, s.omega, s.u)
model.integrator.diag(s)
umax = np.abs(s.u.x).max() / mesh.dx
s.u.x[...] *= 1 / umax
s.u.y[...] *= 1 / umax
model.integrator.diag(s)
model.integrator.upload(s)
e = mesh.engine
dts = []
for k in range(steps):
    if k == 10:
        st0 = e.solver_stats()
    model.set_dt(on_device=True)
    dts.append(model.time.dt)
    model.integrator.step_resident(model.time.dt, 1)
    model.time.pushforward()
st = e.solver_stats()
print(f"n={n} steps={steps} dt {dts[0]:.4e}..{dts[-1]:.4e} (rel. change/step {np.abs(np.diff(dts)).mean()/dts[0]:.2e}) "
      f"iters/solve after warm-up {(st['niters']-st0['niters'])/(st['nsolves']-st0['nsolves']):.2f} max_relres {st['max_relres']:.2e}")
