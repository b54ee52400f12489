"""time the hot kernels alone at 4096^2 (python scripts/kbench.py [names...])"""
import sys
sys.path.insert(0, ".")
import bench
import fluids2d_b200 as f2d

f2d.Param._quiet = True
p = bench.param_for(bench.CONFIGS["euler4096"], 4096, f2d.Param)
m = f2d.Model(p)
s = m.state
s.omega[...] = bench.turbulence_vorticity(m.mesh.x("v"), m.mesh.y("v"), m.mesh.area)
s.omega[...] *= m.mesh.mskv
f2d.tools.set_uv_from_omega(m, s.omega, s.u)
m.integrator.diag(s)
m.integrator.upload(s)
e = m.mesh.engine
e.step(1e-4, 2)
names = sys.argv[1:] or e.bench_kernel_names()
print({k: round(e.bench_kernel(k, 30)[0], 4) for k in names})
