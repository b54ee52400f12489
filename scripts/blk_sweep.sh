for blk in 64x4 32x8 128x2 32x4 64x2 128x1 16x16; do
F2D_BLK=$blk python - <<PY
import sys; sys.path.insert(0,'.')
import numpy as np, bench
import fluids2d_b200 as f2d
f2d.Param._quiet=True
p=bench.param_for(4096,f2d.Param); m=f2d.Model(p)
s=m.state; s.omega[...]=bench.turbulence_vorticity(m.mesh.x("v"),m.mesh.y("v"),m.mesh.area); s.omega[...]*=m.mesh.mskv
f2d.tools.set_uv_from_omega(m,s.omega,s.u); m.integrator.diag(s); m.integrator.upload(s)
e=m.mesh.engine
print("$blk", {k: round(e.bench_kernel(k,20)[0],4) for k in ("advection","divergence")})
PY
done
