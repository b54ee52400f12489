"""one launch of each hot kernel inside an NVTX range, for
ncu --set full --nvtx --nvtx-include "f2dprof/" python scripts/kprof.py [--config NAME] [--n N] [names...]
(default: the 4096^2 Euler workload of bench.py; --config rsw8192 --n 4096 profiles the rsw kernels)"""
import sys
sys.path.insert(0, ".")
import torch
import bench
import fluids2d_b200 as f2d

args = sys.argv[1:]
config, n = "euler4096", None
while args and args[0].startswith("--"):
    if args[0] == "--config":
        config = args[1]
    elif args[0] == "--n":
        n = int(args[1])
    args = args[2:]
cfg = bench.CONFIGS[config]
f2d.Param._quiet = True
p = bench.param_for(cfg, n or cfg["n"] or 0, f2d.Param)
m = f2d.Model(p)
s = m.state
bench.initial_condition(config, cfg, m, f2d, lambda v: v)
m.integrator.upload(s)
e = m.mesh.engine
m.set_dt()
e.step(m.time.dt if config != "euler4096" else 1e-4, 2)
names = args or e.bench_kernel_names()
for k in names:
    e.bench_kernel(k, 3)          # warm
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("f2dprof")
for k in names:
    e.bench_kernel(k, 1)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
