#!/bin/bash
# Re-measure the rows DESIGN.md section 9 marks (†): they predate the open-tile
# multigrid kernels.  Run on the GPU box, one call per N (gpurun --gpus N):
#   gpurun --timeout 300 -- 'bash scripts/remeasure_scaling.sh 1'
#   gpurun --gpus 8 --timeout 400 -- 'bash scripts/remeasure_scaling.sh 8'
# JSON lines land in gpurun_out/; copy the ones to keep into profiles/.
set -u
N=${1:-1}
mkdir -p gpurun_out
run() {  # run <tag> <bench flags...>
    local tag=$1; shift
    if [ "$N" -eq 1 ]; then
        python bench.py --gpus 1 "$@" > "gpurun_out/bench_${tag}_n1.json" 2> "gpurun_out/bench_${tag}_n1.err"
    else
        python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
            --master-port 29533 bench.py --gpus "$N" "$@" \
            > "gpurun_out/bench_${tag}_n${N}.json" 2> "gpurun_out/bench_${tag}_n${N}.err"
    fi
    tail -c 600 "gpurun_out/bench_${tag}_n${N}.json"; echo
}
if [ "$N" -eq 1 ]; then
    run 4096 --steps 10 --warmup 6
    run 16384 --grid 16384 --steps 4 --warmup 6 --no-cpu --no-kernels
else
    run weak --steps 10 --warmup 6 --no-cpu --no-kernels                    # 4096^2 per GPU
    run strong16384 --grid 16384 --strong --steps 4 --warmup 6 --no-cpu --no-kernels
fi
