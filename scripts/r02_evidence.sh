#!/bin/bash
# gpurun calls that produce the round-2 evidence under gpurun_out/ (copied to profiles/ afterwards).
#   ./scripts/r02_evidence.sh bench     tests + bench lines (1 GPU)
#   ./scripts/r02_evidence.sh ncu       launch list + ncu --set full of the hot kernels, exported to CSV on the box
#                                       (gpurun_out/ is only copied back below 64 MiB: the .ncu-rep stays there)
set -x
O=gpurun_out
mkdir -p $O
if [ "$1" = "bench" ]; then
  (python scripts/cpu_rows.py 512 1024 > $O/r02_cpu_rows.jsonl 2> $O/r02_cpu_rows.err &)
  python -m pytest tests -m gpu -q --durations=8 > $O/r02_gpu_tests.txt 2>&1
  tail -5 $O/r02_gpu_tests.txt
  python bench.py --steps 20 --warmup 6 > $O/r02_bench_n1_4096.json 2> $O/r02_bench_n1_4096.err
  python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err
  for c in vortex rsw8192 qgrsw8192; do python bench.py --config $c --steps 10 --warmup 6 > $O/r02_bench_$c.json 2> $O/r02_bench_$c.err; done
  timeout 420 python bench.py --config bouss16384 --steps 3 --warmup 5 --no-cpu --no-kernels > $O/r02_bench_bouss16384.json 2> $O/r02_bench_bouss16384.err
  tail -c 600 $O/r02_bench_bouss16384.err
  sleep 100   # the 1024^2 CPU row
else
  ncu --metrics gpu__time_duration.sum --clock-control none -s 2700 -c 520 --csv --log-file $O/r02_launches_bench_4096.csv python bench.py --steps 3 --warmup 6 --no-cpu --no-kernels > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "f2dprof/" -o /tmp/r02_ncu_full python scripts/kprof.py > $O/r02_ncu_full.log 2>&1
  python scripts/ncu_summary.py /tmp/r02_ncu_full.ncu-rep > $O/r02_ncu_full.csv
  for k in k_mg_up k_mg_down k_cg_update_p k_cg_dir_apply k_stage_tma k_diag_tma; do
    ncu -i /tmp/r02_ncu_full.ncu-rep --page source --csv --kernel-name regex:$k 2>/dev/null | python scripts/ncu_source_top.py 40 > $O/r02_ncu_source_$k.txt
  done
fi
ls -la $O | tail -20
