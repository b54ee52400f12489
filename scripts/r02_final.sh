#!/bin/bash
# final round-2 validation after the one-kernel scalar transport: the whole GPU suite, the headline
# bench (regression check), the two configurations the transport kernel serves, the smoke hook
set -x
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q --durations=8 > $O/r02_gpu_tests.txt 2>&1
tail -6 $O/r02_gpu_tests.txt
python bench.py --steps 20 --warmup 6 > $O/r02_bench_n1_4096.json 2> $O/r02_bench_n1_4096.err
python bench.py --config rsw8192 --steps 10 --warmup 6 > $O/r02_bench_rsw8192.json 2> $O/r02_bench_rsw8192.err
timeout 420 python bench.py --config bouss16384 --steps 3 --warmup 5 --no-cpu > $O/r02_bench_bouss16384_n1.json 2> $O/r02_bench_bouss16384_n1.err
tail -c 400 $O/r02_bench_bouss16384_n1.err
python __graft_entry__.py --smoke > $O/r02_smoke.txt 2>&1
tail -3 $O/r02_smoke.txt
cat $O/r02_bench_n1_4096.json $O/r02_bench_rsw8192.json $O/r02_bench_bouss16384_n1.json | cut -c 1-400
