#!/bin/bash
# persistent greedy-loop kernel: GPU tests, bench A/B against the launch-per-GEMM path
mkdir -p gpurun_out
DCAP_LOOP_DEBUG=1 timeout 300 python tools/dbg_train_loop.py 2>&1 | tail -12
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2c_tests.log 2>&1; tail -4 gpurun_out/r2c_tests.log
b() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-sub --no-cpu-baseline 2>/dev/null | python tools/bench_line.py; }
b DCAP_GREEDY_LOOP=1 > gpurun_out/r2c_bench_ab.log 2>&1
b DCAP_GREEDY_LOOP=0 >> gpurun_out/r2c_bench_ab.log 2>&1
b DCAP_GREEDY_LOOP=1 >> gpurun_out/r2c_bench_ab.log 2>&1
cat gpurun_out/r2c_bench_ab.log
