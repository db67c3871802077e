#!/bin/bash
export DCAP_LOOP_DEBUG=1
timeout 300 python tools/loop_check.py --sizes 37,300,1000,8000 --time > gpurun_out/loop6.log 2>&1; echo rc=$?
DCAP_NO_GRAPHS=1 timeout 200 python tools/loop_trace_run.py gpurun_out/trace6.bin > gpurun_out/trace6.log 2>&1; echo rc=$?
for kv in DCAP_LOOP_NPF=0 DCAP_LOOP_SKEW=19 DCAP_LOOP_SKEW=25 DCAP_LOOP_PFENCE=1; do
  env $kv timeout 300 python tools/loop_check.py --sizes 8000 --time > gpurun_out/loop6_$kv.log 2>&1
done
unset DCAP_LOOP_DEBUG
b() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-sub --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline'])"; }
b DCAP_GREEDY_LOOP=1 > gpurun_out/loop6_bench.log 2>&1
b DCAP_GREEDY_LOOP=0 >> gpurun_out/loop6_bench.log 2>&1
