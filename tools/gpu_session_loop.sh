#!/bin/bash
# persistent greedy-loop kernel: correctness A/B against the launch-per-GEMM path, timing, one trace, knob sweep
export DCAP_LOOP_DEBUG=1
timeout 300 python tools/loop_check.py --sizes 300,37,1000,8000 --time > gpurun_out/loop4.log 2>&1; echo rc=$?
DCAP_NO_GRAPHS=1 timeout 200 python tools/loop_trace_run.py gpurun_out/trace4.bin > gpurun_out/trace4.log 2>&1; echo rc=$?
for kv in DCAP_LOOP_FOLD=0 DCAP_LOOP_PFENCE=1 DCAP_LOOP_PFENCE=0 DCAP_LOOP_EPIACQ=1 DCAP_LOOP_SKEW=26 DCAP_LOOP_SKEW=13; do
  env $kv timeout 300 python tools/loop_check.py --sizes 8000 --time > gpurun_out/loop4_$kv.log 2>&1
done
