#!/bin/bash
# persistent greedy-loop kernel: correctness A/B against the launch-per-GEMM path, timing, one trace, knob sweep
export DCAP_LOOP_DEBUG=1
timeout 300 python tools/loop_check.py --sizes 300,37,1000,8000 --time > gpurun_out/loop3.log 2>&1; echo rc=$?
DCAP_NO_GRAPHS=1 timeout 200 python tools/loop_trace_run.py gpurun_out/trace3.bin > gpurun_out/trace3.log 2>&1; echo rc=$?
for kv in DCAP_LOOP_WPF=1 DCAP_LOOP_L2PF=0 DCAP_LOOP_SKEW=12 DCAP_LOOP_SKEW=26; do
  env $kv timeout 300 python tools/loop_check.py --sizes 8000 --time > gpurun_out/loop3_$kv.log 2>&1
done
