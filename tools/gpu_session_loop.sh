#!/bin/bash
export DCAP_LOOP_DEBUG=1
timeout 300 python tools/loop_check.py --sizes 37,300,1000,2500,8000 --time > gpurun_out/loop12.log 2>&1; echo rc=$?
DCAP_NO_GRAPHS=1 timeout 200 python tools/loop_trace_run.py gpurun_out/trace12.bin > gpurun_out/trace12.log 2>&1; echo rc=$?
for kv in DCAP_LOOP_SKEW=26 DCAP_LOOP_DEFER=0 DCAP_LOOP_SKEW=21; do
  env $kv timeout 300 python tools/loop_check.py --sizes 8000 --time > gpurun_out/loop12_$kv.log 2>&1
done
