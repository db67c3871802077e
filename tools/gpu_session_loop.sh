#!/bin/bash
export DCAP_LOOP_DEBUG=1
timeout 300 python tools/loop_check.py --sizes 37,300,1000,8000 --time > gpurun_out/loop7.log 2>&1; echo rc=$?
DCAP_NO_GRAPHS=1 timeout 200 python tools/loop_trace_run.py gpurun_out/trace7.bin > gpurun_out/trace7.log 2>&1; echo rc=$?
for kv in DCAP_LOOP_AHEAD=2 DCAP_LOOP_NPF=1 DCAP_LOOP_SKEW=25 DCAP_LOOP_AHEAD=1; do
  env $kv timeout 300 python tools/loop_check.py --sizes 8000 --time > gpurun_out/loop7_$kv.log 2>&1
done
DCAP_LOOP_AHEAD=2 DCAP_NO_GRAPHS=1 timeout 200 python tools/loop_trace_run.py gpurun_out/trace7_a2.bin > gpurun_out/trace7_a2.log 2>&1
