#!/bin/bash
export DCAP_LOOP_DEBUG=1
timeout 600 python tools/loop_check.py --sizes 37,300,1000,2500,8000 --time > gpurun_out/loop16.log 2>&1; echo rc=$?
grep "loop=2\|agreement\|LOOP_CHECK" gpurun_out/loop16.log
timeout 600 python tools/loop_fuzz.py --cases 40 --seed 3 2>&1 | tail -2
DCAP_NO_GRAPHS=1 timeout 200 python tools/loop_trace_run.py gpurun_out/trace15.bin > gpurun_out/trace15.log 2>&1; echo rc=$?
