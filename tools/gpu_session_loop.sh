#!/bin/bash
export DCAP_LOOP_DEBUG=1
timeout 600 python tools/loop_check.py --sizes 1,37,300,600,1000,1300,1800,2500,4000,8000 --time > gpurun_out/loop15.log 2>&1; echo rc=$?
grep "ms per call\|agreement\|LOOP_CHECK" gpurun_out/loop15.log
