#!/bin/bash
DCAP_NO_GRAPHS=1 timeout 200 python tools/loop_trace_run.py gpurun_out/trace13.bin > gpurun_out/trace13.log 2>&1; echo rc=$?
