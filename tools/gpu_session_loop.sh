#!/bin/bash
DCAP_NO_GRAPHS=1 timeout 200 python tools/loop_trace_run.py gpurun_out/trace8.bin > gpurun_out/trace8.log 2>&1; echo rc=$?
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_gemm_gpu.py tests/test_postprocess.py tests/test_threading_gpu.py -x -q -m gpu > gpurun_out/loop8_tests.log 2>&1; echo rc=$?
tail -5 gpurun_out/loop8_tests.log
