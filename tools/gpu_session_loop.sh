#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_gemm_gpu.py tests/test_train_gpu.py -q -m gpu > gpurun_out/r2d_tests.log 2>&1; tail -3 gpurun_out/r2d_tests.log
b() { echo "== $*"; env "$@" timeout 400 python bench.py --workload beam --steps 3 --warmup 1 --no-e2e --no-cpu-baseline 2>/dev/null | python tools/bench_line.py; }
b DCAP_BEAM_BLOCKED=1 > gpurun_out/r2d_beam_ab.log 2>&1
b DCAP_BEAM_BLOCKED=0 >> gpurun_out/r2d_beam_ab.log 2>&1
b DCAP_BEAM_BLOCKED=1 >> gpurun_out/r2d_beam_ab.log 2>&1
cat gpurun_out/r2d_beam_ab.log
