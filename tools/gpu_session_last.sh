#!/bin/bash
# last check of the round: the suites that exercise the re-layout kernel and fit_generator's queue, the training line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_train_gpu.py tests/test_greedy_loop_gpu.py -x -q -m gpu > gpurun_out/r2y_tests.log 2>&1; tail -3 gpurun_out/r2y_tests.log
timeout 300 python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2y_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); b=d['breakdown']
print('train step %.4f fb %.4f opt %.4f e2e %.0f' % (d['ms_per_step'], b['forward_backward_ms'], b['optimizer_ms'], d['e2e']['value']))"
