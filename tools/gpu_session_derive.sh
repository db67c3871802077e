#!/bin/bash
# batched operand re-derivation + split-K exemption of the small-launch rule: training / decoder tests, then the step at 512 / 4096 rows
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train_gpu.py tests/test_decoder_gpu.py tests/test_gemm_gpu.py -x -q -m gpu > gpurun_out/r2i_tests.log 2>&1; tail -3 gpurun_out/r2i_tests.log
for B in 512 4096 4096; do
DCAP_TRAIN_BATCH=$B timeout 300 python bench.py --workload train --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2> gpurun_out/r2i_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); b=d['breakdown']
print('B=$B step %.4f fb %.4f enqueue %.4f opt %.4f' % (d['ms_per_step'], b['forward_backward_ms'], b['forward_backward_host_enqueue_ms'], b['optimizer_ms']))"
done
