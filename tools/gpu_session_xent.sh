#!/bin/bash
# fused softmax / cross-entropy / bias-gradient kernel: parity tests, A/B of the cfg3 step, kernel duration under ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -x -q -m gpu > gpurun_out/r2h_tests.log 2>&1; tail -3 gpurun_out/r2h_tests.log
for F in 1 0 1 0; do
DCAP_XENT_FUSED=$F timeout 300 python bench.py --workload train --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2> gpurun_out/r2h_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); b=d['breakdown']
print('fused=$F step %.4f fb %.4f opt %.4f loss %s %s' % (d['ms_per_step'], b['forward_backward_ms'], b['optimizer_ms'], d['config']['loss_first'], d['config']['loss_last']))"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:softmax_xent_colsum --launch-skip 3 -c 1 \
  -o gpurun_out/r2h_xent_full -f python bench.py --workload train --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2h_ncu.log 2>&1
echo ncu rc=$?
