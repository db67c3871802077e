#!/bin/bash
# per-launch list of one cfg3 training step (ncu, serialised) + the small-batch step with the dispatcher's own choice
mkdir -p gpurun_out
for B in 512 1024 4096; do
DCAP_TRAIN_BATCH=$B timeout 300 python bench.py --workload train --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2> gpurun_out/r2g_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); b=d['breakdown']
print('B=$B default step %.4f fb %.4f enqueue %.4f opt %.4f' % (d['ms_per_step'], b['forward_backward_ms'], b['forward_backward_host_enqueue_ms'], b['optimizer_ms']))"
done
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file gpurun_out/r2g_launches_train.csv \
  python bench.py --workload train --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2g_ncu_train.log 2>&1
echo ncu rc=$?
python tools/ncu_launches.py gpurun_out/r2g_launches_train.csv | head -30
