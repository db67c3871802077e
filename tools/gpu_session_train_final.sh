#!/bin/bash
# training path after the round's last changes: tests, the cfg3 line with its e2e, one step's launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -x -q -m gpu > gpurun_out/r2m_tests.log 2>&1; tail -3 gpurun_out/r2m_tests.log
timeout 600 python bench.py --workload train --no-cpu-baseline > gpurun_out/r2m_train.json 2> gpurun_out/r2m_err.log; echo rc=$?
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2m_train.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e'], d['breakdown'])
P
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --launch-skip 150 -c 420 --csv --log-file gpurun_out/r2m_launches_train.csv \
  python bench.py --workload train --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2m_ncu.log 2>&1
echo ncu rc=$?
