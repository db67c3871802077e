#!/bin/bash
# persistent greedy-loop kernel: its GPU test, ncu launch list of one captions step and a full capture of the kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_greedy_loop_gpu.py -q -m gpu > gpurun_out/r2b_looptest.log 2>&1; tail -3 gpurun_out/r2b_looptest.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-sub"
DCAP_NO_GRAPHS=1 timeout 300 $CMD > gpurun_out/r2b_plain_cap.log 2>&1 && \
DCAP_NO_GRAPHS=1 timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2b_launches_cap.csv $CMD > gpurun_out/r2b_ncu_cap.log 2>&1
DCAP_NO_GRAPHS=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:greedy_loop -s 4 -c 1 -o gpurun_out/r2b_loop_full -f $CMD > gpurun_out/r2b_ncu_loop.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r2b_launches_*.csv 2>/dev/null | tail -4
