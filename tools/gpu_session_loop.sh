#!/bin/bash
# persistent greedy-loop kernel: bench contract test, ncu launch list + full capture, default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bench_contract.py -q -m gpu > gpurun_out/r2c_contract.log 2>&1; tail -3 gpurun_out/r2c_contract.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-sub"
DCAP_NO_GRAPHS=1 timeout 300 $CMD > gpurun_out/r2c_plain_cap.log 2>&1 && \
DCAP_NO_GRAPHS=1 timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2c_launches_cap.csv $CMD > gpurun_out/r2c_ncu_cap.log 2>&1
DCAP_NO_GRAPHS=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:greedy_loop -s 4 -c 1 -o gpurun_out/r2c_loop_full -f $CMD > gpurun_out/r2c_ncu_loop.log 2>&1
timeout 900 python bench.py > gpurun_out/r2c_bench_full.json 2> gpurun_out/r2c_bench_full.err; echo rc=$?
python tools/bench_line.py < gpurun_out/r2c_bench_full.json
