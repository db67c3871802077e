#!/bin/bash
# GPU window C (1 GPU): the whole GPU suite, the default bench line, ncu launch lists + full captures for profiles/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_gputests4.log 2>&1; tail -8 gpurun_out/r2_gputests4.log
timeout 900 python bench.py > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; tail -c 1500 gpurun_out/r2_bench2.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench2_ref.json 2>&1; tail -c 600 gpurun_out/r2_bench2_ref.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
# one step of the captions workload, every kernel visible (graphs off): warm-up 3 steps + first timed step
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-sub"
DCAP_NO_GRAPHS=1 timeout 300 $CMD > gpurun_out/r2_plain_cap.log 2>&1 && \
DCAP_NO_GRAPHS=1 timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches_cap.csv $CMD > gpurun_out/r2_ncu_cap.log 2>&1
CMDT="python bench.py --workload train --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $CMDT > gpurun_out/r2_plain_train.log 2>&1 && \
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches_train.csv $CMDT > gpurun_out/r2_ncu_train.log 2>&1
CMDR="python bench.py --workload roi_features --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $CMDR > gpurun_out/r2_plain_roi.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:roi_ -s 12 -c 3 -o gpurun_out/r2_roi_full -f $CMDR > gpurun_out/r2_ncu_roi.log 2>&1
DCAP_NO_GRAPHS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 330 -c 5 -o gpurun_out/r2_gemm_full -f $CMD > gpurun_out/r2_ncu_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches_*.csv 2>/dev/null | tail -8
