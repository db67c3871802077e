#!/bin/bash
# 2 GPUs: NCCL tests of the data-parallel trainer, then the default line (with its sub-records) as the driver launches it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parallel_gpu.py -q -m gpu 2>&1 | tail -3
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2n_bench_n2.json 2> gpurun_out/r2n_bench_n2.err; tail -c 400 gpurun_out/r2n_bench_n2.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2n_bench_n2.json").read().strip().splitlines()[-1])
print("captions N=2", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["frac_of_host_copy_ceiling"], d["e2e"]["h2d_gbs_per_rank_all_ranks_copying"], d["clocks"])
for k, v in d.get("workloads", {}).items():
    print(k, {kk: v.get(kk) for kk in ("value", "ms_per_step", "error")}, (v.get("e2e") or {}).get("value"), v.get("breakdown"))
PY
