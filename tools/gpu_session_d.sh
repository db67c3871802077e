#!/bin/bash
# GPU window D (2 GPUs): NCCL data-parallel training test with the real model, and both bench arms under torchrun as the driver launches them.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parallel_gpu.py -q 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2_bench_n2_ref.json 2> gpurun_out/r2_bench_n2_ref.err; tail -c 700 gpurun_out/r2_bench_n2_ref.json
NCCL_DEBUG=WARN timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -c 1500 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n2.json").read().strip().splitlines()[-1])
print("captions N=2", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["frac_of_host_copy_ceiling"], d["e2e"]["h2d_gbs_per_rank_all_ranks_copying"])
for k, v in d.get("workloads", {}).items():
    print(k, {kk: v.get(kk) for kk in ("value", "ms_per_step", "error")}, (v.get("e2e") or {}).get("value"), v.get("breakdown"))
PY
