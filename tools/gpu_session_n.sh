#!/bin/bash
# usage: tools/gpu_session_n.sh N  -- both bench arms under torchrun at N GPUs, as the driver launches them
N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r2b_bench_n${N}_ref.json 2> gpurun_out/r2b_bench_n${N}_ref.err; tail -c 300 gpurun_out/r2b_bench_n${N}_ref.json
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2b_bench_n${N}.json 2> gpurun_out/r2b_bench_n${N}.err; tail -c 800 gpurun_out/r2b_bench_n${N}.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2b_bench_n${N}.json").read().strip().splitlines()[-1])
print("captions N=${N}", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["frac_of_host_copy_ceiling"], d["e2e"]["h2d_gbs_per_rank_all_ranks_copying"], d["e2e"]["numa"], d["clocks"])
for k, v in d.get("workloads", {}).items():
    print(k, {kk: v.get(kk) for kk in ("value", "ms_per_step", "error")}, (v.get("e2e") or {}).get("value"), v.get("breakdown"))
PY
