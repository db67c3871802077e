#!/bin/bash
# sharded optimiser: 2-GPU NCCL test, train workload A/B at N GPUs
N=${1:-2}
mkdir -p gpurun_out
[ -z "$SKIP_TEST" ] && timeout 900 python -m pytest tests/test_parallel_gpu.py -q -m gpu 2>&1 | tail -3
for so in ${SO_LIST:-1 0 1 0}; do
  echo "== DCAP_SHARD_OPT=$so N=$N"
  DCAP_SHARD_OPT=$so timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$so bench.py --workload train --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | python tools/bench_line.py
done
