#!/bin/bash
# tuning: the per-rank shapes of the N-rank training step on ONE GPU, with the CTA-pair kernels limited to launches of
# at least n pair tiles (DCAP_2CTA=n), and the host's enqueue time next to the device time
mkdir -p gpurun_out
out=gpurun_out/r2f_smallm2.log; : > $out
for B in 2048 4096; do
  for T in 1 20 33 37 70; do
    DCAP_TRAIN_BATCH=$B DCAP_2CTA=$T timeout 300 python bench.py --workload train --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2> gpurun_out/r2f_err.log | \
      python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); b=d['breakdown']
print('B=$B 2CTA>=$T step %.4f fb %.4f enqueue %.4f opt %.4f' % (d['ms_per_step'], b['forward_backward_ms'], b['forward_backward_host_enqueue_ms'], b['optimizer_ms']))" >> $out 2>&1
  done
done
for T in 1 20 37; do
  echo "== greedy DCAP_2CTA=$T" >> $out
  DCAP_2CTA=$T timeout 300 python tools/loop_check.py --sizes 1000,300,2400 --time 2>&1 | grep -v Warning | tail -8 >> $out
done
cat $out
