#!/bin/bash
# cfg5 (5000 images x 300 RoIs): images per caption_rois call
mkdir -p gpurun_out
for nb in 8 16 24 32; do
DCAP_VG_BATCH=$nb timeout 300 python bench.py --workload captions_vg --steps 2 --warmup 1 --no-cpu-baseline 2> gpurun_out/r2k_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('VG_BATCH=$nb value %.0f ms %.1f e2e %.0f clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz']))"
done
