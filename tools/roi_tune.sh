#!/bin/bash
# A/B sweep of the ROIAlign forward paths (run on the GPU box): register-gather (round 1) vs shared-memory ring.
out=${1:-gpurun_out/roi_tune.log}
: > $out
run() { echo "== $*" >> $out; env "$@" timeout 180 python bench.py --workload roi_features --steps 30 --warmup 5 --no-e2e >> $out 2>&1; }
run DCAP_ROI_PATH=0
for dt in f32; do
for ctas in 1 2 3; do
  for warps in 7 14; do
    run DCAP_ROI_PATH=1 DCAP_ROI_CTAS=$ctas DCAP_ROI_WARPS=$warps
  done
done
done
run DCAP_ROI_PATH=1 DCAP_ROI_CTAS=2 DCAP_ROI_RING=4
run DCAP_ROI_PATH=1 DCAP_ROI_CTAS=2 DCAP_ROI_RING=5
run DCAP_ROI_PATH=1 DCAP_ROI_CTAS=1 DCAP_ROI_RING=8
cat $out
