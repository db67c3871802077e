#!/bin/bash
# A/B sweep of the ROIAlign forward paths (run on the GPU box): register-gather (round 1) vs shared-memory ring.
out=${1:-gpurun_out/roi_tune.log}
: > $out
run() { echo "== $*" >> $out; env "$@" timeout 180 python bench.py --workload roi_features --steps 32 --warmup 5 --no-e2e 2>&1 | grep -v "^$" | tail -4 >> $out; }
run DCAP_ROI_PATH=0
for sync in 0 1 2 3; do
  run DCAP_ROI_PATH=1 DCAP_ROI_CTAS=2 DCAP_ROI_SYNC=$sync DCAP_ROI_PROF=1
  run DCAP_ROI_PATH=1 DCAP_ROI_CTAS=2 DCAP_ROI_SYNC=$sync DCAP_ROI_PROF=1 DCAP_ROI_DIAG=7
done
run DCAP_ROI_PATH=1 DCAP_ROI_CTAS=1 DCAP_ROI_SYNC=3 DCAP_ROI_PROF=1
run DCAP_ROI_PATH=1 DCAP_ROI_CTAS=3 DCAP_ROI_SYNC=3 DCAP_ROI_PROF=1
run DCAP_ROI_PATH=1 DCAP_ROI_CTAS=2 DCAP_ROI_SYNC=3
cat $out
