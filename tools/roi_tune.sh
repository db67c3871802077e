#!/bin/bash
# A/B sweep of the ROIAlign forward paths (run on the GPU box): register-gather (round 1) vs shared-memory ring.
out=${1:-gpurun_out/roi_tune.log}
: > $out
run() { echo "== $*" >> $out; env "$@" timeout 180 python bench.py --workload roi_features --steps 30 --warmup 5 --no-e2e >> $out 2>&1; }
run DCAP_ROI_PATH=0
for ctas in 1 2 3 4; do
  run DCAP_ROI_PATH=1 DCAP_ROI_CTAS=$ctas
done
run DCAP_ROI_PATH=1 DCAP_ROI_CTAS=2 DCAP_ROI_DIAG=1
run DCAP_ROI_PATH=1 DCAP_ROI_CTAS=2 DCAP_ROI_DIAG=2
run DCAP_ROI_PATH=1 DCAP_ROI_CTAS=2 DCAP_ROI_DIAG=3
run DCAP_ROI_PATH=1 DCAP_ROI_CTAS=2 DCAP_ROI_DIAG=4
run DCAP_ROI_PATH=1 DCAP_ROI_CTAS=2 DCAP_ROI_DIAG=7
run DCAP_ROI_PATH=1 DCAP_ROI_CTAS=1 DCAP_ROI_DIAG=7
run DCAP_ROI_PATH=1 DCAP_ROI_CTAS=1 DCAP_ROI_DIAG=5
cat $out
