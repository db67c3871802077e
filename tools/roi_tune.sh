#!/bin/bash
# A/B sweep of the ROIAlign forward paths (run on the GPU box).
#   DCAP_ROI_PATH=0 round-1 prepare (one CTA per image) + register-gather stream kernel
#   DCAP_ROI_PATH=2 order kernel + wide record pre-pass + the same gather kernel
#   DCAP_ROI_PATH=3 order kernel + ring-record pre-pass + shared-memory ring kernel (bulk async copies)   [default]
#   DCAP_ROI_PATH=1 first ring kernel (producer computes the records itself)
out=${1:-gpurun_out/roi_tune.log}
: > $out
run() { echo "== $*" >> $out; env "$@" timeout 180 python bench.py --workload roi_features --steps 32 --warmup 5 --no-e2e 2>&1 | grep -v "^$" | tail -3 >> $out; }
run DCAP_ROI_PATH=0
run DCAP_ROI_PATH=2
run DCAP_ROI_PATH=3
run DCAP_ROI_PATH=3 DCAP_ROI_CTAS=1
run DCAP_ROI_PATH=3 DCAP_ROI_PROF=1
run DCAP_ROI_PATH=3 DCAP_ROI_PROF=1 DCAP_ROI_DIAG=7
run DCAP_ROI_PATH=3 DCAP_ROI_DIAG=2
run DCAP_ROI_PATH=3 DCAP_ROI_DIAG=4
run DCAP_ROI_PATH=3 DCAP_ROI_WARPS=14
run DCAP_ROI_PATH=3 DCAP_ROI_RING=5
cat $out
