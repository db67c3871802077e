#!/bin/bash
# A/B sweep of the ROIAlign forward paths (run on the GPU box).
#   DCAP_ROI_PATH=0 round-1 prepare (one CTA per image) + register-gather stream kernel
#   DCAP_ROI_PATH=2 order kernel + wide record pre-pass + the same gather kernel (DCAP_ROI_VARIANT: 1/2/4 = in-register tap re-use)
#   DCAP_ROI_PATH=3 order kernel + ring-record pre-pass + shared-memory ring kernel (bulk async copies)
#   DCAP_ROI_PATH=1 first ring kernel (producer computes the records itself)
out=${1:-gpurun_out/roi_tune.log}
: > $out
run() { echo "== $*" >> $out; env "$@" timeout 180 python bench.py --workload roi_features --steps 32 --warmup 5 --no-e2e 2>&1 | grep -v "^$" | tail -3 >> $out; }
run DCAP_ROI_PATH=2
for v in 1 2 3 4; do run DCAP_ROI_PATH=2 DCAP_ROI_VARIANT=$v; done
for c in 3 5 6 8; do run DCAP_ROI_PATH=2 DCAP_ROI_CTAS=$c; done
run DCAP_ROI_PATH=2 DCAP_ROI_VARIANT=1 DCAP_ROI_CTAS=5
run DCAP_ROI_PATH=2 DCAP_ROI_VARIANT=2 DCAP_ROI_CTAS=2
run DCAP_ROI_PATH=2 DCAP_ROI_VARIANT=2 DCAP_ROI_CTAS=3
run DCAP_ROI_PATH=3
run DCAP_ROI_PATH=3 DCAP_ROI_ONLY_NEW=1
run DCAP_ROI_PATH=3 DCAP_ROI_PROD=2
run DCAP_ROI_PATH=3 DCAP_ROI_PROD=2 DCAP_ROI_ONLY_NEW=1
run DCAP_ROI_PATH=3 DCAP_ROI_PROD=2 DCAP_ROI_ONLY_NEW=1 DCAP_ROI_DIAG=7
run DCAP_ROI_PATH=3 DCAP_ROI_CTAS=1 DCAP_ROI_PROD=2 DCAP_ROI_GROUPS=2 DCAP_ROI_ONLY_NEW=1
run DCAP_ROI_PATH=3 DCAP_ROI_CTAS=1 DCAP_ROI_PROD=2 DCAP_ROI_GROUPS=2 DCAP_ROI_ONLY_NEW=1 DCAP_ROI_PROF=1
run DCAP_ROI_PATH=3 DCAP_ROI_CTAS=1 DCAP_ROI_PROD=1 DCAP_ROI_GROUPS=2 DCAP_ROI_ONLY_NEW=1
cat $out
