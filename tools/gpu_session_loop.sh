#!/bin/bash
r() { echo "== $*"; env "$@" timeout 200 python bench.py --workload roi_features --steps 30 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | python tools/bench_line.py; }
for kv in DCAP_ROI_ADJ=0 DCAP_ROI_ADJ=4 DCAP_ROI_ADJ=2 "DCAP_ROI_ADJ=4 DCAP_ROI_CTAS=8" "DCAP_ROI_ADJ=2 DCAP_ROI_CTAS=8" "DCAP_ROI_ADJ=8 DCAP_ROI_CTAS=8" "DCAP_ROI_ADJ=0 DCAP_ROI_CTAS=8" DCAP_ROI_ADJ=0; do r $kv; done > gpurun_out/roi_adj.log 2>&1
cat gpurun_out/roi_adj.log
