#!/bin/bash
# GEMM knob A/B on the captions workload (device-timed decoder ms per 8000 RoIs)
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-sub --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['decoder_ms'], d['roofline']['frac'])"; }
run DCAP_X=0
run DCAP_TMA_STORE_MASK=15
run DCAP_TMA_STORE_MASK=7
run DCAP_TMA_STORE_MASK=3
run DCAP_X=0
