#!/bin/bash
# One GPU window: parity of the new ROIAlign paths, the ROIAlign sweep, the whole GPU suite, decoder lane A/B, precision study.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_roi_align_gpu.py -q -x 2>&1 | tail -3
DCAP_ROI_PATH=2 timeout 400 python -m pytest tests/test_roi_align_gpu.py -q -x 2>&1 | tail -3
tools/roi_tune.sh gpurun_out/r2_roi_tune4.log > /dev/null; cat gpurun_out/r2_roi_tune4.log | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_gputests2.log 2>&1; tail -25 gpurun_out/r2_gputests2.log
for lanes in 1 2; do
  echo "== captions DCAP_LANES=$lanes"; DCAP_LANES=$lanes timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['decoder_ms'], d['roofline']['frac'], d['roofline_hbm']['roi_align_ms'], d['roofline_hbm']['frac'])"
done
echo "== captions DCAP_LANES=2 DCAP_CELL_TMA=0"; DCAP_LANES=2 DCAP_CELL_TMA=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['decoder_ms'], d['roofline']['frac'])"
echo "== captions DCAP_LANES=1 DCAP_CELL_TMA=0"; DCAP_LANES=1 DCAP_CELL_TMA=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['decoder_ms'], d['roofline']['frac'])"
timeout 600 python tools/bf16_error_study.py 2>&1 | tail -4
