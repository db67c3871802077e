#!/bin/bash
mkdir -p gpurun_out
for cfg in "DCAP_ROI_PATH=2 DCAP_ROI_VARIANT=1" "DCAP_ROI_PATH=2 DCAP_ROI_VARIANT=4" "DCAP_ROI_PATH=3 DCAP_ROI_PROD=2 DCAP_ROI_ONLY_NEW=1" "DCAP_ROI_PATH=3 DCAP_ROI_CTAS=1 DCAP_ROI_PROD=2 DCAP_ROI_GROUPS=2 DCAP_ROI_ONLY_NEW=1"; do
  echo "== roi tests with $cfg"; env $cfg timeout 400 python -m pytest tests/test_roi_align_gpu.py -q -x 2>&1 | tail -2
done
tools/roi_tune.sh gpurun_out/r2_roi_tune5.log > /dev/null; cat gpurun_out/r2_roi_tune5.log | cut -c1-330
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_gputests3.log 2>&1; tail -12 gpurun_out/r2_gputests3.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; tail -c 3000 gpurun_out/r2_bench1.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench1.json").read().strip().splitlines()[-1])
    print("captions", d["value"], d["ms_per_step"], "roofline", d["roofline"]["frac"], "hbm", d["roofline_hbm"]["frac"], "e2e", d["e2e"]["value"], d["e2e"].get("frac_of_host_copy_ceiling"))
    for k, v in d.get("workloads", {}).items():
        print(k, {kk: v.get(kk) for kk in ("value", "ms_per_step", "error")}, (v.get("roofline") or {}).get("frac"), (v.get("e2e") or {}).get("value"), v.get("breakdown"))
except Exception as e:
    print("bench parse failed", e)
PY
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
