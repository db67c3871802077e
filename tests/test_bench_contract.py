"""bench.py contract (CPU part): the reference arm runs without a GPU and prints ONE JSON line with the
keys the driver reads; the GPU arm's line is checked by tests/test_bench_gpu.py."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def _run(args, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "roi_captions_per_sec" and d["unit"] == "RoI captions/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
def test_default_arm_line():
    d = _run(["--steps", "3", "--warmup", "3", "--e2e-steps", "1"])
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline", "cpu_baseline"} <= set(d)
    assert d["metric"] == "roi_captions_per_sec" and d["n_gpus"] == 1 and d["vs_baseline"] is None
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "tensor"
    assert 0.2 < r["frac"] < 1.0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert d["e2e"]["h2d_bytes_per_step"] > 7e8 and d["e2e"]["d2h_bytes_per_step"] == 8000 * 15 * 4
    assert d["e2e"]["token_agreement_with_device_run"] == 1.0
    # per step: ROIAlign 3 + head 2 + merged hoisted-term GEMM 1 + token fill 1 + embedding gather 1 + the persistent greedy loop 1
    # (profiles/r2_launches_captions_loop.txt)
    assert d["gpu_launches"] == 3 * 9 and d["config"]["public_api_matches"] is True
    assert d["config"]["token_agreement_with_launch_per_gemm_path"] == 1.0
    assert d["cpu_baseline"]["value"] > 0 and d["cpu_baseline"]["cores"] >= 1


def test_l2_path_helper():
    """roofline_hbm.l2_path: taps + stores against the L2 -> SM fabric ceiling (6300 B per clock at the sampled SM clock)."""
    sys.path.insert(0, ROOT)
    import bench
    r = bench.l2_path(tap_pixels=4 * 49 * 8000, out_bytes=2 * 256 * 49 * 8000, ms=0.16, sm_mhz=1965.0)
    assert r["tap_bytes"] == 4 * 49 * 8000 * 4 * 256 and r["out_bytes"] == 2 * 256 * 49 * 8000
    assert abs(r["cap_tbs"] - 6300 * 1965e6 / 1e12) < 0.01
    assert abs(r["achieved_tbs"] - (r["tap_bytes"] + r["out_bytes"]) / 0.16e-3 / 1e12) < 0.01
    assert abs(r["frac"] - r["achieved_tbs"] / r["cap_tbs"]) < 2e-3
    assert bench.l2_path(1, 1, 1.0, None)["cap_tbs"] == round(6300 * 1965e6 / 1e12, 2)      # no clock sample: the part's maximum
