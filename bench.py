#!/usr/bin/env python
"""Benchmark of the per-RoI captioning hot path (contract in the task prompt, section 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload captions|roi_features|train|beam|proposals]

One "step" = one pass of the hot path over one batch of synthetic input.  Workloads:

  captions      (default) the BASELINE metric "RoI captions/sec": per GPU 8 images x 1000 RoIs
                (configs[1]'s RoI stage: FPN P2-P5 of a 1024x1024 image, 256 ch fp32) ->
                PyramidROIAlign -> RoI head -> v1 inject-LSTM greedy decoding (configs[0]/[4]'s
                decoder: hidden 512, vocab 10k, embedding 300, P = 15) -> [8000, 15] token ids.
                bf16 tensor-core decoder, fp32 ROIAlign arithmetic.
  roi_features  BASELINE.json configs[1] alone: 8 images x 1000 RoIs -> [8000, 7, 7, 256] fp32.
  train         configs[2]: one v1 decoder training step (bf16, global batch 4096 x P=16, AMSGrad), data parallel.
  beam          configs[3]: width-3 beam decoding of 100k pre-extracted RoI feature vectors.
  proposals     SURVEY.md 8f rank 3: ProposalLayer on 32 images x 261888 anchors per GPU (top 6000 -> NMS -> 1000).

With N GPUs every rank owns its own 8 images (weak scaling, images sharded, no collective).

`value` is measured with inputs resident in HBM; `e2e` goes through the host-buffer C-ABI entry
point with pinned host buffers (H2D of the pyramid and D2H of the features inside the timed
region).  `--impl reference` times the CPU restatement of the reference (oracle/c, all host
threads) on a bounded sample of the same workload -- the reference itself is TensorFlow-1.x
Python and cannot run here.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU reference arm (rank 0 only) must use the host's cores,
# so the thread count is set explicitly BEFORE numpy / OpenBLAS / libgomp are loaded and recorded in the line.
if "reference" in sys.argv[1:]:
    _n = str(os.cpu_count() or 1)
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = _n

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

POOL = (7, 7)
IMAGE_SHAPE = (1024, 1024, 3)
CHANNELS = 256
IMAGES_PER_GPU = 8
ROIS_PER_IMAGE = 1000
SEED_CFG2 = 1002           # 1000 + config index
SEED_DECODER = 1005
VOCAB, EMBED, UNITS, PADDING = 10000, 300, 512, 15
# head 27.79 + hoisted 6.29 MFLOP once, 29.05 MFLOP per step, P = 15
FLOP_PER_ROI_GREEDY = 27787264 + 4194304 + 2097152 + PADDING * (1228800 + 2097152 + 4194304 + 1048576 + 20480000)


# --------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------

def measured_peaks(key="hbm_gbs"):
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p[key]), "measured (MEASURED_PEAKS.json %s)" % key
    fallback = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
    return fallback[key], "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons while the timed region runs: NVML in-process every 5 ms (the timed regions
    are 0.1-1 s long; spawning nvidia-smi takes ~50 ms per sample), nvidia-smi as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []            # (sm MHz, max MHz, set of reasons)
        self.stop_flag = threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and all(x.strip().isdigit() for x in visible.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.bits = {"hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown,
                         "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                         "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                         "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap}
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        self.samples.append((mhz, self.max_mhz, {k for k, b in self.bits.items() if mask & b}))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            f = [x.strip() for x in out.split(",")]
            self.samples.append((float(f[0]), float(f[1]), {n for n, v in zip(self.NAMES, f[2:6]) if v.lower().startswith("active")}))

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                if self.nvml is not None:
                    self.nvml = None                     # fall back to nvidia-smi
            self.stop_flag.wait(0.005 if self.nvml is not None else 0.1)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = [s[0] for s in self.samples]
        mx = [s[1] for s in self.samples]
        reasons = set()
        for s in self.samples:
            reasons |= s[2]
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def tap_unique_pixels(boxes, levels, fm_shapes, pool):
    """T_unique of SURVEY.md section 8(d): distinct (image, level, y, x) pixels touched by any
    in-range bilinear tap.  `levels` come from the GPU kernel; this is byte accounting for the
    roofline, not part of the product path."""
    B, N = boxes.shape[:2]
    fb = boxes.reshape(-1, 4).astype(np.float32)
    lv = levels.reshape(-1)
    img = np.repeat(np.arange(B), N)
    ph, pw = pool
    total = 0
    f32 = np.float32
    for i in range(4):
        sel = np.nonzero(lv == i + 2)[0]
        if sel.size == 0:
            continue
        H, W = fm_shapes[i]
        bx = fb[sel]
        Hm1, Wm1 = f32(H - 1), f32(W - 1)
        hs = ((bx[:, 2] - bx[:, 0]) * Hm1) / f32(ph - 1)
        ws = ((bx[:, 3] - bx[:, 1]) * Wm1) / f32(pw - 1)
        in_y = ((bx[:, 0] * Hm1)[:, None] + np.arange(ph, dtype=f32)[None] * hs[:, None]).astype(f32)
        in_x = ((bx[:, 1] * Wm1)[:, None] + np.arange(pw, dtype=f32)[None] * ws[:, None]).astype(f32)
        y_ok = (in_y >= 0) & (in_y <= Hm1)
        x_ok = (in_x >= 0) & (in_x <= Wm1)
        ys = np.stack([np.floor(in_y), np.ceil(in_y)], -1).astype(np.int64)
        xs = np.stack([np.floor(in_x), np.ceil(in_x)], -1).astype(np.int64)
        shape = (len(sel), ph, 2, pw, 2)
        ok = np.broadcast_to(y_ok[:, :, None, None, None] & x_ok[:, None, None, :, None], shape)
        yy = np.broadcast_to(ys[:, :, :, None, None], shape)
        xx = np.broadcast_to(xs[:, None, None, :, :], shape)
        ii = np.broadcast_to(img[sel][:, None, None, None, None], shape)
        total += np.unique((ii[ok] * H + yy[ok]) * W + xx[ok]).size
    return int(total)



def l2_path(tap_pixels, out_bytes, ms, sm_mhz):
    """Why ROIAlign stops short of the HBM roofline: every bilinear tap (4 per sample, 196 per RoI) crosses the L2 -> SM
    fabric even when DRAM serves each distinct pixel once (overlapping RoIs share taps 3.5x at cfg2), and the stores cross
    it too.  The fabric's measured ceiling is ~6300 B per clock for the whole chip whatever the access path
    (/opt/skills/guides/B300_MICROARCH.md, "LTS throughput cap"; measured there on B300 -- the same L2 slice count).
    Returns that traffic (before the ~13 % the L1 filters, profiles/r2_roi_align_full.txt) against the cap at the SM clock
    sampled during the run."""
    tap_bytes = int(tap_pixels) * 4 * CHANNELS
    cap_tbs = 6300.0 * (sm_mhz or 1965.0) * 1e6 / 1e12
    tbs = (tap_bytes + out_bytes) / (ms * 1e-3) / 1e12
    return {"tap_bytes": tap_bytes, "out_bytes": int(out_bytes), "achieved_tbs": round(tbs, 2), "cap_tbs": round(cap_tbs, 2),
            "frac": round(tbs / cap_tbs, 3), "note": "taps + stores through the L2 -> SM fabric against its ~6300 B/clk ceiling"}

def dist_env():
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs of its GPU's NUMA node BEFORE any pinned host memory is allocated (first touch
    then places the staging buffers next to the PCIe root the GPU hangs off).  Returns what was done, for the line."""
    try:
        bdf = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bdf.count(":") == 2 and len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]                                            # 00000000:1b:00.0 -> 0000:1b:00.0
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if node < 0 or len(nodes) < 2:
            return {"numa_node": node, "numa_nodes": len(nodes), "bound": False}
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "numa_nodes": len(nodes), "bound": bool(allowed), "cpus": len(allowed)}
    except Exception as e:
        return {"bound": False, "error": "%s: %s" % (type(e).__name__, e)}


class Ctx(object):
    """One process per GPU: rank / device / NCCL group shared by every workload of a run."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.rank, self.local_rank, self.world = dist_env()
        self.numa = bind_to_gpu_numa_node(self.local_rank)
        if args.gpus != self.world and self.world > 1:
            raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, self.world))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.dist = dist
        self.torch = torch

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def tensor_roofline(achieved_tf, timed_seconds, **extra):
    """Fraction of the MEASURED bf16 peak.  A timed region shorter than ~2 s runs at boost clocks, so the burst
    figure is the denominator; long regions settle at the power cap and are held against the sustained figure.
    Both fractions are always reported."""
    burst, src = measured_peaks("bf16_tflops")
    sustained, _ = measured_peaks("bf16_tflops_sustained")
    use_burst = timed_seconds < 2.0
    peak = burst if use_burst else sustained
    r = {"bound": "tensor", "achieved": round(achieved_tf, 1), "peak": peak,
         "peak_source": src + (" burst (timed region %.2f s)" % timed_seconds if use_burst else
                               " sustained (timed region %.1f s)" % timed_seconds),
         "unit": "TFLOP/s", "frac": round(achieved_tf / peak, 4), "frac_of_burst": round(achieved_tf / burst, 4),
         "frac_of_sustained": round(achieved_tf / sustained, 4), "traffic": None}
    r.update(extra)
    return r


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------

def run_ours_roi_features(args, ctx):
    import torch
    import torch.distributed as dist
    import image_captioning_b200 as pkg
    from image_captioning_b200 import synth

    rank, local_rank, world, dev = ctx.rank, ctx.local_rank, ctx.world, ctx.dev
    lib = pkg._lib.load()
    sms, cc = pkg._lib.device_info()

    # ---- synthetic workload (cfg2): every rank owns IMAGES_PER_GPU images ----
    B, N = IMAGES_PER_GPU, ROIS_PER_IMAGE
    rng = np.random.default_rng(SEED_CFG2 + 7919 * rank)
    boxes_np = synth.synth_boxes(rng, B, N, float(IMAGE_SHAPE[0]))
    gen = torch.Generator(device=dev).manual_seed(SEED_CFG2 + rank)
    fms = [torch.randn((B, IMAGE_SHAPE[0] >> l, IMAGE_SHAPE[1] >> l, CHANNELS), device=dev,
                       generator=gen) for l in range(2, 6)]
    boxes = torch.from_numpy(boxes_np).to(dev)
    out = torch.empty((B * N, POOL[0], POOL[1], CHANNELS), device=dev)
    levels = torch.empty((B, N), dtype=torch.int32, device=dev)

    def step():
        pkg.pyramid_roi_align(boxes, fms, POOL, IMAGE_SHAPE, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    _, levels = pkg.pyramid_roi_align(boxes, fms, POOL, IMAGE_SHAPE, out=out, return_levels=True)
    torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- device-resident timing: CUDA events on the launching (torch current) stream ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    clocks = sampler.summary()
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    ms_per_step = total_ms_max / args.steps
    rois_per_step = B * N * world
    value = rois_per_step / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (roi_align_kernel), algorithmic bytes ----
    t_unique = tap_unique_pixels(boxes_np, levels.cpu().numpy(), [tuple(f.shape[1:3]) for f in fms], POOL)
    alg_bytes = 4 * CHANNELS * (POOL[0] * POOL[1] * B * N + t_unique)
    k_ms = float(np.mean(kernel_ms))
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    peak, peak_src = measured_peaks()
    roofline = {"bound": "hbm", "kernel": "roi_order_kernel + roi_gather_records_kernel + roi_align_stream_kernel (fp32 output)", "achieved": round(achieved, 1),
                "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "frac_of_nominal_8000": round(achieved / 8000.0, 4),
                "traffic": None, "algorithmic_bytes_per_launch": alg_bytes,
                "t_unique_pixels": t_unique, "kernel_ms": round(k_ms, 5),
                "l2_path": l2_path(4 * POOL[0] * POOL[1] * B * N, 4 * CHANNELS * POOL[0] * POOL[1] * B * N, k_ms, (clocks or {}).get("sm_mhz"))}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            with open(prof) as f:
                roofline["traffic"] = json.load(f).get("roi_align_f32_dram_bytes_per_launch")
        except Exception:
            pass

    if args.no_e2e:
        return {"ms_per_step": round(ms_per_step, 5), "roofline_frac": roofline["frac"]}
    # ---- e2e: host buffers through the C-ABI host entry point ----
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    h_boxes = torch.from_numpy(boxes_np).pin_memory()
    h_fms = [f.cpu().pin_memory() for f in fms]
    h_out = torch.empty((B * N, POOL[0], POOL[1], CHANNELS)).pin_memory()
    ptrs = (ctypes.c_void_p * 4)(*[f.data_ptr() for f in h_fms])
    hs = (ctypes.c_int * 4)(*[f.shape[1] for f in h_fms])
    ws = (ctypes.c_int * 4)(*[f.shape[2] for f in h_fms])

    def e2e_step():
        pkg._lib.check(lib.dc_pyramid_roi_align_host_f32(
            ctypes.c_void_p(h_boxes.data_ptr()), ptrs, hs, ws, B, N, CHANNELS, POOL[0], POOL[1],
            IMAGE_SHAPE[0], IMAGE_SHAPE[1], ctypes.c_void_p(h_out.data_ptr()), None))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    ok = bool(torch.equal(h_out, out.cpu()))
    h2d = int(sum(f.numel() * 4 for f in h_fms) + h_boxes.numel() * 4)
    d2h = int(h_out.numel() * 4)
    e2e = {"value": round(rois_per_step / e2e_s, 1), "unit": "RoI/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "steps": e2e_steps, "matches_device_result": ok,
           "api": "dc_pyramid_roi_align_host_f32 (pinned host buffers)"}

    line = {
        "metric": "roi_features_per_sec", "value": round(value, 1), "unit": "RoI/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2 roi_features: %d images x %d RoIs per GPU, P2-P5 of 1024x1024, "
                               "256 ch fp32 -> 7x7x256" % (B, N),
                   "rois_per_step": rois_per_step, "sharding": "images per rank, no collective",
                   "l2": "inputs larger than L2 (pyramid %d MB, output %d MB per GPU)"
                         % (sum(f.numel() for f in fms) * 4 // 2 ** 20, out.numel() * 4 // 2 ** 20),
                   "sm_count": sms, "cc": cc},
        "clocks": clocks, "e2e": e2e, "gpu_launches": args.steps, "roofline": roofline,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(boxes_np, [f.cpu().numpy() for f in h_fms], budget_s=12.0)
    return line


# --------------------------------------------------------------------------------------------
# CPU baseline / reference arm (the ONLY places bench.py touches oracle/)
# --------------------------------------------------------------------------------------------

def cpu_baseline(boxes_np, fms_np, budget_s):
    """Times the reference's CPU algorithm (literal form: 4 per-level crop_and_resize calls +
    re-sort gather, boxes sharded over all host threads) restated in C (oracle/c), on as many
    whole images of the same workload as fit in ~budget_s seconds."""
    from tests import _c_oracle
    n_img = boxes_np.shape[0]
    N = boxes_np.shape[1]
    out = np.empty((N, POOL[0], POOL[1], CHANNELS), np.float32)
    scratch = np.empty_like(out)
    _c_oracle.pyramid_roi_align(boxes_np[:1], [f[:1] for f in fms_np], POOL, IMAGE_SHAPE, literal=True,
                                out=out, scratch=scratch)          # warm-up
    done, t0 = 0, time.perf_counter()
    while True:
        i = done % n_img
        _c_oracle.pyramid_roi_align(boxes_np[i:i + 1], [f[i:i + 1] for f in fms_np], POOL, IMAGE_SHAPE,
                                    literal=True, out=out, scratch=scratch)
        done += 1
        el = time.perf_counter() - t0
        if el > budget_s:
            break
    return {"value": round(done * N / el, 1), "unit": "RoI/s", "cores": _c_oracle.num_threads(),
            "kind": "port", "sample": "%d images x %d RoIs of the same cfg2 workload, literal "
            "reference form (4x crop_and_resize + re-sort), C/OpenMP restatement" % (done, N),
            "seconds": round(el, 2)}


def _host_threads():
    """Threads the CPU arm really has: OpenMP (C oracle) and BLAS (numpy) pools."""
    from tests import _c_oracle
    info = {"omp": _c_oracle.num_threads(), "cpu_count": os.cpu_count()}
    try:
        from threadpoolctl import threadpool_info
        info["blas"] = max([p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"] or [1])
    except Exception:
        info["blas"] = None
    return info


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    from image_captioning_b200 import synth
    rng = np.random.default_rng(SEED_CFG2)
    B, N = 2, ROIS_PER_IMAGE           # bounded sample: 2 images of the cfg2 workload per step
    boxes_np = synth.synth_boxes(rng, B, N, float(IMAGE_SHAPE[0]))
    fms = [rng.standard_normal((B, IMAGE_SHAPE[0] >> l, IMAGE_SHAPE[1] >> l, CHANNELS), dtype=np.float32)
           for l in range(2, 6)]
    from tests import _c_oracle
    out = np.empty((B * N, POOL[0], POOL[1], CHANNELS), np.float32)
    scratch = np.empty_like(out)

    def step():
        _c_oracle.pyramid_roi_align(boxes_np, fms, POOL, IMAGE_SHAPE, literal=True, out=out, scratch=scratch)

    for _ in range(max(1, min(args.warmup, 3))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    value = args.steps * B * N / el
    sample = ("%d images x %d RoIs per step (bounded sample of cfg2), literal reference form, "
              "C/OpenMP restatement of the TF CPU path" % (B, N))
    line = {"impl": "reference", "metric": "roi_features_per_sec", "value": round(value, 1),
            "unit": "RoI/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(el / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg2 roi_features (bounded sample: %d images x %d RoIs per step)" % (B, N)},
            "cpu_baseline": {"value": round(value, 1), "unit": "RoI/s", "cores": _c_oracle.num_threads(),
                             "kind": "port", "sample": sample},
            "e2e": {"value": round(value, 1), "unit": "RoI/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "threads": _host_threads(),
            "note": "the reference is TF-1.x/Keras Python and cannot be installed here; this arm "
                    "times the C restatement of its CPU algorithm (oracle/c) on all host threads (set explicitly: "
                    "torchrun's OMP_NUM_THREADS=1 is overridden)"}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# captions workload (default): ROIAlign -> head -> greedy decode
# --------------------------------------------------------------------------------------------

def _captions_workload(B=IMAGES_PER_GPU, N=ROIS_PER_IMAGE):
    return ("captions: per GPU %d images x %d RoIs (cfg2 RoI stage, P2-P5 of 1024x1024, 256 ch "
            "fp32) -> PyramidROIAlign -> RoI head -> v1 inject-LSTM greedy decode (hidden %d, "
            "vocab %d, embedding %d, P=%d)" % (B, N, UNITS, VOCAB, EMBED, PADDING))


def _decoder_weights():
    from image_captioning_b200 import synth
    return synth.synth_weights_v1(np.random.default_rng(SEED_DECODER), V=VOCAB, E=EMBED, U=UNITS, C=CHANNELS)


def run_ours_captions(args, ctx):
    import torch
    import torch.distributed as dist
    import image_captioning_b200 as pkg
    from image_captioning_b200 import synth

    rank, local_rank, world, dev = ctx.rank, ctx.local_rank, ctx.world, ctx.dev
    sms, cc = pkg._lib.device_info()

    B, N = IMAGES_PER_GPU, ROIS_PER_IMAGE
    R = B * N
    rng = np.random.default_rng(SEED_CFG2 + 7919 * rank)
    boxes_np = synth.synth_boxes(rng, B, N, float(IMAGE_SHAPE[0]))
    gen = torch.Generator(device=dev).manual_seed(SEED_CFG2 + rank)
    fms = [torch.randn((B, IMAGE_SHAPE[0] >> l, IMAGE_SHAPE[1] >> l, CHANNELS), device=dev, generator=gen)
           for l in range(2, 6)]
    boxes = torch.from_numpy(boxes_np).to(dev)
    w = _decoder_weights()
    cfg = pkg.DenseCapConfig(VOCAB, w["imgcap_embedding_layer/embeddings"], 1, PADDING)
    model = pkg.build_lstm_model([POOL[0], POOL[1], CHANNELS], cfg, UNITS, "inference", dtype="bfloat16", device=dev)
    model.set_weights(w)
    feats = torch.empty((R, POOL[0], POOL[1], CHANNELS), dtype=torch.bfloat16, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    _, levels = pkg.pyramid_roi_align(boxes, fms, POOL, IMAGE_SHAPE, out_dtype=torch.bfloat16, out=feats,
                                      return_levels=True)
    tokens = model.generate(feats)
    tokens_fused = model.caption_rois(boxes, fms, IMAGE_SHAPE)          # the single-call public API
    torch.cuda.synchronize()
    same_as_fused = bool(torch.equal(tokens, tokens_fused))
    # the persistent greedy-loop kernel against the launch-per-GEMM form of the same decoder (a fresh feature buffer: the
    # handle's CUDA graphs are keyed by the input pointer, so this call runs eagerly with the other path)
    loop_vs_launches = None
    if os.environ.get("DCAP_GREEDY_LOOP", "1") != "0":
        prev = os.environ.get("DCAP_GREEDY_LOOP")
        os.environ["DCAP_GREEDY_LOOP"] = "0"
        try:
            tokens_launches = model.generate(feats.clone())
            torch.cuda.synchronize()
            loop_vs_launches = float((tokens_launches == tokens).float().mean().item())
        finally:
            if prev is None:
                os.environ.pop("DCAP_GREEDY_LOOP", None)
            else:
                os.environ["DCAP_GREEDY_LOOP"] = prev
    for _ in range(max(args.warmup, 3)):
        pkg.pyramid_roi_align(boxes, fms, POOL, IMAGE_SHAPE, out_dtype=torch.bfloat16, out=feats)
        model.generate(feats)
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    K = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    barrier()
    for i in range(K):
        ev[i][0].record()
        pkg.pyramid_roi_align(boxes, fms, POOL, IMAGE_SHAPE, out_dtype=torch.bfloat16, out=feats)
        ev[i][1].record()
        tokens = model.generate(feats)
        ev[i][2].record()
    barrier()
    total_ms = ev[0][0].elapsed_time(ev[-1][2])
    roi_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    dec_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    clocks = sampler.summary()
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / K
    value = R * world / (ms_per_step * 1e-3)

    # rooflines: decoder GEMMs (tensor, dominant share of the step) and ROIAlign (HBM)
    flops = FLOP_PER_ROI_GREEDY * R
    achieved_tf = flops / (dec_ms * 1e-3) / 1e12
    loop_kernel = os.environ.get("DCAP_GREEDY_LOOP", "1") != "0"          # csrc/greedy_loop.cu: the 15 steps as ONE persistent kernel
    roofline = tensor_roofline(achieved_tf, total_ms * 1e-3,
                               kernel=("greedy_loop_kernel (all %d decode steps in one persistent launch) + gemm_bf16_tc2_kernel (head, hoisted terms)" % PADDING)
                               if loop_kernel else
                               "gemm_bf16_tc2_kernel (head + 15 decode steps: %d launches per step)" % (4 + 4 * PADDING),
                               algorithmic_flops_per_step=flops, decoder_ms=round(dec_ms, 4),
                               share_of_step=round(dec_ms / (roi_ms + dec_ms), 3))
    t_unique = tap_unique_pixels(boxes_np, levels.cpu().numpy(), [tuple(f.shape[1:3]) for f in fms], POOL)
    alg_bytes = 4 * CHANNELS * t_unique + 2 * CHANNELS * POOL[0] * POOL[1] * R        # fp32 taps in, bf16 rows out
    hbm_peak, hbm_src = measured_peaks("hbm_gbs")
    achieved_gb = alg_bytes / (roi_ms * 1e-3) / 1e9
    roofline_hbm = {"bound": "hbm", "kernel": "roi_order_kernel + roi_gather_records_kernel + roi_align_stream_kernel (bf16 output)",
                    "achieved": round(achieved_gb, 1), "peak": hbm_peak, "peak_source": hbm_src, "unit": "GB/s",
                    "frac": round(achieved_gb / hbm_peak, 4), "frac_of_nominal_8000": round(achieved_gb / 8000.0, 4),
                    "traffic": None, "algorithmic_bytes_per_step": alg_bytes, "t_unique_pixels": t_unique,
                    "roi_align_ms": round(roi_ms, 4), "share_of_step": round(roi_ms / (roi_ms + dec_ms), 3),
                    "l2_path": l2_path(4 * POOL[0] * POOL[1] * R, 2 * CHANNELS * POOL[0] * POOL[1] * R, roi_ms, (clocks or {}).get("sm_mhz"))}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            with open(prof) as f:
                tr = json.load(f)
            roofline["traffic"] = tr.get("gemm_dram_bytes_per_step")
            roofline_hbm["traffic"] = tr.get("roi_align_bf16_dram_bytes_per_launch")
        except Exception:
            pass

    line = {
        "metric": "roi_captions_per_sec", "value": round(value, 1), "unit": "RoI captions/s",
        "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": _captions_workload(B, N),
                   "rois_per_step": R * world, "sharding": "images per rank, no collective",
                   "precision": "ROIAlign fp32 arithmetic with bf16 output; decoder bf16 operands, fp32 accumulate/state",
                   "l2": "inputs larger than L2 (pyramid %d MB per GPU; decoder weights + activations %d MB)"
                         % (sum(f.numel() for f in fms) * 4 // 2 ** 20, (63 + R * (12544 * 2 + 4 * 2048 * 2) // 2 ** 20)),
                   "sm_count": sms, "cc": cc, "public_api_matches": same_as_fused,
                   "token_agreement_with_launch_per_gemm_path": loop_vs_launches},
        "clocks": clocks,
        # per step: ROIAlign (order + records + stream), head (2 GEMMs), ONE merged hoisted-term GEMM, token fill, first
        # embedding gather, ONE greedy_loop_kernel launch = 9 -- or, with DCAP_GREEDY_LOOP=0: head 2, bf16 cast, 2 hoisted-term
        # GEMMs, fill, gather, then P x (2 gate GEMMs + dense1 + vocabulary GEMM + merge); profiles/r2_launches_captions*.txt
        "gpu_launches": K * ((3 + 2 + 1 + 1 + 1 + 1) if loop_kernel else (3 + 7 + 5 * PADDING)), "roofline": roofline,
        "roofline_hbm": roofline_hbm,
    }
    if not args.no_e2e:
        e2e_steps = max(3, min(K, args.e2e_steps))
        h_boxes = torch.from_numpy(boxes_np).pin_memory()
        h_fms = [f.cpu().pin_memory() for f in fms]
        h_np = [f.numpy() for f in h_fms]
        model.caption_rois(h_boxes.numpy(), h_np, IMAGE_SHAPE)          # warm-up (allocations)
        model.caption_rois(h_boxes.numpy(), h_np, IMAGE_SHAPE)
        # the host-side ceiling: every rank uploads its pinned pyramid at the same time, nothing else running
        h2d_bytes = int(sum(f.numel() * 4 for f in h_fms) + h_boxes.numel() * 4)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        c0.record()
        for _ in range(3):
            for f, hf in zip(fms, h_fms):
                f.copy_(hf, non_blocking=True)
        c1.record()
        barrier()
        copy_s = ctx.max_over_ranks(c0.elapsed_time(c1) * 1e-3 / 3)
        barrier()
        # pipelined public API: two calls in flight, the upload of call k+1 under the decode tail of call k
        outs = []
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            outs.append(model.caption_rois(h_boxes.numpy(), h_np, IMAGE_SHAPE, wait=False))
            if i >= 1:
                model.caption_rois_wait()
        model.caption_rois_wait()
        barrier()
        e2e_s = ctx.max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        h_tok = outs[-1]
        agree = float((torch.from_numpy(h_tok).to(dev) == tokens).float().mean().item())
        same_all = all(np.array_equal(o, outs[0]) for o in outs)
        line["e2e"] = {"value": round(R * world / e2e_s, 1), "unit": "RoI captions/s",
                       "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(R * PADDING * 4), "steps": e2e_steps,
                       "token_agreement_with_device_run": round(agree, 5), "all_steps_identical": same_all,
                       "h2d_gbs_per_rank_all_ranks_copying": round(h2d_bytes / copy_s / 1e9, 2),
                       "host_copy_ceiling": round(R * world / copy_s, 1),
                       "frac_of_host_copy_ceiling": round(copy_s / e2e_s, 4), "numa": ctx.numa,
                       "api": "dc_caption_rois_host_submit / _wait (pinned host pyramid + boxes in, token ids out; images "
                              "uploaded one by one under ROIAlign + decode of the previous one, two calls in flight); "
                              "host_copy_ceiling = the same metric if the pinned H2D copy of the pyramids were the only cost, "
                              "measured with all ranks copying at once"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_captions(boxes_np[:1], [f[:1].cpu().numpy() for f in fms], w, 12.0)
    return line


def _cpu_caption_pass(boxes_np, fms_np, w, out, scratch):
    """The reference's CPU algorithm restated (oracle/): literal PyramidROIAlign (C/OpenMP) ->
    head -> v1 greedy (numpy fp32 on the BLAS threads; incremental batched scan, i.e. the
    best-effort vectorised form, NOT the reference's batch-1 O(P^2) loop)."""
    from tests import _c_oracle
    from oracle import decoder as dec
    _c_oracle.pyramid_roi_align(boxes_np, fms_np, POOL, IMAGE_SHAPE, literal=True, out=out, scratch=scratch)
    tok, _ = dec.greedy_v1(dec.head(out, w), w, PADDING, return_logits=True)
    return tok


def cpu_baseline_captions(boxes_np, fms_np, w, budget_s, n_rois=250):
    from tests import _c_oracle
    b = np.ascontiguousarray(boxes_np[:, :n_rois])
    out = np.empty((n_rois, POOL[0], POOL[1], CHANNELS), np.float32)
    scratch = np.empty_like(out)
    _cpu_caption_pass(b, fms_np, w, out, scratch)
    done, t0 = 0, time.perf_counter()
    while True:
        _cpu_caption_pass(b, fms_np, w, out, scratch)
        done += 1
        el = time.perf_counter() - t0
        if el > budget_s:
            break
    return {"value": round(done * n_rois / el, 1), "unit": "RoI captions/s", "cores": _c_oracle.num_threads(),
            "kind": "port", "seconds": round(el, 2),
            "sample": "%d passes over 1 image x %d RoIs of the same workload: literal PyramidROIAlign (C/OpenMP) + "
                      "head + v1 greedy P=%d as a batched incremental numpy/BLAS scan (the reference itself runs "
                      "batch 1 and O(P^2) LSTM steps, which is slower)" % (done, n_rois, PADDING)}


def run_reference_captions(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    from image_captioning_b200 import synth
    from tests import _c_oracle
    rng = np.random.default_rng(SEED_CFG2)
    n_rois = 250
    boxes_np = synth.synth_boxes(rng, 1, n_rois, float(IMAGE_SHAPE[0]))
    fms = [rng.standard_normal((1, IMAGE_SHAPE[0] >> l, IMAGE_SHAPE[1] >> l, CHANNELS), dtype=np.float32)
           for l in range(2, 6)]
    w = _decoder_weights()
    out = np.empty((n_rois, POOL[0], POOL[1], CHANNELS), np.float32)
    scratch = np.empty_like(out)
    for _ in range(max(1, min(args.warmup, 2))):
        _cpu_caption_pass(boxes_np, fms, w, out, scratch)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _cpu_caption_pass(boxes_np, fms, w, out, scratch)
    el = time.perf_counter() - t0
    value = args.steps * n_rois / el
    sample = ("1 image x %d RoIs per step (bounded sample of the captions workload): literal PyramidROIAlign "
              "(C/OpenMP) + head + v1 greedy P=%d, batched incremental numpy/BLAS scan" % (n_rois, PADDING))
    line = {"impl": "reference", "metric": "roi_captions_per_sec", "value": round(value, 1),
            "unit": "RoI captions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(el / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": _captions_workload(), "sample_per_step": "1 image x %d RoIs" % n_rois},
            "cpu_baseline": {"value": round(value, 1), "unit": "RoI captions/s", "cores": _c_oracle.num_threads(),
                             "kind": "port", "sample": sample},
            "e2e": {"value": round(value, 1), "unit": "RoI captions/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "threads": _host_threads(),
            "note": "the reference is TF-1.x/Keras Python and cannot be installed here (no TensorFlow); this arm "
                    "times the CPU restatement of its algorithm (oracle/) on all host threads (set explicitly: "
                    "torchrun's OMP_NUM_THREADS=1 is overridden)"}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# train workload (BASELINE.json configs[2]): v1 decoder training step, bf16, global batch 4096,
# P = 16, data-parallel with one NCCL gradient all-reduce per step
# --------------------------------------------------------------------------------------------
TRAIN_BATCH, TRAIN_P = int(os.environ.get("DCAP_TRAIN_BATCH", "4096")), 16      # the override is for tuning runs only
FLOP_PER_ROI_FWD_TRAIN = 27787264 + 4194304 + 2097152 + TRAIN_P * (1228800 + 2097152 + 4194304 + 1048576 + 20480000)


def run_ours_train(args, ctx):
    import torch
    import torch.distributed as dist
    import image_captioning_b200 as pkg
    from image_captioning_b200 import synth, parallel

    rank, local_rank, world, dev = ctx.rank, ctx.local_rank, ctx.world, ctx.dev
    sms, cc = pkg._lib.device_info()
    lo, hi = parallel.shard_bounds(TRAIN_BATCH, rank, world)         # global batch split evenly: strong scaling
    Bl = hi - lo
    # random initialisation as at the start of training (Glorot / orthogonal, no trained-like logit sharpening)
    w = synth.synth_weights_v1(np.random.default_rng(SEED_DECODER), V=VOCAB, E=EMBED, U=UNITS, C=CHANNELS,
                               trained_like=False)
    cfg = pkg.DenseCapConfig(VOCAB, w["imgcap_embedding_layer/embeddings"], Bl, TRAIN_P)
    model = pkg.build_lstm_model([POOL[0], POOL[1], CHANNELS], cfg, UNITS, "training", dtype="bfloat16", device=dev)
    model.set_weights(w)
    model.compile(optimizer=pkg.Adam(amsgrad=True), loss=pkg.roi_caption_loss)
    # optimiser: replicated behind an all-reduce, or sharded (reduce-scatter, update of the rank's own ranges, all-gather of
    # the updated parameters) -- "auto" shards from 4 ranks on (measured crossover); DCAP_SHARD_OPT=0/1 forces one
    so = os.environ.get("DCAP_SHARD_OPT")
    trainer = parallel.DataParallelTrainer(model, overlap=os.environ.get("DCAP_NO_OVERLAP") is None,
                                           shard_optimizer="auto" if so is None else so != "0")
    rng = np.random.default_rng(1003)
    gt_np = synth.synth_captions(rng, TRAIN_BATCH, TRAIN_P, VOCAB)[lo:hi]
    gen = torch.Generator(device=dev).manual_seed(1003 + rank)
    feats = torch.randn((Bl, POOL[0], POOL[1], CHANNELS), device=dev, generator=gen)       # fp32 RoI features (ROIAlign output)
    gt = torch.from_numpy(gt_np).to(dev).to(torch.int32)
    npos = float(TRAIN_BATCH * TRAIN_P)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    losses = []
    for _ in range(max(args.warmup, 3)):
        losses.append(trainer.train_step(feats, gt, None, npos))
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    K = args.steps
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    barrier()
    ev[0].record()
    for i in range(K):
        losses.append(trainer.train_step(feats, gt, None, npos))
        ev[i + 1].record()
    barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    clocks = sampler.summary()
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / K
    value = TRAIN_BATCH / (ms_per_step * 1e-3)
    flops = 3.0 * FLOP_PER_ROI_FWD_TRAIN * Bl
    achieved = flops / (ms_per_step * 1e-3) / 1e12
    loss_hist = [float(l.item()) for l in losses]
    line = {
        "metric": "train_rois_per_sec", "value": round(value, 1), "unit": "RoI/s", "n_gpus": world, "steps": K,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "cfg3 train: v1 inject-LSTM decoder training step (head trainable), global batch %d RoIs "
                               "x P=%d, hidden %d, vocab %d, embedding %d; bf16 operands / fp32 master weights; "
                               "forward + roi_caption_loss + BPTT + NCCL gradient all-reduce + Keras AMSGrad update"
                               % (TRAIN_BATCH, TRAIN_P, UNITS, VOCAB, EMBED),
                   "rois_per_step": TRAIN_BATCH, "rois_per_rank": Bl, "sharding": "global batch split over ranks; "
                   "one all-reduce(sum) of %d fp32 gradients per step" % model.grad_buffer().numel(),
                   "optimizer": "sharded over the ranks (reduce-scatter, range update, all-gather)" if trainer.shard_optimizer
                                else "replicated on every rank",
                   "l2": "activations larger than L2 (logits %d MB per rank)" % (Bl * TRAIN_P * VOCAB * 4 // 2 ** 20),
                   "sm_count": sms, "cc": cc, "loss_first": round(loss_hist[0], 4), "loss_last": round(loss_hist[-1], 4)},
        "clocks": clocks, "gpu_launches": K * 132,      # profiles/r2_launches_train_final.txt
        "roofline": tensor_roofline(achieved, total_ms * 1e-3,
                                    kernel="gemm_bf16_tc2_kernel (forward scan + time-batched dense/vocab GEMMs + dgrad/wgrad GEMMs)",
                                    algorithmic_flops_per_step_per_rank=flops,
                                    note="whole step time (GEMMs + softmax/xent + cell backward + optimiser + all-reduce) "
                                         "against 3x forward FLOPs"),
    }
    # where the step goes (names the strong-scaling limiter): local forward+backward alone, optimiser alone; the rest of
    # ms_per_step is the exposed part of the NCCL all-reduce plus launch gaps
    def timed(fn, n=4):
        fn()
        a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.barrier()
        a.record()
        for _ in range(n):
            fn()
        b2.record()
        torch.cuda.synchronize()
        return ctx.max_over_ranks(a.elapsed_time(b2) / n)
    fb_ms = timed(lambda: model.train_step_device(feats, gt, None, 1.0 / npos))
    # host side of the same call: seconds the CPU needs to ENQUEUE one forward+backward pass (launches, tensor maps,
    # events) with an empty queue ahead of it -- when this approaches forward_backward_ms the step is launch-bound
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        model.train_step_device(feats, gt, None, 1.0 / npos)
    enq_ms = (time.perf_counter() - t0) / 4 * 1e3
    torch.cuda.synchronize()
    saved_iter = model.optimizer.iterations
    opt_ms = timed(lambda: model.apply_gradients())
    model.optimizer.iterations = saved_iter
    line["breakdown"] = {"forward_backward_ms": round(fb_ms, 4), "forward_backward_host_enqueue_ms": round(enq_ms, 4),
                         "optimizer_ms": round(opt_ms, 4),
                         "allreduce_exposed_plus_gaps_ms": round(max(0.0, ms_per_step - fb_ms - opt_ms), 4),
                         "allreduce_bytes": int(model.grad_buffer().numel() * 4) if world > 1 else 0,
                         "kernel_launches_per_step": 132 + (5 if world > 1 else 0),
                         "note": "per-rank batch %d: the %d recurrent step kernels run %d-row GEMMs" % (Bl, 6 * TRAIN_P, Bl)}
    if not args.no_e2e:
        e2e_steps = max(1, min(K, args.e2e_steps))
        h_feats = feats.cpu().pin_memory()
        h_gt = gt.cpu().pin_memory()
        # every step's batch crosses PCIe inside the timed region, one step ahead of the step that trains on it
        # (parallel.HostBatchPrefetcher: the upload of batch k+1 runs under the kernels of step k); the loss of every
        # step is read back (D2H + sync)
        pf = parallel.HostBatchPrefetcher(dev)

        def e2e_run(n):
            out = []
            pf.put(h_feats, h_gt)
            for k in range(n):
                if k + 1 < n:
                    pf.put(h_feats, h_gt)
                d_f, d_g = pf.get()
                loss = trainer.train_step(d_f, d_g, None, npos)
                pf.done()
                out.append(float(loss.item()))
            return out
        e2e_run(2)
        barrier()
        t0 = time.perf_counter()
        e2e_run(e2e_steps)
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        line["e2e"] = {"value": round(TRAIN_BATCH / float(t.item()), 1), "unit": "RoI/s",
                       "h2d_bytes_per_step": int(h_feats.numel() * 4 + h_gt.numel() * 4), "d2h_bytes_per_step": 4,
                       "steps": e2e_steps, "api": "parallel.HostBatchPrefetcher + DataParallelTrainer.train_step (pinned host "
                       "features + captions in one step ahead, scalar loss out every step)"}
    return line


# --------------------------------------------------------------------------------------------
# beam workload (BASELINE.json configs[3]): width-3 beam decoding of 100k pre-extracted 1024-d RoI
# feature vectors, rows sharded over ranks (no collective)
# --------------------------------------------------------------------------------------------
BEAM_ROIS, BEAM_K, BEAM_CHUNK = 100000, 3, int(os.environ.get("DCAP_BEAM_CHUNK", "33334"))
# hoisted terms once + (1 + (P-2)*k) word-model rows of 29.05 MFLOP (first step has one live beam)
FLOP_PER_ROI_BEAM = 4194304 + 2097152 + (1 + (PADDING - 2) * BEAM_K) * (1228800 + 2097152 + 4194304 + 1048576 + 20480000)


def run_ours_beam(args, ctx):
    import torch
    import torch.distributed as dist
    import image_captioning_b200 as pkg
    from image_captioning_b200 import parallel

    rank, local_rank, world, dev = ctx.rank, ctx.local_rank, ctx.world, ctx.dev
    sms, cc = pkg._lib.device_info()
    lo, hi = parallel.shard_bounds(BEAM_ROIS, rank, world)
    n = hi - lo
    w = _decoder_weights()
    cfg = pkg.DenseCapConfig(VOCAB, w["imgcap_embedding_layer/embeddings"], 1, PADDING)
    model = pkg.build_lstm_model([POOL[0], POOL[1], CHANNELS], cfg, UNITS, "inference", dtype="bfloat16", device=dev)
    model.set_weights(w)
    gen = torch.Generator(device=dev).manual_seed(1004 + rank)
    feats = torch.relu(torch.randn((n, 1024), device=dev, generator=gen))       # post-ReLU head features

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(min(args.warmup, 3), 1)):
        tokens, scores = model.beam_search(feats, beam_width=BEAM_K, chunk=BEAM_CHUNK)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    K = args.steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        tokens, scores = model.beam_search(feats, beam_width=BEAM_K, chunk=BEAM_CHUNK)
    e1.record()
    barrier()
    total_ms = e0.elapsed_time(e1)
    clocks = sampler.summary()
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / K
    value = BEAM_ROIS / (ms_per_step * 1e-3)
    flops = FLOP_PER_ROI_BEAM * n
    achieved = flops / (ms_per_step * 1e-3) / 1e12
    chunks = (n + BEAM_CHUNK - 1) // BEAM_CHUNK
    line = {
        "metric": "beam_captions_per_sec", "value": round(value, 1), "unit": "RoI captions/s", "n_gpus": world, "steps": K,
        "warmup": max(min(args.warmup, 3), 1), "ms_per_step": round(ms_per_step, 3), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "cfg4 beam: width-%d beam decoding (probability-sum scores, P=%d) of %d pre-extracted 1024-d "
                               "RoI feature vectors, hidden %d, vocab %d; rows sharded over ranks, %d-RoI chunks"
                               % (BEAM_K, PADDING, BEAM_ROIS, UNITS, VOCAB, BEAM_CHUNK),
                   "rois_per_step": BEAM_ROIS, "rois_per_rank": n, "sharding": "contiguous RoI rows per rank, no collective",
                   "l2": "inputs larger than L2 (features %d MB, beam state %d MB per chunk)"
                         % (n * 4096 // 2 ** 20, BEAM_CHUNK * BEAM_K * (832 + 1024) * 4 // 2 ** 20),
                   "sm_count": sms, "cc": cc},
        "clocks": clocks, "gpu_launches": K * chunks * (8 + 8 * (PADDING - 1)),
        "roofline": tensor_roofline(achieved, total_ms * 1e-3,
                                    kernel="gemm_bf16_tc2_kernel (gate GEMMs + fused cell, dense1, vocabulary GEMM + fused top-k epilogue)",
                                    algorithmic_flops_per_step_per_rank=flops),
    }
    if not args.no_e2e:
        h_feats = feats.cpu().pin_memory()
        barrier()
        t0 = time.perf_counter()
        d = h_feats.to(dev, non_blocking=True)
        tk, sc = model.beam_search(d, beam_width=BEAM_K, chunk=BEAM_CHUNK)
        h_tok, h_sc = tk.cpu(), sc.cpu()
        barrier()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        line["e2e"] = {"value": round(BEAM_ROIS / float(t.item()), 1), "unit": "RoI captions/s",
                       "h2d_bytes_per_step": int(h_feats.numel() * 4),
                       "d2h_bytes_per_step": int(h_tok.numel() * 4 + h_sc.numel() * 8), "steps": 1,
                       "api": "RoiCaptionModel.beam_search (host features in, [N,k,P] ids + [N,k] scores out)"}
    return line


# --------------------------------------------------------------------------------------------
# proposals workload: the box front-end that feeds PyramidROIAlign when RPN
# proposals are used -- ProposalLayer on the reference's configuration, images sharded over ranks
# --------------------------------------------------------------------------------------------
PROP_IMAGES, PROP_COUNT, PROP_LIMIT, PROP_NMS = 32, 1000, 6000, 0.7


def _rpn_like(gen, n_images, anchors, dev, n_objects=40):
    """RPN-like synthetic outputs: per image a few dozen "objects"; an anchor's foreground score falls off with
    its distance and scale mismatch to the nearest object, and its deltas regress towards that object (plus
    noise).  The top-6000 anchors therefore cluster around the objects and NMS has real work (with i.i.d. boxes
    almost nothing is suppressed and the scan would stop after ~1000 boxes)."""
    import torch
    a = torch.from_numpy(anchors.astype(np.float32)).to(dev)
    ah, aw = a[:, 2] - a[:, 0], a[:, 3] - a[:, 1]
    acy, acx = a[:, 0] + 0.5 * ah, a[:, 1] + 0.5 * aw
    asz = torch.sqrt(ah * aw)
    A = a.shape[0]
    probs = torch.empty((n_images, A, 2), device=dev)
    bbox = torch.empty((n_images, A, 4), device=dev)
    std = torch.tensor([0.1, 0.1, 0.2, 0.2], device=dev)
    for b in range(n_images):
        oc = torch.rand((n_objects, 2), device=dev, generator=gen) * 1024.0
        osz = torch.exp(torch.rand((n_objects,), device=dev, generator=gen) * (np.log(512.0) - np.log(32.0)) + np.log(32.0))
        oasp = torch.exp((torch.rand((n_objects,), device=dev, generator=gen) - 0.5) * 1.2)
        d2 = ((acy[:, None] - oc[None, :, 0]) ** 2 + (acx[:, None] - oc[None, :, 1]) ** 2) / (osz[None] ** 2)
        aff = torch.exp(-3.0 * d2 - torch.log2(asz[:, None] / osz[None]) ** 2)
        best, which = aff.max(1)
        logit = 9.0 * best - 6.0 + 0.7 * torch.randn((A,), device=dev, generator=gen)
        fg = torch.sigmoid(logit)
        probs[b, :, 0], probs[b, :, 1] = 1 - fg, fg
        oh, ow = osz[which] * torch.sqrt(oasp[which]), osz[which] / torch.sqrt(oasp[which])
        tgt = torch.stack([(oc[which, 0] - acy) / ah, (oc[which, 1] - acx) / aw, torch.log(oh / ah), torch.log(ow / aw)], 1)
        near = (best > 0.05).float()[:, None]                  # far anchors regress nothing in particular
        noise = torch.randn((A, 4), device=dev, generator=gen)
        bbox[b] = (tgt * near + noise * (0.2 + 0.4 * (1 - near))) / std
    return probs.contiguous(), bbox.contiguous()


def run_ours_proposals(args, ctx):
    import torch
    import torch.distributed as dist
    import image_captioning_b200 as pkg

    rank, local_rank, world, dev = ctx.rank, ctx.local_rank, ctx.world, ctx.dev
    sms, cc = pkg._lib.device_info()
    cfg = pkg.ProposalConfig()
    anchors = cfg.anchors()
    A = anchors.shape[0]
    layer = pkg.ProposalLayer(PROP_COUNT, PROP_NMS, anchors, cfg)
    gen = torch.Generator(device=dev).manual_seed(1005 + rank)
    probs, bbox = _rpn_like(gen, PROP_IMAGES, anchors, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W = max(args.warmup, 3)
    for _ in range(W):
        rois = layer([probs, bbox])
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    K = args.steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        rois = layer([probs, bbox])
    e1.record()
    barrier()
    total_ms = e0.elapsed_time(e1)
    clocks = sampler.summary()
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / K
    value = PROP_IMAGES * world / (ms_per_step * 1e-3)
    # algorithmic bytes per image: the [A,2] score rows once, the gathered delta + anchor rows of the candidates,
    # the IoU bit masks written and (at most) read once, the padded output
    n_blk = (PROP_LIMIT + 63) // 64
    bytes_img = A * 8 + PROP_LIMIT * 32 + 2 * (PROP_LIMIT * n_blk * 8 // 2) + PROP_COUNT * 16
    hbm_peak, hbm_src = measured_peaks("hbm_gbs")
    achieved = bytes_img * PROP_IMAGES / (ms_per_step * 1e-3) / 1e9
    line = {
        "metric": "proposal_images_per_sec", "value": round(value, 1), "unit": "images/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "proposals: ProposalLayer on %d images per GPU, %d anchors of a 1024x1024 image, top %d -> "
                               "refine/clip -> NMS %.1f -> %d proposals; synthetic RPN outputs clustered around 40 objects per image"
                               % (PROP_IMAGES, A, PROP_LIMIT, PROP_NMS, PROP_COUNT),
                   "images_per_step": PROP_IMAGES * world, "sharding": "images per rank, no collective",
                   "l2": "inputs larger than L2 (%d MB of RPN outputs per step)" % (PROP_IMAGES * A * 24 // 2 ** 20),
                   "sm_count": sms, "cc": cc},
        # 1 select + (IoU mask + scan) per band of 1024 candidates, at most 4 bands (csrc/proposals.cu)
        "clocks": clocks, "gpu_launches": K * (1 + 2 * min(4, (n_blk + 15) // 16)),
        "roofline": {"bound": "hbm", "kernel": "proposal_select_kernel + proposal_iou_mask_kernel + proposal_nms_scan_kernel "
                     "(not HBM-bound: the IoU mask kernel is issue-bound, select and scan are chains of short dependent phases; "
                     "profiles/r1_proposals_full.txt)", "achieved": round(achieved, 1), "peak": hbm_peak,
                     "peak_source": hbm_src, "unit": "GB/s", "frac": round(achieved / hbm_peak, 4), "traffic": None,
                     "algorithmic_bytes_per_image": bytes_img},
    }
    n_valid = layer([probs, bbox], return_details=True)[1]
    line["config"]["proposals_found_per_image"] = [int(n_valid.min()), float(n_valid.float().mean()), int(n_valid.max())]
    if not args.no_e2e:
        h_p, h_b = probs.cpu().pin_memory(), bbox.cpu().pin_memory()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            r = layer([h_p.to(dev, non_blocking=True), h_b.to(dev, non_blocking=True)])
            h_r = r.cpu()
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.e2e_steps
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        line["e2e"] = {"value": round(PROP_IMAGES * world / float(t.item()), 1), "unit": "images/s",
                       "h2d_bytes_per_step": int(h_p.numel() * 4 + h_b.numel() * 4), "d2h_bytes_per_step": int(h_r.numel() * 4),
                       "steps": args.e2e_steps, "api": "ProposalLayer.__call__ (pinned host RPN outputs in, proposals out)"}
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import proposals as opr
        p_np, b_np = probs[:2].cpu().numpy(), bbox[:2].cpu().numpy()
        t0, n = time.perf_counter(), 0
        while time.perf_counter() - t0 < 10.0:
            want = opr.proposal_layer(p_np, b_np, anchors, PROP_COUNT, PROP_NMS, cfg.IMAGE_SHAPE)
            n += 2
        dt = time.perf_counter() - t0
        got = rois[:2].cpu().numpy()
        line["cpu_baseline"] = {"value": round(n / dt, 2), "unit": "images/s", "cores": 1, "kind": "port",
                                "sample": "%d images of the same workload through oracle/proposals.py (numpy)" % n,
                                "bit_exact_vs_gpu": bool(np.array_equal(got.view(np.uint32), want.view(np.uint32)))}
    return line


def run_reference_proposals(args):
    """Reference arm of the proposals workload: the numpy restatement of ProposalLayer (oracle/proposals.py) on one
    image of the same configuration per step (the reference's own layer is a TensorFlow graph)."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    import image_captioning_b200.proposals as prod
    from oracle import proposals as opr
    cfg = prod.ProposalConfig()
    anchors = cfg.anchors()
    A = anchors.shape[0]
    rng = np.random.default_rng(1005)
    fg = (1.0 / (1.0 + np.exp(-(rng.standard_normal((1, A)) * 2.5 - 4.0)))).astype(np.float32)
    probs = np.stack([1 - fg, fg], -1)
    bbox = (rng.standard_normal((1, A, 4)) * 1.5).astype(np.float32)
    for _ in range(max(1, min(args.warmup, 2))):
        opr.proposal_layer(probs, bbox, anchors, PROP_COUNT, PROP_NMS, cfg.IMAGE_SHAPE)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        opr.proposal_layer(probs, bbox, anchors, PROP_COUNT, PROP_NMS, cfg.IMAGE_SHAPE)
    el = time.perf_counter() - t0
    value = args.steps / el
    line = {"impl": "reference", "metric": "proposal_images_per_sec", "value": round(value, 2), "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(el / args.steps * 1e3, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "proposals (bounded sample: 1 image x %d anchors per step)" % A},
            "cpu_baseline": {"value": round(value, 2), "unit": "images/s", "cores": 1, "kind": "port",
                             "sample": "1 image per step through oracle/proposals.py (numpy)"},
            "e2e": {"value": round(value, 2), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# dense-captioning workload shaped like BASELINE.json configs[4]: 5000 images x 300 RoIs, images dealt
# round-robin over the ranks, FPN pyramids already on the device (the backbone's output), boxes in / ids out
# --------------------------------------------------------------------------------------------
VG_IMAGES, VG_ROIS = 5000, 300
VG_BATCH = int(os.environ.get("DCAP_VG_BATCH", "24"))         # images per caption_rois call: 7200 RoIs (measured 8 / 16 / 24 / 32 images: 1.73 / 1.98 / 2.10 / 2.11 M captions/s)
VG_POOL = 2 * VG_BATCH


def run_ours_captions_vg(args, ctx):
    import torch
    import image_captioning_b200 as pkg
    from image_captioning_b200 import synth

    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    my_images = list(range(rank, VG_IMAGES, world))                      # round-robin image sharding
    n_batches = (len(my_images) + VG_BATCH - 1) // VG_BATCH
    rng = np.random.default_rng(1004 + 7919 * rank)
    boxes_np = synth.synth_boxes(rng, n_batches * VG_BATCH, VG_ROIS, float(IMAGE_SHAPE[0]))
    gen = torch.Generator(device=dev).manual_seed(1004 + rank)
    # a pool of VG_POOL distinct pyramids stands in for the backbone's per-image output (generating 5000 pyramids
    # = 445 GB would time torch.randn, not the path); every batch reads 8 of them with its own boxes
    pool = [torch.randn((VG_POOL, IMAGE_SHAPE[0] >> l, IMAGE_SHAPE[1] >> l, CHANNELS), device=dev, generator=gen)
            for l in range(2, 6)]
    w = _decoder_weights()
    cfg = pkg.DenseCapConfig(VOCAB, w["imgcap_embedding_layer/embeddings"], 1, PADDING)
    model = pkg.build_lstm_model([POOL[0], POOL[1], CHANNELS], cfg, UNITS, "inference", dtype="bfloat16", device=dev)
    model.set_weights(w)
    h_boxes = torch.from_numpy(boxes_np).pin_memory()
    d_boxes = torch.empty((n_batches * VG_BATCH, VG_ROIS, 4), device=dev)
    h_tok = torch.empty((n_batches * VG_BATCH * VG_ROIS, PADDING), dtype=torch.int32).pin_memory()
    views = [[f[(i * VG_BATCH) % VG_POOL:(i * VG_BATCH) % VG_POOL + VG_BATCH] for f in pool] for i in range(VG_POOL // VG_BATCH)]

    def job(host_io):
        """boxes (pinned host) -> device, per VG_BATCH-image batch ROIAlign + head + greedy decode, ids -> pinned host"""
        if host_io:
            d_boxes.copy_(h_boxes, non_blocking=True)
        for i in range(n_batches):
            n_img = min(VG_BATCH, len(my_images) - i * VG_BATCH)
            fm = [f[:n_img] for f in views[i % len(views)]]
            tok = model.caption_rois(d_boxes[i * VG_BATCH:i * VG_BATCH + n_img], fm, IMAGE_SHAPE)
            if host_io:
                h_tok[i * VG_BATCH * VG_ROIS:(i * VG_BATCH + n_img) * VG_ROIS].copy_(tok, non_blocking=True)

    d_boxes.copy_(h_boxes)
    for _ in range(max(1, min(args.warmup, 2))):
        job(False)
    ctx.barrier()
    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    K = max(1, min(args.steps, 3))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    e0.record()
    for _ in range(K):
        job(False)
    e1.record()
    ctx.barrier()
    total_ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.summary()
    ms_per_step = total_ms / K
    total_rois = VG_IMAGES * VG_ROIS
    flops = FLOP_PER_ROI_GREEDY * len(my_images) * VG_ROIS
    achieved = flops / (ms_per_step * 1e-3) / 1e12
    ctx.barrier()
    t0 = time.perf_counter()
    job(True)
    ctx.barrier()
    e2e_s = ctx.max_over_ranks(time.perf_counter() - t0)
    return {
        "metric": "roi_captions_per_sec", "value": round(total_rois / (ms_per_step * 1e-3), 1), "unit": "RoI captions/s",
        "n_gpus": world, "steps": K, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": round(ms_per_step, 3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "cfg5 dense captioning: %d images x %d RoIs dealt round-robin over the ranks, batches of %d "
                               "images; FPN pyramids resident on the device (a pool of %d distinct synthetic pyramids stands "
                               "in for the backbone output), PyramidROIAlign -> head -> v1 greedy P=%d"
                               % (VG_IMAGES, VG_ROIS, VG_BATCH, VG_POOL, PADDING),
                   "rois_per_step": total_rois, "images_per_rank": len(my_images), "sharding": "round-robin images, no collective",
                   "l2": "inputs larger than L2 (pyramid pool %d MB)" % (sum(f.numel() for f in pool) * 4 // 2 ** 20)},
        "clocks": clocks,
        "gpu_launches": K * n_batches * (9 if os.environ.get("DCAP_GREEDY_LOOP", "1") != "0" else 3 + 7 + 5 * PADDING),
        "roofline": tensor_roofline(achieved, total_ms * 1e-3, kernel="gemm_bf16_tc2_kernel (decoder GEMMs; whole job time incl. ROIAlign)",
                                    algorithmic_flops_per_step_per_rank=flops),
        "e2e": {"value": round(total_rois / e2e_s, 1), "unit": "RoI captions/s", "h2d_bytes_per_step": int(h_boxes.numel() * 4) * world,
                "d2h_bytes_per_step": int(h_tok.numel() * 4) * world, "steps": 1,
                "api": "RoiCaptionModel.caption_rois per %d-image batch (pinned host boxes in, token ids out to pinned host)" % VG_BATCH},
    }


SUB_KEYS = ("metric", "value", "unit", "ms_per_step", "steps", "scaling", "dtype", "config", "clocks", "roofline",
            "roofline_hbm", "e2e", "gpu_launches", "breakdown")


def sub_records(args, ctx):
    """Bounded runs of the other BASELINE configs appended to the default line, so that the driver's N = 1/2/4/8
    sweep also records cfg2 (fp32 ROIAlign), cfg3 (the path's only collective: the NCCL gradient all-reduce),
    cfg4 (beam, rows split over ranks) and cfg5 (5000 x 300 end to end)."""
    import copy
    out = {}
    plan = [("roi_features", run_ours_roi_features, dict(steps=20, warmup=3, e2e_steps=4)),
            ("train", run_ours_train, dict(steps=10, warmup=3, e2e_steps=8)),       # 8 steps: the un-overlapped first upload is 1/8 of the leg
            ("beam", run_ours_beam, dict(steps=1, warmup=1, e2e_steps=1)),
            ("captions_vg", run_ours_captions_vg, dict(steps=2, warmup=1, e2e_steps=1))]
    for name, fn, over in plan:
        a = copy.copy(args)
        for k, v in over.items():
            setattr(a, k, v)
        a.no_cpu_baseline = True
        a.no_e2e = False
        try:
            line = fn(a, ctx)
            out[name] = {k: line[k] for k in SUB_KEYS if k in line}
        except Exception as e:                                      # a sub-record must never take the headline down
            out[name] = {"error": "%s: %s" % (type(e).__name__, e)}
        ctx.torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="captions",
                    choices=["captions", "roi_features", "train", "beam", "proposals", "captions_vg"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="tuning runs only: skip the host-buffer leg")
    ap.add_argument("--no-sub", action="store_true", help="captions only: skip the bounded cfg2/cfg3/cfg4/cfg5 sub-records")
    args = ap.parse_args()
    if args.impl == "reference":
        {"roi_features": run_reference, "proposals": run_reference_proposals}.get(args.workload, run_reference_captions)(args)
        return
    ctx = Ctx(args)
    fn = {"captions": run_ours_captions, "roi_features": run_ours_roi_features, "train": run_ours_train, "beam": run_ours_beam,
          "proposals": run_ours_proposals, "captions_vg": run_ours_captions_vg}[args.workload]
    line = fn(args, ctx)
    if args.workload == "captions" and not args.no_sub and not args.no_e2e:
        line["workloads"] = sub_records(args, ctx)
    if ctx.rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
