"""Deterministic synthetic inputs shared by tests, smoke() and bench.py.

Pure input generation (numpy RNG draws in documented distributions); no reference arithmetic lives
here.  Seeds are fixed per BASELINE config: 1000 + config index.
"""
import numpy as np

F32 = np.float32


def synth_boxes(rng, n_images, n_boxes, image_size=1024.0, pad_frac=0.02, straddle_frac=0.01):
    """sqrt(area) log-uniform in [32,512] px, aspect log-uniform [0.5,2], centre uniform,
    clipped to [0,1]; ``pad_frac`` zero rows (padding) and ``straddle_frac`` boxes that cross
    the border before clipping."""
    n = n_images * n_boxes
    side = np.exp(rng.uniform(np.log(32.0), np.log(512.0), n))
    aspect = np.exp(rng.uniform(np.log(0.5), np.log(2.0), n))
    h = side * np.sqrt(aspect) / image_size
    w = side / np.sqrt(aspect) / image_size
    cy = rng.uniform(0, 1, n)
    cx = rng.uniform(0, 1, n)
    inside = rng.uniform(0, 1, n) >= straddle_frac
    # boxes that must not straddle are shifted inside the image first
    cy = np.where(inside, np.clip(cy, h / 2, 1 - h / 2), cy)
    cx = np.where(inside, np.clip(cx, w / 2, 1 - w / 2), cx)
    b = np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], 1)
    b = np.clip(b, 0.0, 1.0)
    pad = rng.uniform(0, 1, n) < pad_frac
    b[pad] = 0.0
    return b.astype(F32).reshape(n_images, n_boxes, 4)


def synth_pyramid(rng, n_images, image_size=1024, channels=256):
    """P2..P5 NHWC fp32 N(0,1) maps, H_l = image_size / 2^l."""
    return [rng.standard_normal((n_images, image_size >> l, image_size >> l, channels),
                                dtype=F32) for l in range(2, 6)]
