"""Generates tests/golden/roi_align_small.npz from the numpy oracle (fixed seed).

The reference cannot be imported in this container (no TensorFlow/Keras), so these vectors pin
the ORACLE, not the reference: they freeze today's oracle output so later edits to the oracle, the
C restatement or the CUDA kernel are all checked against the same bytes.  Run from the repo root:
    python tests/golden/gen_golden_roi_align.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import roi_align as ra  # noqa: E402

rng = np.random.default_rng(20261018)
IMG = 1024          # image_shape used for the level rule; maps are deliberately tiny
FM0 = 32            # P2 side (P3..P5 = 16, 8, 4)
C = 8
B, N = 2, 48
image_shape = (IMG, IMG, 3)
fms = [rng.standard_normal((B, FM0 >> i, FM0 >> i, C), dtype=np.float32) for i in range(4)]
side = np.exp(rng.uniform(np.log(24.0), np.log(700.0), B * N))
asp = np.exp(rng.uniform(np.log(0.5), np.log(2.0), B * N))
h, w = side * np.sqrt(asp) / IMG, side / np.sqrt(asp) / IMG
cy, cx = rng.uniform(0, 1, B * N), rng.uniform(0, 1, B * N)
boxes = np.clip(np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], 1), 0, 1).astype(np.float32)
boxes[::11] = 0.0                                     # zero-padded rows
boxes[5] = [0.25, 0.25, 0.75, 0.75]                   # integer-aligned samples on some levels
boxes[6] = [-0.1, 0.2, 0.5, 1.2]                      # extrapolation (outside [0,1])
boxes[7] = [0.6, 0.6, 0.4, 0.4]                       # inverted box (positive area, reversed walk)
boxes[8] = [0.2, 0.6, 0.5, 0.4]                       # negative area -> NaN level -> 2
boxes = boxes.reshape(B, N, 4)
out, lv = ra.pyramid_roi_align_literal(boxes, fms, (7, 7), image_shape)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "roi_align_small.npz"),
                    boxes=boxes, p2=fms[0], p3=fms[1], p4=fms[2], p5=fms[3],
                    image_shape=np.array(image_shape), pooled=out, levels=lv)
print("levels histogram", np.bincount(lv.ravel(), minlength=6)[2:], "bytes",
      os.path.getsize(os.path.join(ROOT, "tests", "golden", "roi_align_small.npz")))
