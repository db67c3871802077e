"""Oracle self-consistency (CPU): numpy oracle vs golden vectors, scalar-loop transcription,
the C restatement, and two independent transcriptions of tf.image.crop_and_resize (TVM's python
reference and torch grid_sample).  Reference lines: evaluate_models/modified_dense_model.py:313-419.
"""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import roi_align as ra
from tests import _c_oracle


def _golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "roi_align_small.npz"))
    fms = [g["p2"], g["p3"], g["p4"], g["p5"]]
    return g, fms, tuple(int(v) for v in g["image_shape"])


def test_golden_literal_and_direct(golden_dir):
    g, fms, ishape = _golden(golden_dir)
    lit, lv = ra.pyramid_roi_align_literal(g["boxes"], fms, (7, 7), ishape)
    assert lit.shape == (1, 96, 7, 7, 8)
    assert np.array_equal(lv, g["levels"])
    assert np.array_equal(lit.view(np.uint32), g["pooled"].view(np.uint32))
    direct, lv2 = ra.pyramid_roi_align(g["boxes"], fms, (7, 7), ishape)
    assert np.array_equal(direct.view(np.uint32), lit.view(np.uint32))
    assert np.array_equal(lv2, lv)
    assert set(np.unique(lv)) == {2, 3, 4, 5}


def test_c_oracle_bit_exact(golden_dir):
    g, fms, ishape = _golden(golden_dir)
    for literal in (False, True):
        out, lv = _c_oracle.pyramid_roi_align(g["boxes"], fms, (7, 7), ishape, literal=literal)
        assert np.array_equal(lv, g["levels"])
        assert np.array_equal(out.view(np.uint32), g["pooled"][0].view(np.uint32))


def test_vectorised_equals_loops():
    rng = np.random.default_rng(3)
    fm = rng.standard_normal((2, 9, 13, 4), dtype=np.float32)
    boxes = rng.uniform(-0.2, 1.2, (20, 4)).astype(np.float32)
    idx = rng.integers(0, 2, 20)
    for cs in [(7, 7), (1, 1), (3, 5), (14, 14)]:
        a = ra.crop_and_resize(fm, boxes, idx, cs)
        b = ra.crop_and_resize_loops(fm, boxes, idx, cs)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), cs


def test_against_tvm_transcription():
    path = None
    try:
        import tilelang  # noqa: F401
        base = os.path.dirname(tilelang.__file__)
        path = os.path.join(base, "3rdparty/tvm/python/tvm/topi/testing/crop_and_resize_python.py")
    except Exception:
        pass
    if not path or not os.path.exists(path):
        pytest.skip("TVM crop_and_resize_python transcription not present in this image")
    spec = importlib.util.spec_from_file_location("_tvm_car", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(5)
    fm = rng.standard_normal((2, 11, 7, 3), dtype=np.float32)
    boxes = rng.uniform(-0.1, 1.1, (12, 4)).astype(np.float32)
    idx = rng.integers(0, 2, 12)
    want = mod.crop_and_resize_python(fm, boxes, idx, (7, 7), "NHWC")
    got = ra.crop_and_resize(fm, boxes, idx, (7, 7))
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)


def test_against_grid_sample_in_range():
    """grid_sample(align_corners=True) equals crop_and_resize for samples inside the map."""
    rng = np.random.default_rng(7)
    H, W, C = 16, 12, 5
    fm = rng.standard_normal((1, H, W, C), dtype=np.float32)
    boxes = np.sort(rng.uniform(0.0, 1.0, (10, 2, 2)), axis=1).reshape(10, 4)[:, [0, 1, 2, 3]]
    boxes = boxes.astype(np.float32)       # (y1,x1,y2,x2) with y1<=y2, x1<=x2, all inside [0,1]
    got = ra.crop_and_resize(fm, boxes, np.zeros(10, int), (7, 7))
    t = torch.from_numpy(fm).permute(0, 3, 1, 2).double()
    for n in range(10):
        y1, x1, y2, x2 = boxes[n].astype(np.float64)
        ys = torch.linspace(0, 1, 7, dtype=torch.float64) * (y2 - y1) + y1
        xs = torch.linspace(0, 1, 7, dtype=torch.float64) * (x2 - x1) + x1
        gy, gx = torch.meshgrid(ys * 2 - 1, xs * 2 - 1, indexing="ij")
        grid = torch.stack([gx, gy], -1)[None]
        want = torch.nn.functional.grid_sample(t, grid, mode="bilinear", align_corners=True)
        want = want[0].permute(1, 2, 0).numpy()
        np.testing.assert_allclose(got[n], want, rtol=1e-4, atol=1e-4)


def test_level_edge_cases():
    ishape = (1024, 1024, 3)
    b = np.array([[0, 0, 0, 0],                 # zero area -> log(0) = -inf -> level 2
                  [0.2, 0.6, 0.5, 0.4],         # negative area -> NaN -> level 2
                  [0.1, 0.1, 0.1 + 224 / 1024, 0.1 + 224 / 1024],     # 224 px -> P4
                  [0.0, 0.0, 1.0, 1.0],         # 1024 px -> clamp 5
                  [0.5, 0.5, 0.5 + 8 / 1024, 0.5 + 8 / 1024],         # 8 px -> clamp 2
                  [0.0, 0.0, 112 / 1024, 112 / 1024],                 # 112 px -> P3
                  [0.0, 0.0, 448 / 1024, 448 / 1024]], np.float32)    # 448 px -> P5
    lv = ra.fpn_level(b, ishape)
    assert lv.tolist() == [2, 2, 4, 5, 2, 3, 5]
    assert _c_oracle.fpn_levels(b, ishape).tolist() == lv.tolist()


def test_level_boundaries_agree_with_c():
    """sqrt(area) = 224*2^(k+0.5): the round-half-even boundary; numpy and C must agree."""
    rng = np.random.default_rng(11)
    ks = rng.integers(-3, 2, 4000)
    side = 224.0 * 2.0 ** (ks + 0.5) / 1024.0 * (1 + rng.integers(-3, 4, 4000) * 2.0 ** -23)
    asp = np.exp(rng.uniform(-0.5, 0.5, 4000))
    b = np.zeros((4000, 4), np.float32)
    b[:, 2] = side * asp
    b[:, 3] = side / asp
    a = ra.fpn_level(b, (1024, 1024, 3))
    c = _c_oracle.fpn_levels(b, (1024, 1024, 3))
    assert np.array_equal(a, c)
    assert ra.level_ambiguity(b, (1024, 1024, 3)).any()


def test_zero_box_reads_p2_origin(golden_dir):
    g, fms, ishape = _golden(golden_dir)
    out, lv = ra.pyramid_roi_align(np.zeros((2, 3, 4), np.float32), fms, (7, 7), ishape)
    assert (lv == 2).all()
    for b in range(2):
        assert np.array_equal(out[0, 3 * b:3 * b + 3], np.broadcast_to(fms[0][b, 0, 0], (3, 7, 7, 8)))


def test_empty_and_limits(golden_dir):
    g, fms, ishape = _golden(golden_dir)
    out, lv = ra.pyramid_roi_align(np.zeros((2, 0, 4), np.float32), fms, (7, 7), ishape)
    assert out.shape == (1, 0, 7, 7, 8) and lv.shape == (2, 0)
    with pytest.raises(ValueError):
        _c_oracle.pyramid_roi_align(np.zeros((1, 100001, 4), np.float32), [f[:1] for f in fms],
                                    (1, 1), ishape, literal=True)


def test_unique_tap_pixels_small():
    fm_shapes = [(8, 8), (4, 4), (2, 2), (1, 1)]
    boxes = np.zeros((1, 2, 4), np.float32)                       # two zero boxes: 1 pixel of P2
    assert ra.unique_tap_pixels(boxes, fm_shapes, (7, 7), (1024, 1024, 3)) == 1
    boxes[0, 1] = [0, 0, 1, 1]                                    # full image -> P5 (1x1 map)
    assert ra.unique_tap_pixels(boxes, fm_shapes, (7, 7), (1024, 1024, 3)) == 2


def test_backward_oracle_is_the_adjoint_of_the_forward():
    """<g, align(F)> == <backward(g), F>: the restated CropAndResizeGradImage is the transpose of the gather."""
    rng = np.random.default_rng(5)
    B, N, C = 1, 12, 4
    from image_captioning_b200 import synth
    boxes = synth.synth_boxes(rng, B, N, 1024.0)
    shapes = [(16 >> i, 16 >> i) for i in range(4)]
    fms = [rng.standard_normal((B, h, w, C)).astype(np.float32) for h, w in shapes]
    g = rng.standard_normal((B * N, 7, 7, C)).astype(np.float32)
    out, _ = ra.pyramid_roi_align(boxes, fms, (7, 7), (1024, 1024, 3))
    grads = ra.pyramid_roi_align_backward(boxes, g, shapes, (7, 7), (1024, 1024, 3))
    lhs = float((out[0].astype(np.float64) * g).sum())
    rhs = float(sum((gr * f.astype(np.float64)).sum() for gr, f in zip(grads, fms)))
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))
