"""Property tests (hypothesis) of the oracle on the edge cases SURVEY.md section 8c lists: zero-area / padded
boxes, boxes touching or exceeding [0,1], integer-aligned samples, level-boundary boxes, token 0 mid-sequence,
arg-max ties.  CPU only; the CUDA kernels are compared with the same oracle in the -m gpu tests."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import decoder as dec
from oracle import roi_align as ra
from oracle import postprocess as pp

F32 = np.float32
coord = st.floats(min_value=-0.25, max_value=1.25, allow_nan=False, width=32)


def _maps(rng, C=4, side=16):
    return [rng.standard_normal((1, side >> i, side >> i, C)).astype(F32) for i in range(4)]


@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(coord, coord, coord, coord), min_size=1, max_size=6), st.integers(0, 2 ** 31 - 1))
def test_literal_and_direct_forms_agree_on_arbitrary_boxes(raw, seed):
    """The reference's 4x crop_and_resize + concat + top_k re-sort + gather equals the one-pass form for ANY boxes:
    reversed (negative area -> NaN level -> 2), outside the image (extrapolation 0), degenerate."""
    rng = np.random.default_rng(seed)
    boxes = np.array(raw, F32)[None]
    fms = _maps(rng)
    lit, lv_lit = ra.pyramid_roi_align_literal(boxes, fms, (7, 7), (1024, 1024, 3))
    direct, lv = ra.pyramid_roi_align(boxes, fms, (7, 7), (1024, 1024, 3))
    assert lit.shape == direct.shape == (1, boxes.shape[1], 7, 7, 4)
    assert np.array_equal(lit.view(np.uint32), direct.view(np.uint32))
    assert ((lv >= 2) & (lv <= 5)).all() and np.array_equal(lv, lv_lit)


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 2 ** 31 - 1))
def test_zero_boxes_read_the_p2_origin_and_outside_samples_are_zero(seed):
    rng = np.random.default_rng(seed)
    fms = _maps(rng)
    boxes = np.zeros((1, 3, 4), F32)
    boxes[0, 1] = [1.5, 1.5, 2.0, 2.0]                      # entirely outside: every sample extrapolates to 0
    boxes[0, 2] = [0.0, 0.0, 1.0, 1.0]                      # integer-aligned end points (floor == ceil at the corners)
    out, lv = ra.pyramid_roi_align(boxes, fms, (7, 7), (1024, 1024, 3))
    assert lv[0, 0] == 2 and np.array_equal(out[0, 0], np.broadcast_to(fms[0][0, 0, 0], (7, 7, 4)))
    assert not out[0, 1].any()
    li = int(lv[0, 2]) - 2
    H = fms[li].shape[1]
    assert np.array_equal(out[0, 2][0, 0], fms[li][0, 0, 0]) and np.array_equal(out[0, 2][6, 6], fms[li][0, H - 1, H - 1])


@settings(max_examples=60, deadline=None)
@given(st.integers(-3, 3), st.floats(min_value=-0.49, max_value=0.49, allow_nan=False),
       st.floats(min_value=0.5, max_value=2.0, allow_nan=False))
def test_level_rule_away_from_the_rounding_boundaries(k, frac, aspect):
    """sqrt(area) = 224 * 2^(k + frac) px -> level clamp(4 + k, 2, 5) when frac is inside (-0.5, 0.5)."""
    side = 224.0 * 2.0 ** (k + frac) / 1024.0
    h, w = side * np.sqrt(aspect), side / np.sqrt(aspect)
    box = np.array([[[0.0, 0.0, h, w]]], F32)
    if abs(abs(frac) - 0.5) < 1e-3:
        return
    assert int(ra.fpn_level(box, (1024, 1024, 3))[0, 0]) == min(5, max(2, 4 + k))


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.lists(st.integers(0, 11), min_size=5, max_size=5))
def test_token_zero_anywhere_carries_the_state(seed, toks):
    """Literal O(P^2) prefix re-runs == incremental masked scan for ANY token sequence, zeros included."""
    from image_captioning_b200 import synth
    rng = np.random.default_rng(seed)
    w = synth.synth_weights_v1(rng, V=12, E=6, F=1024, U=8, pool=2, C=4)
    f = dec.head(rng.standard_normal((2, 2, 2, 4)).astype(F32), w)
    gt = np.array([toks, toks[::-1]], F32)
    lit = dec.train_forward_v1_literal(f, gt, w)
    inc = dec.train_forward_v1(f, gt, w)
    np.testing.assert_allclose(lit, inc, rtol=1e-5, atol=1e-7)
    for b in range(2):
        for t in range(1, 5):
            if gt[b, t] == 0:
                np.testing.assert_allclose(inc[b, t], inc[b, t - 1], rtol=1e-6)


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 40), st.floats(min_value=0.05, max_value=0.95))
def test_nms_survivors_never_overlap_more_than_the_threshold(seed, n, thr):
    rng = np.random.default_rng(seed)
    c = rng.uniform(0.2, 0.8, (n, 2)); h = rng.uniform(0.0, 0.3, (n, 2))
    b = np.concatenate([c - h, c + h], 1).astype(F32)
    s = np.round(rng.standard_normal(n), 1).astype(F32)              # coarse scores: many ties
    keep = pp.non_max_suppression(b, s, thr)
    assert len(set(keep.tolist())) == len(keep) and (np.diff(s[keep]) <= 0).all()
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    for i, k in enumerate(keep[1:], 1):
        ov = pp.compute_overlap(b[k], b[keep[:i]], area[k], area[keep[:i]])
        assert not (ov > F32(thr)).any()
    removed = sorted(set(range(n)) - set(keep.tolist()))
    for r in removed:                                                  # every removed box is covered by a better survivor
        better = [k for k in keep if s[k] >= s[r]]
        ov = pp.compute_overlap(b[r], b[better], area[r], area[better])
        assert (ov > F32(thr)).any()


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 60), st.floats(min_value=0.0, max_value=0.95), st.integers(1, 70))
def test_tf_nms_invariants(seed, n, thr, max_out):
    """tf.image.non_max_suppression restated: survivors are in score order, pairwise IoU <= thr, and every candidate that
    was passed over before the output filled up overlaps an earlier survivor by more than thr."""
    from oracle import proposals as pr
    rng = np.random.default_rng(seed)
    c = rng.uniform(0.2, 0.8, (n, 2)); h = rng.uniform(0.0, 0.3, (n, 2))
    b = np.concatenate([c - h, c + h], 1).astype(F32)
    b[::6, 2:] = b[::6, :2]                                            # empty boxes: IoU 0 with everything
    s = np.round(rng.standard_normal(n), 1).astype(F32)
    keep = pr.tf_non_max_suppression(b, s, max_out, thr)
    assert len(keep) <= max_out and len(set(keep.tolist())) == len(keep) and (np.diff(s[keep]) <= 0).all()
    for i, k in enumerate(keep):
        for j in keep[:i]:
            assert not pr.tf_iou(b[j], b[k]) > F32(thr)
    order = np.argsort(-s, kind="stable")
    last = int(np.nonzero(order == keep[-1])[0][0])
    kept = set(keep.tolist())
    for pos in range(last):
        r = order[pos]
        if r not in kept:
            earlier = [k for k in keep if int(np.nonzero(order == k)[0][0]) < pos]
            assert any(pr.tf_iou(b[k], b[r]) > F32(thr) for k in earlier)


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 40), st.integers(1, 30))
def test_proposal_layer_outputs_are_clipped_ordered_and_padded(seed, count, limit):
    from oracle import proposals as pr
    rng = np.random.default_rng(seed)
    anchors = pr.generate_pyramid_anchors((32, 64), [0.5, 1, 2], [[8, 8], [4, 4]], [16, 32], 1)
    A = anchors.shape[0]
    probs = rng.uniform(0, 1, (1, A, 2)).astype(F32)
    bbox = (rng.standard_normal((1, A, 4)) * 3).astype(F32)
    out, picked = pr.proposal_layer(probs, bbox, anchors, count, 0.7, (128, 128, 3), pre_nms_limit=limit, return_indices=True)
    n = len(picked[0])
    assert out.shape == (1, count, 4) and n <= min(count, limit, A)
    assert (out >= 0).all() and (out <= 1).all() and not out[0, n:].any()
    assert (out[0, :n, 2] >= out[0, :n, 0]).all() and (out[0, :n, 3] >= out[0, :n, 1]).all()    # exp > 0: never inverted
    top = pr.top_k_indices(probs[0, :, 1], limit)
    assert set(picked[0].tolist()) <= set(top.tolist()) and (np.diff(probs[0, picked[0], 1]) <= 0).all()
