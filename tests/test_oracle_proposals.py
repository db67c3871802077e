"""CPU tests of the box front-end oracle (oracle/proposals.py): the vectorised NMS against a literal restatement of
tf.image.non_max_suppression's candidate-vs-selected loop, top-k tie order, padding, clipping."""
import numpy as np

from oracle import proposals as pr

F32 = np.float32


def _literal_tf_nms(boxes, scores, max_out, thr):
    order = np.argsort(-np.asarray(scores, F32), kind="stable")
    sel = []
    for i in order:
        if len(sel) >= max_out:
            break
        if all(not (pr.tf_iou(boxes[i], boxes[j]) > F32(thr)) for j in reversed(sel)):
            sel.append(i)
    return np.array(sel, np.int32)


def test_vectorised_nms_equals_the_literal_loop():
    rng = np.random.default_rng(5)
    for n, thr, mx in ((1, 0.7, 5), (60, 0.7, 100), (150, 0.3, 40), (150, 0.7, 10)):
        c = rng.uniform(0.2, 0.8, (n, 2)); h = rng.uniform(0.0, 0.25, (n, 2))
        b = np.concatenate([c - h, c + h], 1).astype(F32)
        b[::7, 2:] = b[::7, :2]                                     # zero-area boxes: IoU 0, never suppress or get suppressed
        if n > 3:
            b[3] = b[3][[2, 3, 0, 1]]                               # flipped corners are normalised by min/max
        s = np.round(rng.standard_normal(n), 1).astype(F32)         # ties
        assert np.array_equal(pr.tf_non_max_suppression(b, s, mx, thr), _literal_tf_nms(b, s, mx, thr))


def test_top_k_ties_prefer_the_lower_index():
    s = np.array([0.5, 0.9, 0.5, 0.9, 0.1, 0.5], F32)
    assert pr.top_k_indices(s, 4).tolist() == [1, 3, 0, 2]


def test_proposal_layer_shapes_padding_and_range():
    rng = np.random.default_rng(6)
    anchors = pr.generate_pyramid_anchors((32, 64), [0.5, 1, 2], [[16, 16], [8, 8]], [8, 16], 1)
    A = anchors.shape[0]
    probs = rng.uniform(0, 1, (2, A, 2)).astype(F32)
    bbox = (rng.standard_normal((2, A, 4)) * 2).astype(F32)
    out, picked = pr.proposal_layer(probs, bbox, anchors, 50, 0.7, (128, 128, 3), pre_nms_limit=300, return_indices=True)
    assert out.shape == (2, 50, 4) and out.dtype == F32
    assert (out >= 0).all() and (out <= 1).all()
    for b in range(2):
        n = len(picked[b])
        assert not out[b, n:].any()
        sc = probs[b, picked[b], 1]
        assert (np.diff(sc) <= 0).all()                               # score order
        top = set(pr.top_k_indices(probs[b, :, 1], 300).tolist())
        assert set(picked[b].tolist()) <= top
    # fewer anchors than the pre-NMS limit and than proposal_count: everything is a candidate, rest is padding
    out = pr.proposal_layer(probs[:, :20], bbox[:, :20], anchors[:20], 50, 0.7, (128, 128, 3))
    assert out.shape == (2, 50, 4) and not out[:, 20:].any()


def test_tf_nms_restatement_agrees_with_torchvision_nms():
    """torchvision.ops.nms is the same greedy rule (descending score, suppress when IoU > threshold) on (x1,y1,x2,y2)
    boxes: for well-formed boxes with distinct scores the restated tf.image.non_max_suppression must keep the same set
    in the same order (an independent implementation of the part TensorFlow is needed for)."""
    import torch
    from torchvision.ops import nms
    rng = np.random.default_rng(7)
    for n, thr in ((50, 0.7), (300, 0.5), (300, 0.3), (800, 0.7)):
        c = rng.uniform(0.2, 0.8, (n, 2)); h = rng.uniform(0.02, 0.25, (n, 2))
        b = np.concatenate([c - h, c + h], 1).astype(F32)                      # (y1, x1, y2, x2), positive area
        s = rng.permutation(n).astype(F32) / n                                  # distinct scores
        got = pr.tf_non_max_suppression(b, s, n, thr)
        want = nms(torch.from_numpy(b[:, [1, 0, 3, 2]].copy()), torch.from_numpy(s), thr).numpy()
        if not np.array_equal(got, want):
            # fp32 IoU values within an ulp of the threshold may be decided differently by the two formulas
            diff = set(got.tolist()) ^ set(want.tolist())
            assert len(diff) <= 2, (n, thr, sorted(diff))
        assert len(got) > 0
        got_cut = pr.tf_non_max_suppression(b, s, 10, thr)
        assert np.array_equal(got_cut, got[:10])                                # max_output_size only truncates
