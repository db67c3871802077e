"""The oracle against outputs of the REFERENCE'S OWN numpy code (tests/golden/reference_numpy.npz, produced by
tests/golden/gen_golden_reference_numpy.py, which imports /root/reference with TensorFlow stubbed out).
These are the pinned parts of the oracle: post-processing NMS / refine_generations, the anchor generator and
the box-delta arithmetic.  CPU only."""
import hashlib
import os

import numpy as np
import pytest

from oracle import postprocess as pp
from oracle import proposals as pr

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_numpy.npz"))
N_PP = 5


@pytest.mark.parametrize("i", range(N_PP))
def test_nms_and_refine_generations_equal_the_reference(i):
    boxes, probs, scores = G["pp%d_boxes" % i], G["pp%d_probs" % i], G["pp%d_scores" % i]
    thr, mx = G["pp%d_params" % i]
    got_scores = pp.caption_scores(probs)
    assert np.array_equal(got_scores.view(np.uint32), scores.view(np.uint32))
    assert np.array_equal(pp.non_max_suppression(boxes, scores, thr), G["pp%d_nms_keep" % i])
    keep = pp.refine_generations(boxes, scores, thr, int(mx))
    assert np.array_equal(boxes[keep], G["pp%d_refined_rois" % i])
    assert np.array_equal(probs[keep][:, 0, :], G["pp%d_refined_first_probs" % i])


def test_dice_overlap_equals_the_reference():
    b = G["pp1_boxes"]
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    got = pp.compute_overlap(b[0], b, area[0], area)
    assert np.array_equal(got.view(np.uint32), G["iou_17_row0"].view(np.uint32))
    for j in range(5):
        np.testing.assert_array_equal(pp.compute_overlap(b[j], b, area[j], area).astype(np.float64), G["overlaps_17"][:, j])


def test_anchor_generator_equals_the_reference():
    scales, ratios, strides = (32, 64, 128, 256, 512), [0.5, 1, 2], [4, 8, 16, 32, 64]
    small = [[-(-128 // s)] * 2 for s in strides]
    full = [[-(-1024 // s)] * 2 for s in strides]
    assert np.array_equal(pr.generate_pyramid_anchors(scales, ratios, small, strides, 1), G["anchors_128"])
    assert np.array_equal(pr.generate_pyramid_anchors(scales[:2], ratios, small[:2], strides[:2], 2), G["anchors_128_stride2"])
    a = pr.generate_pyramid_anchors(scales, ratios, full, strides, 1)
    assert list(a.shape) == list(G["anchors_1024_shape"]) == [261888, 4]
    digest = hashlib.sha256(np.ascontiguousarray(a.astype(np.float32)).tobytes()).digest()
    assert digest == G["anchors_1024_sha256"].tobytes()
    # the product's host-side generator (image_captioning_b200.proposals) is the same function of the config
    from image_captioning_b200 import proposals as prod
    assert np.array_equal(prod.generate_pyramid_anchors(scales, ratios, full, strides, 1), a)


def test_box_deltas_equal_the_reference():
    """Same arithmetic, op by op: with numpy's own fp32 exp the reference's numpy twin is reproduced bit for bit.
    With the correctly rounded exp the oracle defines (numpy's SIMD expf is up to 2 ulp away from it) a corner
    moves by at most a few ulp of the box coordinates."""
    anc = G["anchors_128"].astype(np.float32)
    want = G["refined_128"]
    got = pr.apply_box_deltas(anc, G["deltas_128"], exp=np.exp)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    got = pr.apply_box_deltas(anc, G["deltas_128"])
    scale = np.maximum(np.abs(want).max(1, keepdims=True), 1.0)
    assert (np.abs(got - want) / scale).max() <= 4 * np.finfo(np.float32).eps


def test_beam_search_equals_the_reference_control_flow():
    """gen_captions ("image captioning/test.py":23-64) run with a stand-in model.predict (the oracle's word model):
    candidates, pooling order, stable sort, float64 probability sums and the kept beams equal oracle.beam_v1's."""
    from image_captioning_b200 import synth
    from oracle import decoder as dec
    V, P, K, R = [int(v) for v in G["beam_params"]]
    w = synth.synth_weights_v1(np.random.default_rng(77), V=V, E=6, F=16, U=8, pool=2, C=4, trained_like=False)
    f = dec.head(np.random.default_rng(78).standard_normal((R, 2, 2, 4)).astype(np.float32), w)
    tok, sc = dec.beam_v1(f, w, P, K)
    assert np.array_equal(tok, G["beam_ref_tokens"])
    assert np.array_equal(sc, G["beam_ref_scores"])
    assert len(np.unique(G["beam_ref_tokens"])) > 4                    # not a degenerate fixture


def test_v2_greedy_loop_equals_the_reference_control_flow():
    """The per-RoI loop of evaluate_models/test_score_dense_captions.py:214-225 run with a stand-in model.predict (the
    oracle's v2 inject model): start id 0, P-1 predicts over the pre-padded argmax history == oracle.greedy_v2."""
    from image_captioning_b200 import synth
    from oracle import decoder as dec
    V, P, R = [int(v) for v in G["v2_loop_params"]]
    w = synth.synth_weights_v2(np.random.default_rng(79), V=V, E=6, F=16, units=8, pool=2, C=4, trained_like=False)
    feat = np.random.default_rng(80).standard_normal((R, 2, 2, 4)).astype(np.float32)
    tok, probs = dec.greedy_v2(feat, w, P)
    want = G["v2_loop_probs"]
    assert probs.shape == want.shape == (R, P - 1, V)
    assert np.array_equal(tok, want.argmax(-1))
    np.testing.assert_allclose(probs, want, rtol=2e-6, atol=1e-9)     # batch-of-R vs batch-of-1 matmuls


def test_pyramid_roi_align_glue_equals_the_reference_layer():
    """PyramidROIAlign.call (evaluate_models/modified_dense_model.py:342-416) EXECUTED from the reference source over a
    numpy stand-in for its TF ops (tests/golden/tf_numpy_shim.py; crop_and_resize = the oracle's restatement): level
    formula, per-level dispatch, concat and the top_k re-sort reproduce the oracle's literal and direct forms bit for bit."""
    from oracle import roi_align as ra
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "roi_align_small.npz"))
    assert bool(G["roi_align_layer_equals_oracle_golden"])
    fms = [g[k] for k in ("p2", "p3", "p4", "p5")]
    shape = tuple(int(v) for v in g["image_shape"])
    for fn in (ra.pyramid_roi_align_literal, ra.pyramid_roi_align):
        out, _ = fn(g["boxes"], fms, (7, 7), shape)
        assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).digest() == G["roi_align_layer_sha256"].tobytes()


def test_proposal_layer_glue_equals_the_reference_layer():
    """ProposalLayer.call (dense_img_cap_separate_models/modified_dense_model.py:247-303, with apply_box_deltas_graph,
    clip_boxes_graph and utils.batch_slice) EXECUTED from the reference source over the numpy stand-in (top_k order,
    exp and non_max_suppression = the oracle's restatements): std-dev scaling, gather order, delta arithmetic, clip,
    normalisation, NMS call and zero padding reproduce oracle.proposal_layer bit for bit."""
    lo, hi = G["proposal_anchor_slice"]
    anchors = G["anchors_128"][lo:hi]
    want = G["proposal_ref"]
    got = pr.proposal_layer(G["proposal_probs"], G["proposal_bbox"], anchors, want.shape[1], 0.7, (128, 128, 3))
    assert got.dtype == want.dtype and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    n_valid = (np.abs(want).sum(-1) > 0).sum(1)
    assert (n_valid > 100).all() and (n_valid < want.shape[1]).all()           # real proposals and real zero padding


def test_v1_model_wiring_equals_the_reference_builders():
    """word_generation_model, ROICaptionInferenceLayer.call (greedy decoding), build_roi_caption_model_training (teacher
    forcing) and roi_caption_loss (text_generation_model.py:130-232, 286-294) EXECUTED from the reference source over eager
    stand-ins for the Keras layers (layer numerics = the oracle's restatements): slicing of [feature | words], the
    [embedding ; feature] and [h2 ; feature] concatenations, mask propagation, the growing post-padded prefix with
    arg-max feedback, the teacher-forced prefixes and the masked mean reproduce the oracle's literal forms bit for bit
    -- and the incremental scans the CUDA path mirrors equal those."""
    from image_captioning_b200 import synth
    from oracle import decoder as dec
    V, P, B, U, F = [int(v) for v in G["v1_params"]]
    w = synth.synth_weights_v1(np.random.default_rng(81), V=V, E=6, F=F, U=U, pool=2, C=4, trained_like=False)
    f = dec.head(np.random.default_rng(82).standard_normal((B, 2, 2, 4)).astype(np.float32), w)
    gt = G["v1_gt"]
    assert gt[1, 2] == 0
    greedy = dec.greedy_v1_literal(f, w, P)
    assert np.array_equal(greedy.view(np.uint32), G["v1_greedy_probs"].view(np.uint32))
    tok, probs = dec.greedy_v1(f, w, P)
    assert np.array_equal(tok, G["v1_greedy_probs"].argmax(-1))
    np.testing.assert_allclose(probs, G["v1_greedy_probs"], rtol=1e-5, atol=1e-8)
    train = dec.train_forward_v1_literal(f, gt, w)
    assert np.array_equal(train.view(np.uint32), G["v1_train_probs"].view(np.uint32))
    np.testing.assert_allclose(dec.train_forward_v1(f, gt, w), G["v1_train_probs"], rtol=1e-5, atol=1e-8)
    ids = dec.targets_from_captions(gt)
    valid = np.ones((B, P), bool)
    valid[2, 3:] = False
    loss = dec.roi_caption_loss(ids, train, valid)
    np.testing.assert_allclose(loss, G["v1_loss"], rtol=2e-6)
    assert float(G["v1_loss_all_masked"]) == 0.0 and dec.roi_caption_loss(ids, train, valid & False) == 0.0


def test_whole_models_equal_the_reference_builders():
    """build_lstm_model in both modes (text_generation_model.py:235-283: RoI head as TimeDistributed Conv2D/BatchNorm/ReLU,
    squeeze, then the caption layer / the training graph) and the v2 build_model(inject=True)
    (text_generation_model_v2.py:140-166) EXECUTED from the reference source over the eager Keras stand-ins: the outputs
    equal head -> greedy_v1_literal / train_forward_v1_literal and v2_inject_predict of the oracle."""
    from image_captioning_b200 import synth
    from oracle import decoder as dec
    V, P, B, U, _ = [int(v) for v in G["v1_params"]]
    w = synth.synth_weights_v1(np.random.default_rng(84), V=V, E=6, F=1024, U=U, pool=2, C=4, trained_like=False)
    feat = np.random.default_rng(82).standard_normal((B, 2, 2, 4)).astype(np.float32)
    f = dec.head(feat, w)
    inf = dec.greedy_v1_literal(f, w, P)
    assert inf.shape == G["v1_model_inference"].shape == (B, P, V)
    np.testing.assert_allclose(inf, G["v1_model_inference"], rtol=1e-5, atol=1e-8)
    assert np.array_equal(inf.argmax(-1), G["v1_model_inference"].argmax(-1))
    np.testing.assert_allclose(dec.train_forward_v1_literal(f, G["v1_gt"], w), G["v1_model_training"], rtol=1e-5, atol=1e-8)
    V2, P2, R2 = [int(v) for v in G["v2_loop_params"]]
    w2 = synth.synth_weights_v2(np.random.default_rng(85), V=V2, E=6, F=1024, units=8, pool=2, C=4, trained_like=False)
    feat2 = np.random.default_rng(80).standard_normal((R2, 2, 2, 4)).astype(np.float32)
    got = dec.v2_inject_predict(feat2, G["v2_model_words"], w2)
    assert got.shape == G["v2_model_probs"].shape == (R2, V2)
    np.testing.assert_allclose(got, G["v2_model_probs"], rtol=1e-5, atol=1e-8)


def test_v2_greedy_loop_from_a_given_first_word_equals_the_reference():
    """evaluate_models/eval_text_generation_model_v2.py:176-186 (prev = [gt[0]], then P-1 predicts) run with the stand-in
    predict(): the decoded ids equal oracle.greedy_v2(start=...)."""
    from image_captioning_b200 import synth
    from oracle import decoder as dec
    V, P, R = [int(v) for v in G["v2_loop_params"]]
    w = synth.synth_weights_v2(np.random.default_rng(79), V=V, E=6, F=16, units=8, pool=2, C=4, trained_like=False)
    feat = np.random.default_rng(80).standard_normal((R, 2, 2, 4)).astype(np.float32)
    tok, _ = dec.greedy_v2(feat, w, P, start=G["v2_eval_start"])
    want = G["v2_eval_predicted"]
    assert want.shape == (R, P) and np.array_equal(want[:, 0], G["v2_eval_start"])
    assert np.array_equal(tok, want[:, 1:])
