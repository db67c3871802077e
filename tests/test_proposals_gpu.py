"""GPU parity of the box front-end (dc_proposal_layer / dc_normalize_boxes through the ProposalLayer mirror) against
oracle/proposals.py: bit-exact boxes, identical selection, on the reference's configuration (261888 anchors of a
1024x1024 image, top 6000, NMS 0.7, 1000 / 2000 proposals) and on edge cases (ties, degenerate boxes, all-equal
scores, fewer anchors than the limits)."""
import numpy as np
import pytest
import torch

from oracle import proposals as pr

pytestmark = pytest.mark.gpu
F32 = np.float32


def _rpn_like(rng, B, anchors, tie_frac=0.0):
    """RPN-like outputs: most foreground scores near 0, a few hundred confident ones clustered so that NMS has work."""
    A = anchors.shape[0]
    logit = rng.standard_normal((B, A)) * 2.5 - 4.0
    fg = (1.0 / (1.0 + np.exp(-logit))).astype(F32)
    if tie_frac:
        fg = np.round(fg, 2).astype(F32)                              # heavy ties, also across the top-k threshold
    probs = np.stack([1 - fg, fg], -1).astype(F32)
    bbox = (rng.standard_normal((B, A, 4)) * 1.5).astype(F32)
    return probs, bbox


def _check(layer, probs, bbox, anchors, count, thr, image_shape, limit=6000):
    out, n_valid, index = layer([probs, bbox], return_details=True)
    want, picked = pr.proposal_layer(probs, bbox, anchors, count, thr, image_shape, pre_nms_limit=limit, return_indices=True)
    for b in range(probs.shape[0]):
        n = len(picked[b])
        assert int(n_valid[b]) == n
        assert np.array_equal(index[b, :n], picked[b]), np.nonzero(index[b, :n] != picked[b])[0][:5]
        assert (index[b, n:] == -1).all()
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32))
    return out


def test_reference_configuration_inference_and_training_counts():
    import image_captioning_b200 as pkg
    cfg = pkg.ProposalConfig()
    anchors = cfg.anchors()
    assert anchors.shape == (261888, 4)
    rng = np.random.default_rng(60)
    probs, bbox = _rpn_like(rng, 2, anchors)
    for count in (cfg.POST_NMS_ROIS_INFERENCE, cfg.POST_NMS_ROIS_TRAINING):
        layer = pkg.ProposalLayer(count, cfg.RPN_NMS_THRESHOLD, anchors, cfg)
        assert layer.compute_output_shape() == (None, count, 4)
        out = _check(layer, probs, bbox, anchors, count, 0.7, cfg.IMAGE_SHAPE)
        assert out.shape == (2, count, 4)


def test_ties_at_the_top_k_threshold_take_the_lower_anchor_index():
    import image_captioning_b200 as pkg
    cfg = pkg.ProposalConfig()
    anchors = cfg.anchors()
    rng = np.random.default_rng(61)
    probs, bbox = _rpn_like(rng, 1, anchors, tie_frac=1.0)
    layer = pkg.ProposalLayer(1000, 0.7, anchors, cfg)
    _check(layer, probs, bbox, anchors, 1000, 0.7, cfg.IMAGE_SHAPE)
    # all scores equal: the first 6000 anchors in index order are the candidates
    probs[:] = 0.5
    out, n_valid, index = layer([probs, bbox], return_details=True)
    assert index[0, :int(n_valid[0])].max() < 6000
    _check(layer, probs, bbox, anchors, 1000, 0.7, cfg.IMAGE_SHAPE)


def test_small_pyramids_degenerate_boxes_and_short_inputs():
    import image_captioning_b200 as pkg
    rng = np.random.default_rng(62)
    cfg = pkg.ProposalConfig(IMAGE_MAX_DIM=128)
    anchors = cfg.anchors()                                           # 4092 anchors < 6000: every anchor is a candidate
    assert anchors.shape[0] == 4092
    probs, bbox = _rpn_like(rng, 3, anchors)
    bbox[0, ::5] = [0.0, 0.0, -80.0, -80.0]                           # exp(-16): zero-area after rounding/clipping
    bbox[1, ::3, :2] = 50.0                                           # pushed outside the image: clipped to a corner (area 0)
    probs[2, 7, 1] = np.nan                                           # NaN sorts last
    probs[2, 9, 1] = -0.0
    for count, thr in ((100, 0.7), (5000, 0.3), (64, 0.0)):
        layer = pkg.ProposalLayer(count, thr, anchors, cfg)
        p = probs.copy()
        want_probs = np.where(np.isnan(p), -np.inf, p)                # the oracle's argsort would put NaN first
        out, n_valid, index = layer([p, bbox], return_details=True)
        want, picked = pr.proposal_layer(want_probs, bbox, anchors, count, thr, cfg.IMAGE_SHAPE, return_indices=True)
        assert np.array_equal(out.view(np.uint32), want.view(np.uint32))
        for b in range(3):
            assert np.array_equal(index[b, :len(picked[b])], picked[b])
    # a handful of anchors, fewer than proposal_count
    layer = pkg.ProposalLayer(50, 0.7, anchors[:20], cfg)
    _check(layer, probs[:, :20].copy(), bbox[:, :20].copy(), anchors[:20], 50, 0.7, cfg.IMAGE_SHAPE)
    layer = pkg.ProposalLayer(3, 0.7, anchors[:1], cfg)
    _check(layer, probs[:1, :1].copy(), bbox[:1, :1].copy(), anchors[:1], 3, 0.7, cfg.IMAGE_SHAPE)


def test_overflowing_deltas_propagate_inf_and_nan_like_the_graph():
    """exp overflow gives +inf sizes and inf - inf = NaN corners; tf.minimum / tf.maximum (and numpy) propagate the
    NaN through the clip, and a NaN box is never suppressed and never suppresses."""
    import image_captioning_b200 as pkg
    rng = np.random.default_rng(66)
    cfg = pkg.ProposalConfig(IMAGE_MAX_DIM=128)
    anchors = cfg.anchors()
    probs, bbox = _rpn_like(rng, 1, anchors)
    bbox[0, ::11, 2] = 1000.0                                         # exp(200) = inf in fp32
    bbox[0, 5::11, 3] = -1000.0                                       # exp(-200) = 0
    layer = pkg.ProposalLayer(300, 0.7, anchors, cfg)
    out, n_valid, index = layer([probs, bbox], return_details=True)
    with np.errstate(all="ignore"):
        want, picked = pr.proposal_layer(probs, bbox, anchors, 300, 0.7, cfg.IMAGE_SHAPE, return_indices=True)
    assert np.isnan(want).any()
    assert np.array_equal(index[0, :len(picked[0])], picked[0])
    assert np.array_equal(np.isnan(out), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.array_equal(out[ok].view(np.uint32), want[ok].view(np.uint32))


def test_anchor_counts_that_do_not_fit_the_shared_memory_cache_use_the_streaming_path():
    import image_captioning_b200 as pkg
    rng = np.random.default_rng(63)
    cfg = pkg.ProposalConfig(IMAGE_MAX_DIM=1280)                      # 409200 anchors > 8 * 49152
    anchors = cfg.anchors()
    assert anchors.shape[0] > 8 * 48 * 1024
    probs, bbox = _rpn_like(rng, 1, anchors)
    layer = pkg.ProposalLayer(500, 0.7, anchors, cfg)
    _check(layer, probs, bbox, anchors, 500, 0.7, cfg.IMAGE_SHAPE)


def test_device_tensors_stay_on_the_device_and_feed_roi_align():
    import image_captioning_b200 as pkg
    cfg = pkg.ProposalConfig()
    anchors = cfg.anchors()
    rng = np.random.default_rng(64)
    probs, bbox = _rpn_like(rng, 1, anchors)
    layer = pkg.ProposalLayer(200, 0.7, anchors, cfg)
    rois = layer([torch.from_numpy(probs).cuda(), torch.from_numpy(bbox).cuda()])
    assert rois.is_cuda and rois.shape == (1, 200, 4)
    want = pr.proposal_layer(probs, bbox, anchors, 200, 0.7, cfg.IMAGE_SHAPE)
    assert np.array_equal(rois.cpu().numpy().view(np.uint32), want.view(np.uint32))
    from oracle import roi_align as ra
    fms = [rng.standard_normal((1, 64 >> i, 64 >> i, 64)).astype(F32) for i in range(4)]
    pooled = pkg.pyramid_roi_align(rois, [torch.from_numpy(f).cuda() for f in fms], (7, 7), (1024, 1024, 3))
    ref, _ = ra.pyramid_roi_align(want, fms, (7, 7), (1024, 1024, 3))
    assert np.array_equal(pooled.float().cpu().numpy().reshape(ref.shape).view(np.uint32), ref.view(np.uint32))


def test_generated_rois_through_the_caption_pipeline():
    """use_generated_rois=True end to end on the device: RPN outputs -> ProposalLayer -> PyramidROIAlign -> head -> greedy
    decoding (fp32 model), every stage bit-exact, so the token ids equal the oracle chain's."""
    import image_captioning_b200 as pkg
    from image_captioning_b200 import synth
    from oracle import decoder as dec
    from oracle import roi_align as ra
    rng = np.random.default_rng(67)
    cfg = pkg.ProposalConfig(IMAGE_MAX_DIM=256)
    anchors = cfg.anchors()
    probs, bbox = _rpn_like(rng, 2, anchors)
    layer = pkg.ProposalLayer(40, 0.7, anchors, cfg)
    rois = layer([torch.from_numpy(probs).cuda(), torch.from_numpy(bbox).cuda()])
    C, V, P = 64, 500, 6
    fms = [rng.standard_normal((2, 64 >> i, 64 >> i, C)).astype(F32) for i in range(4)]
    w = synth.synth_weights_v1(np.random.default_rng(68), V=V, E=48, U=128, C=C)
    mcfg = pkg.DenseCapConfig(V, w["imgcap_embedding_layer/embeddings"], 1, P)
    model = pkg.build_lstm_model([7, 7, C], mcfg, 128, "inference", dtype="float32")
    model.set_weights(w)
    tok = model.caption_rois(rois, [torch.from_numpy(f).cuda() for f in fms], tuple(cfg.IMAGE_SHAPE)).cpu().numpy()
    rois_want = pr.proposal_layer(probs, bbox, anchors, 40, 0.7, cfg.IMAGE_SHAPE)
    pooled, _ = ra.pyramid_roi_align(rois_want, fms, (7, 7), tuple(cfg.IMAGE_SHAPE))
    tok_want, _ = dec.greedy_v1(dec.head(pooled[0], w), w, P)
    assert tok.shape == tok_want.shape == (80, P)
    assert np.array_equal(tok, tok_want)


def test_normalize_boxes_is_an_fp32_division():
    import image_captioning_b200 as pkg
    rng = np.random.default_rng(65)
    px = rng.uniform(-20, 1100, (3, 77, 4)).astype(F32)
    got = pkg.normalize_boxes(px, (1024, 768, 3))
    want = pr.normalize_boxes(px, 1024, 768)
    assert got.shape == px.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert pkg.normalize_boxes(px[:0], (1024, 768, 3)).shape == (0, 77, 4)
