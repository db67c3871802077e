"""Caption post-processing: oracle self-checks on CPU, device
``refine_generations`` and fused caption scores against the oracle on the GPU."""
import numpy as np
import pytest

from oracle import postprocess as pp


def _boxes(rng, n):
    c = rng.uniform(0.1, 0.9, (n, 2))
    h = rng.uniform(0.02, 0.3, (n, 2))
    b = np.concatenate([c - h, c + h], 1).astype(np.float32)
    return np.clip(b, 0, 1)


def test_oracle_nms_basic_properties():
    rng = np.random.default_rng(0)
    b = _boxes(rng, 60)
    s = rng.standard_normal(60).astype(np.float32)
    keep = pp.non_max_suppression(b, s, 0.5)
    assert keep[0] == s.argmax() and len(set(keep.tolist())) == len(keep)
    assert (np.diff(s[keep]) <= 0).all()                      # descending score order
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    for i, k in enumerate(keep):                              # survivors do not overlap an earlier survivor
        ov = pp.compute_overlap(b[k], b[keep[:i]], area[k], area[keep[:i]]) if i else np.zeros(0)
        assert (ov <= 0.5).all()
    # the overlap is the Dice coefficient of this copy of utils.py: identical boxes -> 1
    assert pp.compute_overlap(b[0], b[:1], area[0], area[:1])[0] == pytest.approx(1.0)
    assert pp.caption_text([5, 6, 2, 7], {5: "a", 6: "dog", 2: ".", 7: "x"}) == "a dog"
    assert len(pp.refine_generations(b, s, 0.99, 10)) == 10


@pytest.mark.gpu
def test_device_refine_generations_matches_oracle():
    import torch
    import image_captioning_b200 as pkg
    rng = np.random.default_rng(1)
    for n, thr, top in ((200, 0.7, 100), (1000, 0.3, 100), (37, 0.5, 5), (1, 0.7, 3)):
        b = _boxes(rng, n)
        s = rng.standard_normal(n).astype(np.float32)
        s[::7] = s[0]                                         # ties: larger index first
        if n > 10:
            b[3] = b[4]                                       # identical boxes (overlap exactly 1)
            b[9] = 0.0                                        # zero-area box: overlap 0/0 = NaN -> never suppressed
            b[10] = 0.0
        want = pp.refine_generations(b, s, thr, top)
        got = pkg.refine_generations(b, s, thr, top)
        assert np.array_equal(got, want), (n, thr)
    # batched, CUDA tensors in -> CUDA tensors out
    bb = np.stack([_boxes(rng, 300) for _ in range(3)])
    ss = rng.standard_normal((3, 300)).astype(np.float32)
    out = pkg.refine_generations(torch.from_numpy(bb).cuda(), torch.from_numpy(ss).cuda(), 0.7, 100)
    for i in range(3):
        assert out[i].is_cuda and np.array_equal(out[i].cpu().numpy(), pp.refine_generations(bb[i], ss[i], 0.7, 100))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_fused_caption_scores(dtype):
    """scores = sum_t log max p without the [N,P,V] dump == the same from the materialised probabilities."""
    import image_captioning_b200 as pkg
    from image_captioning_b200 import synth
    rng = np.random.default_rng(2)
    V, E, U, C, P, B = 2000, 300, 512, 256, 6, 40
    w = synth.synth_weights_v1(rng, V=V, E=E, U=U, C=C)
    cfg = pkg.DenseCapConfig(V, w["imgcap_embedding_layer/embeddings"], 1, P)
    m = pkg.build_lstm_model([7, 7, C], cfg, U, "inference", dtype=dtype)
    m.set_weights(w)
    feat = rng.standard_normal((B, 7, 7, C)).astype(np.float32)
    tok, scores = m.generate(feat, return_scores=True)
    tok2, probs = m.generate(feat, return_probs=True)
    assert np.array_equal(tok, tok2)
    want = pp.caption_scores(probs)
    np.testing.assert_allclose(scores, want, rtol=2e-3, atol=2e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_v2_dense_captioning_flow_scores_and_refine(dtype):
    """The evaluation flow of test_score_dense_captions.py:207-237 for one image on the device: v2 greedy with
    fused caption scores -> refine_generations (NMS + top-100) -> caption text cut at ' .'."""
    import image_captioning_b200 as pkg
    from image_captioning_b200 import synth
    from oracle import decoder as dec
    rng = np.random.default_rng(3)
    V, E, units, C, P, N = 2000, 300, 256, 256, 10, 200
    w = synth.synth_weights_v2(rng, V=V, E=E, units=units, C=C)
    cfg = pkg.DenseCapConfig(V, w["imgcap_embedding_layer/embeddings"], 1, P)
    m = pkg.build_model((7, 7, C), (P,), cfg, units, inject=True, dtype=dtype)
    m.set_weights(w)
    feat = rng.standard_normal((N, 7, 7, C)).astype(np.float32)
    rois = _boxes(rng, N)
    tok, scores = m.generate(feat, return_scores=True)
    tok2, probs = m.generate(feat, return_probs=True)
    assert np.array_equal(tok, tok2)
    np.testing.assert_allclose(scores, pp.caption_scores(probs), rtol=2e-3, atol=2e-3)
    keep = pkg.refine_generations(rois, scores, 0.7, 100)
    assert np.array_equal(keep, pp.refine_generations(rois, scores, 0.7, 100))
    if dtype == "float32":
        tok_want, p_want = dec.greedy_v2(feat[:16], w, P)
        assert np.array_equal(tok[:16], tok_want)
    id_to_word = {i: ("." if i == 2 else "w%d" % i) for i in range(V)}
    text = pkg.caption_text(tok[keep[0]], id_to_word)
    assert " ." not in text and text == pp.caption_text(tok[keep[0]], id_to_word)
