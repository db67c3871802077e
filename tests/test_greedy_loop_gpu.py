"""The persistent greedy-loop kernel (csrc/greedy_loop.cu: all decoding steps of the v1 word model in one launch, work
items over (step, stage, row block) with dependency counters) against the launch-per-GEMM path it replaces
(DCAP_GREEDY_LOOP=0, itself held to the oracle in test_decoder_gpu.py): the same tcgen05 contractions in the same
order, so token ids must be IDENTICAL and caption scores equal to fp32 rounding -- on ragged batch sizes (partial
128-row blocks, a single row block, fewer row blocks than the wavefront skew) and at the BASELINE size."""
import os

import numpy as np
import pytest
import torch

from image_captioning_b200 import synth

pytestmark = pytest.mark.gpu

V, E, U, C, P = 10000, 300, 512, 256, 15


def _model(w):
    import image_captioning_b200 as pkg
    cfg = pkg.DenseCapConfig(V, w["imgcap_embedding_layer/embeddings"], 1, P)
    m = pkg.build_lstm_model([7, 7, C], cfg, U, "inference", dtype="bfloat16")
    m.set_weights(w)
    return m


def _run(w, feats, loop):
    old = os.environ.get("DCAP_GREEDY_LOOP")
    os.environ["DCAP_GREEDY_LOOP"] = "2" if loop else "0"          # 2 = the loop kernel at every batch size
    try:
        m = _model(w)
        calls = [m.generate(feats).cpu().numpy() for _ in range(3)]          # eager, graph capture, graph replay
        tok_s, sc = m.generate(feats, return_scores=True)
        torch.cuda.synchronize()
        return calls, tok_s.cpu().numpy(), sc.cpu().numpy()
    finally:
        if old is None:
            os.environ.pop("DCAP_GREEDY_LOOP", None)
        else:
            os.environ["DCAP_GREEDY_LOOP"] = old


@pytest.mark.parametrize("B", [1, 37, 300, 1000, 2500])
def test_loop_kernel_equals_launch_per_gemm_path(B):
    w = synth.synth_weights_v1(np.random.default_rng(1005), V=V, E=E, U=U, C=C)
    feats = torch.randn((B, 1024), device="cuda", generator=torch.Generator(device="cuda").manual_seed(5 + B)).relu()
    want_calls, want_tok, want_sc = _run(w, feats, loop=False)
    got_calls, got_tok, got_sc = _run(w, feats, loop=True)
    for c in got_calls:
        assert np.array_equal(c, want_calls[0])
    assert np.array_equal(got_tok, want_tok)
    np.testing.assert_allclose(got_sc, want_sc, rtol=0, atol=2e-5)
    if B >= 300:
        assert len(np.unique(got_calls[0])) >= 150, "degenerate captions"


def test_loop_kernel_full_size_and_back_to_back_calls():
    """8000 RoIs (the bench shape): identical ids; then alternating batch sizes on one handle (the counters, the
    flagged partials and the blocked cell state are re-initialised by every call)."""
    w = synth.synth_weights_v1(np.random.default_rng(1005), V=V, E=E, U=U, C=C)
    feats = torch.randn((8000, 1024), device="cuda", generator=torch.Generator(device="cuda").manual_seed(5)).relu()
    want, _, _ = _run(w, feats, loop=False)
    got, _, _ = _run(w, feats, loop=True)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[2], want[0])
    m = _model(w)
    a = m.generate(feats).cpu().numpy()
    b = m.generate(feats[:333].contiguous()).cpu().numpy()
    c = m.generate(feats).cpu().numpy()
    d = m.generate(feats[:333].contiguous()).cpu().numpy()
    assert np.array_equal(a, want[0]) and np.array_equal(c, a)
    assert np.array_equal(b, a[:333]) and np.array_equal(d, b)


@pytest.mark.parametrize("B", [16, 700])
def test_loop_kernel_from_roi_features(B):
    """RoI features in (the head runs inside the call, its bf16 output feeds the merged hoist GEMM directly): same ids
    as the launch-per-GEMM path, and close to the fp32 oracle."""
    from oracle import decoder as dec
    rng = np.random.default_rng(77)
    w = synth.synth_weights_v1(rng, V=V, E=E, U=U, C=C)
    roi = torch.from_numpy(rng.standard_normal((B, 7, 7, C)).astype(np.float32)).cuda()
    want_calls, want_tok, want_sc = _run(w, roi, loop=False)
    got_calls, got_tok, got_sc = _run(w, roi, loop=True)
    for c in got_calls:
        assert np.array_equal(c, want_calls[0])
    assert np.array_equal(got_tok, want_tok)
    np.testing.assert_allclose(got_sc, want_sc, rtol=0, atol=2e-5)
    if B <= 64:
        tok32, _ = dec.greedy_v1(dec.head(roi.cpu().numpy(), w), w, P)
        assert (got_calls[0] == tok32).mean() >= 0.9
