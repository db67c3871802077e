"""Oracle self-consistency for the decoder (CPU only): the literal O(P^2) forms the reference
executes equal the incremental scans the CUDA path implements; the fp64 BPTT gradients equal
finite differences; loss / AMSGrad restatements behave as Keras documents."""
import numpy as np
import pytest

from image_captioning_b200 import synth
from oracle import decoder as dec

SMALL = dict(V=50, E=12, F=1024, U=16, pool=2, C=4)


def _small(seed=0, trained_like=True):
    rng = np.random.default_rng(seed)
    w = synth.synth_weights_v1(rng, trained_like=trained_like, **SMALL)
    feat = rng.standard_normal((5, SMALL["pool"], SMALL["pool"], SMALL["C"])).astype(np.float32)
    return rng, w, feat


def test_greedy_literal_equals_incremental():
    rng, w, feat = _small(1)
    f = dec.head(feat, w)
    P = 6
    lit = dec.greedy_v1_literal(f, w, P)
    tok, inc = dec.greedy_v1(f, w, P)
    assert np.array_equal(lit.argmax(-1), tok)
    np.testing.assert_allclose(lit, inc, rtol=1e-5, atol=1e-7)


def test_token_zero_mid_sequence_carries_state():
    """A generated/ground-truth id 0 in the middle is masked: state carried, output recomputed."""
    rng, w, feat = _small(2)
    f = dec.head(feat, w)
    gt = np.array([[1, 7, 0, 9, 2, 0], [1, 0, 0, 3, 2, 0], [1, 4, 5, 6, 7, 2], [1, 2, 0, 0, 0, 0],
                   [1, 9, 9, 0, 9, 2]], np.float32)
    lit = dec.train_forward_v1_literal(f, gt, w)
    inc = dec.train_forward_v1(f, gt, w)
    np.testing.assert_allclose(lit, inc, rtol=1e-5, atol=1e-7)
    # masked step re-emits the distribution of the previous step
    np.testing.assert_allclose(inc[0, 2], inc[0, 1], rtol=1e-6)


def test_fp64_shadow_close_to_fp32():
    rng, w, feat = _small(3)
    a = dec.greedy_v1(dec.head(feat, w), w, 5)[1]
    b = dec.greedy_v1(dec.head(feat, w, np.float64), w, 5, np.float64)[1]
    np.testing.assert_allclose(a, b, rtol=2e-3, atol=1e-6)


def test_loss_and_targets():
    gt = np.array([[1, 5, 2, 0]], np.float32)
    assert dec.targets_from_captions(gt).tolist() == [[5, 2, 0, 0]]
    p = np.full((1, 4, 6), 0.1, np.float32)
    p[0, :, 5] = 0.5
    loss = dec.roi_caption_loss(dec.targets_from_captions(gt), p)
    want = -(np.log(0.5) + 3 * np.log(0.1)) / 4
    assert abs(loss - want) < 1e-6
    # clipping at 1e-7
    p2 = np.zeros((1, 1, 3), np.float32); p2[0, 0, 0] = 1.0
    assert abs(dec.roi_caption_loss(np.array([[1]]), p2) + np.log(1e-7)) < 1e-3


def test_keras_amsgrad_first_steps():
    p = np.array([1.0, -2.0], np.float32)
    g = np.array([0.5, -0.25], np.float32)
    m = v = vh = np.zeros(2, np.float32)
    p1, m, v, vh = dec.keras_adam_amsgrad(p, g, m, v, vh, 1)
    # t=1: lr_t = lr*sqrt(1-b2)/(1-b1); m=(1-b1)g; v=(1-b2)g^2 -> step ~ lr*sign(g)
    np.testing.assert_allclose(p - p1, 1e-3 * np.sign(g), rtol=1e-3)
    p2, m, v, vh2 = dec.keras_adam_amsgrad(p1, g * 0.1, m, v, vh, 2)
    assert (vh2 >= v).all() and (vh2 >= vh).all()


def test_bptt_gradients_match_finite_differences():
    rng, w, feat = _small(4, trained_like=False)
    w = {k: v.astype(np.float64) for k, v in w.items()}
    gt = synth.synth_captions(rng, 5, 5, SMALL["V"])
    gt[1, 2] = 0                                   # masked step in the middle
    loss, G = dec.train_loss_and_grads_v1(feat, gt, w)
    probs = dec.train_forward_v1(dec.head(feat, w, np.float64), gt, w, np.float64)
    ref = dec.roi_caption_loss(dec.targets_from_captions(gt), probs)
    assert abs(loss - ref) < 1e-9
    checks = [("imgcap_lstm_d2/kernel", (3, 7)), ("imgcap_lstm_d1/kernel", (20, 5)), ("imgcap_lstm2/kernel", (3, 40)),
              ("imgcap_lstm2/recurrent_kernel", (2, 9)), ("imgcap_lstm1/kernel", (5, 20)),
              ("imgcap_lstm1/kernel", (100, 33)), ("imgcap_lstm1/recurrent_kernel", (1, 50)),
              ("imgcap_lstm1/bias", (17,)), ("mrcnn_class_conv2/kernel", (0, 0, 3, 4)),
              ("mrcnn_class_conv1/kernel", (1, 0, 2, 7)), ("mrcnn_class_conv1/bias", (3,)),
              ("mrcnn_class_bn1/gamma", (11,)), ("mrcnn_class_bn1/beta", (12,)), ("mrcnn_class_bn2/gamma", (5,)),
              ("mrcnn_class_bn2/beta", (6,)), ("imgcap_lstm_d2/bias", (9,)), ("imgcap_lstm_d1/bias", (100,))]
    eps = 1e-5
    for name, idx in checks:
        w2 = dict(w); a = w[name].copy(); a[idx] += eps; w2[name] = a
        lp = dec.roi_caption_loss(dec.targets_from_captions(gt),
                                  dec.train_forward_v1(dec.head(feat, w2, np.float64), gt, w2, np.float64))
        a = w[name].copy(); a[idx] -= eps; w2[name] = a
        lm = dec.roi_caption_loss(dec.targets_from_captions(gt),
                                  dec.train_forward_v1(dec.head(feat, w2, np.float64), gt, w2, np.float64))
        fd = (lp - lm) / (2 * eps)
        assert abs(fd - G[name][idx]) < 1e-6 + 1e-4 * abs(fd), (name, idx, fd, G[name][idx])


def test_v2_greedy_incremental_property():
    """v2: the pre-padded window never truncates during the reference loop, so the word LSTM
    state can be carried; check predict() on the full prefix equals a carried-state run."""
    rng = np.random.default_rng(5)
    w = synth.synth_weights_v2(rng, V=40, E=10, F=1024, units=8, pool=2, C=4)
    feat = rng.standard_normal((3, 2, 2, 4)).astype(np.float32)
    tok, probs = dec.greedy_v2(feat, w, P=6)
    assert tok.shape == (3, 5)
    # re-predict the last step from the explicit padded prefix
    seqs = [[0] + tok[i, :4].tolist() for i in range(3)]
    p_last = dec.v2_inject_predict(feat, dec.pad_sequences_pre(seqs, 6), w)
    np.testing.assert_allclose(p_last, probs[:, 4], rtol=1e-6)
    # all-masked prefix -> word vector 0
    p0 = dec.v2_inject_predict(feat, np.zeros((3, 6), np.int32), w)
    np.testing.assert_allclose(p0, probs[:, 0], rtol=1e-6)


def test_beam_width1_equals_greedy_and_scores_accumulate():
    rng, w, feat = _small(6)
    f = dec.head(feat, w)
    P = 5
    tok, probs = dec.greedy_v1(f, w, P)
    bt, bs = dec.beam_v1(f, w, P, 1)
    assert np.array_equal(bt[:, 0, 0], np.ones(5, np.int32))
    assert np.array_equal(bt[:, 0, 1:], tok[:, :P - 1])
    np.testing.assert_allclose(bs[:, 0], probs[:, :P - 1].max(-1).astype(np.float64).sum(1), rtol=1e-5)
    bt3, bs3 = dec.beam_v1(f, w, P, 3)
    assert (np.diff(bs3, axis=1) >= 0).all()            # ascending, best last
    assert (bs3[:, -1] >= bs[:, 0] - 1e-9).all()


def test_oracle_reproduces_golden_decoder_vectors(golden_dir):
    """tests/golden/decoder_small.npz freezes the oracle (parity unpinned by the reference)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("gen_golden_decoder", os.path.join(golden_dir, "gen_golden_decoder.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    now = mod.build()
    g = np.load(os.path.join(golden_dir, "decoder_small.npz"))
    for k in ("greedy_tok", "beam_tok", "v2_tok", "gt"):
        assert np.array_equal(now[k], g[k]), k
    for k in ("head", "greedy_probs", "teacher_forced_probs", "beam_scores", "adam_p1", "adam_m1", "adam_v1", "v2_probs_last"):
        np.testing.assert_allclose(now[k], g[k], rtol=1e-5, atol=1e-8, err_msg=k)
    assert abs(float(now["loss"]) - float(g["loss"])) < 1e-6
    np.testing.assert_allclose(now["grad_l2"], g["grad_l2"], rtol=1e-9)
    np.testing.assert_allclose(now["grad_sum"], g["grad_sum"], rtol=1e-7, atol=1e-12)


def test_amsgrad_mechanics_agree_with_an_independent_implementation():
    """torch.optim.Adam(amsgrad=True) is the same recurrence as Keras' up to where epsilon enters (Keras: sqrt(vhat) + eps
    after folding both bias corrections into lr_t; torch: sqrt(vhat / (1 - b2^t)) + eps).  With a negligible epsilon
    the two must coincide: an independent check of the moment updates, the running maximum and the bias corrections."""
    import torch
    rng = np.random.default_rng(17)
    p0 = rng.standard_normal((5, 7))
    grads = [rng.standard_normal((5, 7)) * s for s in (1.0, 0.1, 3.0, 0.01, 1.0, 0.5)]     # vhat must hold earlier maxima
    tp = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    opt = torch.optim.Adam([tp], lr=1e-3, betas=(0.9, 0.999), eps=1e-30, amsgrad=True)
    p, m, v, vh = p0.copy(), np.zeros_like(p0), np.zeros_like(p0), np.zeros_like(p0)
    for t, g in enumerate(grads, 1):
        tp.grad = torch.from_numpy(g.copy())
        opt.step()
        p, m, v, vh = dec.keras_adam_amsgrad(p, g, m, v, vh, t, eps=1e-30)
        np.testing.assert_allclose(p, tp.detach().numpy(), rtol=1e-12, atol=1e-15)
    assert (vh >= v).all() and (vh > v).any()


def test_lstm_cell_wiring_agrees_with_an_independent_implementation():
    """torch.nn.LSTMCell has the same gate order (i, f, g, o) and state update as Keras' LSTMCell and differs only in the
    recurrent activation (logistic sigmoid vs hard_sigmoid): with that swapped in, the oracle's cell must reproduce it.
    And Keras' hard_sigmoid is the line 0.2 x + 0.5 clipped to [0, 1] (tangent-free piecewise linear, exact at 0, +-2.5)."""
    import torch
    rng = np.random.default_rng(18)
    n_in, u, B = 5, 4, 3
    cell = torch.nn.LSTMCell(n_in, u).double()
    x, h, c = (rng.standard_normal(s) for s in ((B, n_in), (B, u), (B, u)))
    with torch.no_grad():
        h_t, c_t = cell(torch.from_numpy(x), (torch.from_numpy(h), torch.from_numpy(c)))
    kernel = cell.weight_ih.detach().numpy().T                  # Keras stores [in, 4u]
    recurrent = cell.weight_hh.detach().numpy().T
    bias = (cell.bias_ih + cell.bias_hh).detach().numpy()
    sigmoid = lambda z: 1.0 / (1.0 + np.exp(-z))
    h_o, c_o = dec.lstm_cell(x, h, c, kernel, recurrent, bias, recurrent_activation=sigmoid)
    np.testing.assert_allclose(h_o, h_t.numpy(), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(c_o, c_t.numpy(), rtol=1e-12, atol=1e-14)
    z = np.array([-3.0, -2.5, -1.0, 0.0, 1.0, 2.5, 3.0])
    assert dec.hard_sigmoid(z).tolist() == [0.0, 0.0, 0.3, 0.5, 0.7, 1.0, 1.0]


def test_batchnorm_and_cross_entropy_agree_with_independent_implementations():
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(19)
    x = rng.standard_normal((6, 9))
    gamma, beta, mean = (rng.standard_normal(9) for _ in range(3))
    var = rng.uniform(0.1, 2.0, 9)
    want = F.batch_norm(torch.from_numpy(x), torch.from_numpy(mean), torch.from_numpy(var), torch.from_numpy(gamma),
                        torch.from_numpy(beta), training=False, eps=dec.BN_EPS).numpy()
    np.testing.assert_allclose(dec.batchnorm_inference(x, gamma, beta, mean, var), want, rtol=1e-12, atol=1e-13)
    logits = rng.standard_normal((4, 5, 11)) * 3
    ids = rng.integers(0, 11, (4, 5))
    probs = dec.softmax(logits)
    want_loss = F.cross_entropy(torch.from_numpy(logits).reshape(-1, 11), torch.from_numpy(ids).reshape(-1)).item()
    np.testing.assert_allclose(dec.roi_caption_loss(ids, probs), want_loss, rtol=1e-10)      # no probability near the 1e-7 clip
    valid = rng.uniform(size=(4, 5)) < 0.6
    want_masked = F.cross_entropy(torch.from_numpy(logits)[torch.from_numpy(valid)], torch.from_numpy(ids)[torch.from_numpy(valid)]).item()
    np.testing.assert_allclose(dec.roi_caption_loss(ids, probs, valid), want_masked, rtol=1e-10)


def test_v2_training_gradients_match_finite_differences():
    """oracle.train_loss_and_grads_v2 (the checker of dc_decoder_v2_train_step): analytic fp64 BPTT through the masked,
    pre-padded word LSTM, the single image-LSTM step and Dense(V) against central differences of the oracle's own
    forward (v2_inject_predict); the image LSTM's recurrent kernel sees only the zero state -> zero gradient."""
    rng = np.random.default_rng(3)
    V, E, units, C, L, B = 30, 6, 5, 4, 5, 7
    w = synth.synth_weights_v2(rng, V=V, E=E, F=1024, units=units, pool=2, C=C, trained_like=False)
    w = {k: v.astype(np.float64) for k, v in w.items()}
    feat = rng.standard_normal((B, 2, 2, C))
    words = np.zeros((B, L), np.int32)
    for i in range(B):
        n = rng.integers(0, L + 1)
        words[i, L - n:] = rng.integers(1, V, n)
    words[2, L - 2] = 0
    y = rng.integers(0, V, B)
    loss, G = dec.train_loss_and_grads_v2(feat, words, y, w)

    def f(wd):
        p = dec.v2_inject_predict(feat, words, wd, np.float64)
        return -np.log(p[np.arange(B), y]).mean()

    assert abs(loss - f(w)) < 1e-12
    assert not G["imgcap_lstm/recurrent_kernel"].any()
    for name, g in G.items():
        for _ in range(5):
            ix = tuple(rng.integers(0, s) for s in g.shape)
            wp, wm = dict(w), dict(w)
            a = w[name].copy(); a[ix] += 1e-6; wp[name] = a
            b = w[name].copy(); b[ix] -= 1e-6; wm[name] = b
            fd = (f(wp) - f(wm)) / 2e-6
            assert abs(fd - g[ix]) <= 1e-6 + 1e-4 * abs(fd), (name, ix, fd, g[ix])


def test_philox_known_answers_and_mask_statistics():
    """Random123 known-answer vectors for Philox-4x32-10 (kat_vectors: zero / all-ones / pi-digits inputs), then the
    recurrent-dropout masks drawn from it: values {0, 1/(1-rate)}, keep rate ~ 1-rate, four different masks per LSTM,
    independent of how the batch is cut into shards (row_offset)."""
    z = np.zeros((1, 4), np.uint32)
    assert [hex(v) for v in dec.philox4x32_10(z, (0, 0))[0]] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    f = np.full((1, 4), 0xFFFFFFFF, np.uint32)
    assert [hex(v) for v in dec.philox4x32_10(f, (0xFFFFFFFF, 0xFFFFFFFF))[0]] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    pi = np.array([[0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344]], np.uint32)
    assert [hex(v) for v in dec.philox4x32_10(pi, (0xA4093822, 0x299F31D0))[0]] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]
    m = dec.philox_masks(0.2, seed=1234, step=7, layer=1, rows=512, units=64)
    assert m.shape == (512, 4, 64) and set(np.unique(m)) == {0.0, 1.25}
    assert abs((m > 0).mean() - 0.8) < 0.01
    assert not np.array_equal(m[:, 0], m[:, 1]) and not np.array_equal(m, dec.philox_masks(0.2, 1234, 8, 1, 512, 64))
    assert not np.array_equal(m, dec.philox_masks(0.2, 1234, 7, 2, 512, 64))
    assert np.array_equal(m[100:], dec.philox_masks(0.2, 1234, 7, 1, 412, 64, row_offset=100))


def test_dropout_gradients_match_finite_differences():
    """fp64 BPTT with recurrent-dropout masks (per gate, time-invariant) against central differences."""
    rng = np.random.default_rng(8)
    shp = dict(V=20, E=5, F=1024, U=6, pool=2, C=3)
    w = {k: v.astype(np.float64) for k, v in synth.synth_weights_v1(rng, trained_like=False, **shp).items()}
    B, P = 5, 4
    feat = rng.standard_normal((B, 2, 2, 3))
    gt = synth.synth_captions(rng, B, P, shp["V"]); gt[1, 2] = 0
    masks = (dec.philox_masks(0.3, 5, 1, 1, B, 6), dec.philox_masks(0.3, 5, 1, 2, B, 6))
    loss, G = dec.train_loss_and_grads_v1(feat, gt, w, rec_masks=masks)
    loss0, _ = dec.train_loss_and_grads_v1(feat, gt, w)
    assert abs(loss - loss0) > 1e-6
    for name in ("imgcap_lstm1/recurrent_kernel", "imgcap_lstm2/recurrent_kernel", "imgcap_lstm1/kernel", "imgcap_lstm_d1/kernel",
                 "mrcnn_class_conv2/kernel"):
        for _ in range(4):
            ix = tuple(rng.integers(0, s) for s in G[name].shape)
            wp, wm = dict(w), dict(w)
            a = w[name].copy(); a[ix] += 1e-6; wp[name] = a
            b = w[name].copy(); b[ix] -= 1e-6; wm[name] = b
            fd = (dec.train_loss_and_grads_v1(feat, gt, wp, rec_masks=masks)[0] - dec.train_loss_and_grads_v1(feat, gt, wm, rec_masks=masks)[0]) / 2e-6
            assert abs(fd - G[name][ix]) <= 1e-7 + 1e-4 * abs(fd), (name, ix, fd, G[name][ix])
