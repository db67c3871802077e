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


# ----------------------------------------------------------------------------------------
# synthetic decoder weights (Keras layouts) and captions
# ----------------------------------------------------------------------------------------

def _glorot(rng, shape, fan_in, fan_out):
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, shape).astype(F32)


def _orthogonal(rng, rows, cols):
    """Keras 'orthogonal' recurrent initializer: QR of a Gaussian [rows, cols]."""
    a = rng.standard_normal((max(rows, cols), max(rows, cols)))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))
    return q[:rows, :cols].astype(F32)


def _lstm_weights(rng, n_in, u):
    kern = _glorot(rng, (n_in, 4 * u), n_in, 4 * u)
    rec = np.concatenate([_orthogonal(rng, u, u) for _ in range(4)], 1)
    bias = np.zeros(4 * u, F32)
    bias[u:2 * u] = 1.0                                   # unit_forget_bias
    return kern, rec, bias


def _head_weights(rng, pool=7, C=256, F=1024):
    w = {}
    w["mrcnn_class_conv1/kernel"] = _glorot(rng, (pool, pool, C, F), pool * pool * C, pool * pool * F)
    # Keras conv fan computation uses receptive field * channels; scale so activations stay O(1)
    w["mrcnn_class_conv1/kernel"] *= F32(np.sqrt((pool * pool * C + pool * pool * F) / (2.0 * pool * pool * C)))
    w["mrcnn_class_conv1/bias"] = (0.1 * rng.standard_normal(F)).astype(F32)
    w["mrcnn_class_conv2/kernel"] = _glorot(rng, (1, 1, F, F), F, F) * F32(1.5)
    w["mrcnn_class_conv2/bias"] = (0.1 * rng.standard_normal(F)).astype(F32)
    for bn in ("mrcnn_class_bn1", "mrcnn_class_bn2"):
        w[bn + "/gamma"] = rng.uniform(0.9, 1.1, F).astype(F32)
        w[bn + "/beta"] = rng.uniform(-0.1, 0.1, F).astype(F32)
        w[bn + "/moving_mean"] = rng.uniform(-0.1, 0.1, F).astype(F32)
        w[bn + "/moving_variance"] = rng.uniform(0.9, 1.1, F).astype(F32)
    return w


def _embedding(rng, V, E):
    e = rng.uniform(-0.5, 0.5, (V, E)).astype(F32)
    e[0] = 0.0                                            # preprocess.py:11 (<unk>/pad row)
    return e


# ---- "trained-like" decoders ---------------------------------------------------------------
# Plain Glorot weights give a near-uniform soft-max whose arg-max is decided by 1e-5-sized gaps and a
# word feedback path too weak to matter; a parity test on such a model says little.  The generators
# below shape the random weights the way a trained captioner's are, so that free-running decoding
#   * never emits <pad>=0 / <start>=1 (strongly negative output bias; a 0 would freeze the masked scan),
#   * uses hundreds of distinct words (Zipf unigram prior, -log(1+rank), over a random permutation of
#     the ids >= 3: rank and id are unrelated, as in a real vocabulary file),
#   * depends on the consumed word (embedding -> gate and state -> dense gains), so tokens change along
#     the caption and a wrong token changes what follows,
#   * is as confident as the reference's own checkpoint: heavy-tailed vocabulary projection scaled to a
#     logit spread of ~3, which puts the mean greedy -log p near the 1.69 validation loss in the
#     reference's checkpoint name (text_generation_model.py:484) and the median top-1/top-2 gap at ~0.6,
#   * favours <end>=2 only late: a few LSTM units are wired as step counters (input/forget/output
#     gates saturated open, constant small candidate) that switch dedicated dense units on after ~10
#     consumed words, which in turn feed the <end> logit.  The reference's loops never stop at <end>
#     (text_generation_model.py:207-229), so what follows it is decoded and compared too.
# tests/test_synth.py asserts these properties so the generator cannot slide back to degenerate captions.

EMBED_GAIN = 8.0          # W[:E] of the LSTM that consumes the word
STATE_GAIN = 5.0          # rows of the dense / image-LSTM kernel fed by the recurrent state
LAYER2_GAIN = 4.0         # input kernel of the second LSTM of the v1 stack
FEATURE_GAIN = 0.5        # rows of the v1 dense kernel fed by the image feature
LOGIT_TAIL = 3.0          # vocabulary projection ~ sign(g) |g|^LOGIT_TAIL, g ~ N(0,1)
LOGIT_SCALE = 3.0         # its column norm, in units of 1/sqrt(fan_in/2)
SPECIAL_BIAS = -30.0      # <pad>, <start>
COUNTER_STEP = 0.06       # tanh(candidate) of a counter unit
V2_STATE_GAIN = 2.0       # v2: rows of the image-LSTM kernel fed by the word vector
V2_LOGIT_SCALE = 14.0     # v2: the vocabulary projection reads 256 bounded LSTM outputs (rms ~0.2), not 1024 ReLUs


def _vocab_kernel(rng, n_in, V, trained_like, scale=LOGIT_SCALE):
    if not trained_like:
        return _glorot(rng, (n_in, V), n_in, V)
    g = rng.standard_normal((n_in, V))
    k = np.sign(g) * np.abs(g) ** LOGIT_TAIL
    k *= scale / np.sqrt(0.5 * n_in) / k.std()
    return k.astype(F32)


def _vocab_bias(rng, V, trained_like):
    if not trained_like:
        return np.zeros(V, F32)
    b = np.full(V, SPECIAL_BIAS, F32)
    if V > 3:
        b[3 + rng.permutation(V - 3)] = -np.log(1.0 + np.arange(V - 3))
    if V > 2:
        b[2] = -np.log(6.0)                               # <end>: a mid-frequency word before the counters fire
    return b


def _n_counters(u):
    return max(1, u // 64)


def _wire_counters(kern, rec, bias, u):
    """Turn the last _n_counters(u) units of a Keras LSTM (blocks i|f|c|o) into step counters:
    c_t = COUNTER_STEP * t, h_t = tanh(c_t), independent of the input."""
    n = _n_counters(u)
    for g, b in ((0, 10.0), (1, 10.0), (2, float(np.arctanh(COUNTER_STEP))), (3, 10.0)):
        cols = slice(g * u + u - n, g * u + u)
        kern[:, cols] = 0.0
        rec[:, cols] = 0.0
        bias[cols] = b
    return n


def synth_weights_v1(rng, V=10000, E=300, F=1024, U=512, pool=7, C=256, trained_like=True):
    """Keras-ordered weights of build_lstm_model; ``trained_like`` as described above, else plain
    Glorot / orthogonal / zero-bias initialisers."""
    w = _head_weights(rng, pool, C, F)
    w["imgcap_embedding_layer/embeddings"] = _embedding(rng, V, E)
    (w["imgcap_lstm1/kernel"], w["imgcap_lstm1/recurrent_kernel"],
     w["imgcap_lstm1/bias"]) = _lstm_weights(rng, E + F, U)
    (w["imgcap_lstm2/kernel"], w["imgcap_lstm2/recurrent_kernel"],
     w["imgcap_lstm2/bias"]) = _lstm_weights(rng, U, U)
    w["imgcap_lstm_d1/kernel"] = _glorot(rng, (U + F, 1024), U + F, 1024)
    w["imgcap_lstm_d1/bias"] = np.zeros(1024, F32)
    w["imgcap_lstm_d2/kernel"] = _vocab_kernel(rng, 1024, V, trained_like)
    w["imgcap_lstm_d2/bias"] = _vocab_bias(rng, V, trained_like)
    if trained_like:
        w["imgcap_lstm1/kernel"][:E] *= F32(EMBED_GAIN)
        w["imgcap_lstm_d1/kernel"][:U] *= F32(STATE_GAIN)
        w["imgcap_lstm_d1/kernel"][U:] *= F32(FEATURE_GAIN)
        w["imgcap_lstm2/kernel"] *= F32(LAYER2_GAIN)
        n = _wire_counters(w["imgcap_lstm2/kernel"], w["imgcap_lstm2/recurrent_kernel"], w["imgcap_lstm2/bias"], U)
        # dense units 1020..1023 = relu(2*8/n * sum(counter h) - 8.5 + noise(f)); they feed only <end>
        k1, k2 = w["imgcap_lstm_d1/kernel"], w["imgcap_lstm_d2/kernel"]
        k1[:U, -4:] = 0.0
        k1[U - n:U, -4:] = 16.0 / n
        w["imgcap_lstm_d1/bias"][-4:] = -8.5
        k2[-4:, :] = 0.0
        if V > 2:
            k2[-4:, 2] = 2.0
    return w


def synth_weights_v2(rng, V=10000, E=300, F=1024, units=256, pool=7, C=256, trained_like=True):
    """Keras-ordered weights of build_model(inject=True) (text_generation_model_v2.py:140-166)."""
    w = _head_weights(rng, pool, C, F)
    w["imgcap_embedding_layer/embeddings"] = _embedding(rng, V, E)
    w["lstm_1/kernel"], w["lstm_1/recurrent_kernel"], w["lstm_1/bias"] = _lstm_weights(rng, E, 1024)
    (w["imgcap_lstm/kernel"], w["imgcap_lstm/recurrent_kernel"],
     w["imgcap_lstm/bias"]) = _lstm_weights(rng, F + 1024, units)
    w["imgcap_d1/kernel"] = _vocab_kernel(rng, units, V, trained_like, V2_LOGIT_SCALE)
    w["imgcap_d1/bias"] = _vocab_bias(rng, V, trained_like)
    if trained_like:
        w["lstm_1/kernel"] *= F32(EMBED_GAIN)
        w["imgcap_lstm/kernel"][F:] *= F32(V2_STATE_GAIN)
        n = _wire_counters(w["lstm_1/kernel"], w["lstm_1/recurrent_kernel"], w["lstm_1/bias"], 1024)
        # image-LSTM units units-2.. (single step from the zero state: h = o*tanh(i*g)) take their
        # candidate from the word-LSTM counters only and feed only <end>
        k, b = w["imgcap_lstm/kernel"], w["imgcap_lstm/bias"]
        for g, bias in ((0, 10.0), (2, -8.5), (3, 10.0)):
            cols = slice(g * units + units - 2, g * units + units)
            k[:, cols] = 0.0
            b[cols] = bias
            if g == 2:
                k[F + 1024 - n:, cols] = 16.0 / n
        w["imgcap_d1/kernel"][-2:, :] = 0.0
        if V > 2:
            w["imgcap_d1/kernel"][-2:, 2] = 12.0
    return w


def synth_captions(rng, B, P, V):
    """[1, w.., 2, 0-pad]; length uniform 3..P including <start>/<end>; ids in [3,V)."""
    gt = np.zeros((B, P), F32)
    for i in range(B):
        L = int(rng.integers(3, P + 1))
        gt[i, 0] = 1
        gt[i, 1:L - 1] = rng.integers(3, V, L - 2)
        gt[i, L - 1] = 2
    return gt
