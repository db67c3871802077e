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


def _vocab_bias(V, trained_like):
    if not trained_like:
        return np.zeros(V, F32)
    # Zipfian unigram prior, as a trained captioner's output bias has (log frequency).
    return (-np.log(1.0 + np.arange(V))).astype(F32)


def synth_weights_v1(rng, V=10000, E=300, F=1024, U=512, pool=7, C=256, trained_like=True):
    """Keras-ordered weights of build_lstm_model.  ``trained_like`` sharpens the output
    distribution (logit temperature + Zipf bias) the way a trained captioner's is; plain
    Glorot gives a near-uniform softmax whose argmax is decided by 1e-3-sized gaps."""
    w = _head_weights(rng, pool, C, F)
    w["imgcap_embedding_layer/embeddings"] = _embedding(rng, V, E)
    (w["imgcap_lstm1/kernel"], w["imgcap_lstm1/recurrent_kernel"],
     w["imgcap_lstm1/bias"]) = _lstm_weights(rng, E + F, U)
    (w["imgcap_lstm2/kernel"], w["imgcap_lstm2/recurrent_kernel"],
     w["imgcap_lstm2/bias"]) = _lstm_weights(rng, U, U)
    w["imgcap_lstm_d1/kernel"] = _glorot(rng, (U + F, 1024), U + F, 1024)
    w["imgcap_lstm_d1/bias"] = np.zeros(1024, F32)
    w["imgcap_lstm_d2/kernel"] = _glorot(rng, (1024, V), 1024, V)
    if trained_like:
        w["imgcap_lstm_d2/kernel"] *= F32(3.0)
    w["imgcap_lstm_d2/bias"] = _vocab_bias(V, trained_like)
    return w


def synth_weights_v2(rng, V=10000, E=300, F=1024, units=256, pool=7, C=256, trained_like=True):
    """Keras-ordered weights of build_model(inject=True) (text_generation_model_v2.py:140-166)."""
    w = _head_weights(rng, pool, C, F)
    w["imgcap_embedding_layer/embeddings"] = _embedding(rng, V, E)
    w["lstm_1/kernel"], w["lstm_1/recurrent_kernel"], w["lstm_1/bias"] = _lstm_weights(rng, E, 1024)
    (w["imgcap_lstm/kernel"], w["imgcap_lstm/recurrent_kernel"],
     w["imgcap_lstm/bias"]) = _lstm_weights(rng, F + 1024, units)
    w["imgcap_d1/kernel"] = _glorot(rng, (units, V), units, V)
    if trained_like:
        w["imgcap_d1/kernel"] *= F32(3.0)
    w["imgcap_d1/bias"] = _vocab_bias(V, trained_like)
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
