"""CPU oracle for the RoI head + "inject" LSTM caption decoder (TEST INFRASTRUCTURE ONLY).

numpy fp32 restatement (optionally fp64 "shadow" by passing ``dtype=np.float64``) of

  * head                      dense_img_cap_separate_models/text_generation_model.py:249-262
                              (v2: text_generation_model_v2.py:141-150); frozen BatchNorm
                              evaluate_models/modified_dense_model.py:52-62
  * word_generation_model     text_generation_model.py:130-156
  * ROICaptionInferenceLayer  text_generation_model.py:192-232      (v1 greedy)
  * training graph            text_generation_model.py:159-189, 264-277
  * roi_caption_loss          text_generation_model.py:286-294
  * Adam(amsgrad=True)        text_generation_model.py:425
  * v2 build_model(inject)    text_generation_model_v2.py:140-166
  * v2 greedy loop            evaluate_models/test_score_dense_captions.py:216-225
  * beam search gen_captions  image captioning/test.py:23-64

(all paths relative to /root/reference).  Keras-2.1 / TF-1.x internals (LSTMCell with
implementation=1 and hard_sigmoid, K.rnn mask handling, BatchNormalization inference,
categorical_crossentropy, Adam) are restated from their published behaviour because their
source is not under /root/reference.  PARITY UNPINNED for the network arithmetic: the reference
holds no tests, golden vectors, weights or vocabularies for this path and cannot be imported here
(no TF/Keras).  PINNED (tests/golden/gen_golden_reference_numpy.py -> tests/test_reference_golden.py):
  * the two decoding loops that are plain Python around model.predict -- beam search gen_captions and
    the v2 greedy loop -- run from /root/reference with a stand-in predict(): beam_v1 / greedy_v2
    reproduce their outputs exactly;
  * the model WIRING: word_generation_model, ROICaptionInferenceLayer.call,
    build_roi_caption_model_training, roi_caption_loss, build_lstm_model (both modes) and the v2
    build_model are executed from the reference source over eager stand-ins for the Keras layers
    (tests/golden/tf_numpy_shim.py; layer numerics = the primitives below): greedy_v1_literal,
    train_forward_v1_literal, roi_caption_loss, head and v2_inject_predict reproduce them.
What remains unpinned are the Keras/TF primitives themselves (LSTM cell, Dense, BatchNorm, crossentropy, Adam).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import numpy as np

F32 = np.float32
BN_EPS = 1e-3            # keras.layers.BatchNormalization default epsilon


# ----------------------------------------------------------------------------------------
# primitives
# ----------------------------------------------------------------------------------------

def hard_sigmoid(x):
    """Keras TF backend: clip(0.2*x + 0.5, 0, 1)."""
    dt = x.dtype.type
    return np.clip(dt(0.2) * x + dt(0.5), dt(0), dt(1))


def batchnorm_inference(x, gamma, beta, mean, var, eps=BN_EPS):
    """tf.nn.batch_normalization: inv = rsqrt(var+eps)*gamma; x*inv + (beta - mean*inv)."""
    dt = x.dtype.type
    inv = (dt(1) / np.sqrt(var.astype(x.dtype) + dt(eps))) * gamma.astype(x.dtype)
    return x * inv + (beta.astype(x.dtype) - mean.astype(x.dtype) * inv)


def softmax(z):
    z = z - z.max(axis=-1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=-1, keepdims=True)


def lstm_cell(x, h, c, kernel, recurrent, bias, recurrent_activation=hard_sigmoid):
    """Keras LSTMCell (implementation=1, tanh / hard_sigmoid, gate blocks i|f|c|o).  ``recurrent_activation`` exists
    for the cross-check against torch.nn.LSTMCell (logistic sigmoid) in tests/test_oracle_decoder.py."""
    z = (x @ kernel + bias) + h @ recurrent
    u = h.shape[-1]
    i = recurrent_activation(z[..., 0 * u:1 * u])
    f = recurrent_activation(z[..., 1 * u:2 * u])
    g = np.tanh(z[..., 2 * u:3 * u])
    o = recurrent_activation(z[..., 3 * u:4 * u])
    c_new = f * c + i * g
    h_new = o * np.tanh(c_new)
    return h_new, c_new


def lstm_masked(xs, mask, kernel, recurrent, bias, return_sequences):
    """K.rnn with a mask: masked step emits the previous h and carries (h, c)."""
    B, T, _ = xs.shape
    u = recurrent.shape[0]
    h = np.zeros((B, u), xs.dtype)
    c = np.zeros((B, u), xs.dtype)
    outs = []
    for t in range(T):
        hn, cn = lstm_cell(xs[:, t], h, c, kernel, recurrent, bias)
        m = mask[:, t][:, None]
        h = np.where(m, hn, h)
        c = np.where(m, cn, c)
        outs.append(h)
    return np.stack(outs, 1) if return_sequences else h


# ----------------------------------------------------------------------------------------
# head (a7)
# ----------------------------------------------------------------------------------------

def head(feat, w, dtype=F32):
    """relu(bn2(relu(bn1(conv7x7_valid(X)))*W2+b2)); feat [B,7,7,256] -> [B,1024]."""
    x = feat.reshape(feat.shape[0], -1).astype(dtype)
    k1 = w["mrcnn_class_conv1/kernel"].reshape(-1, w["mrcnn_class_conv1/kernel"].shape[-1])
    a = x @ k1.astype(dtype) + w["mrcnn_class_conv1/bias"].astype(dtype)
    a = batchnorm_inference(a, w["mrcnn_class_bn1/gamma"], w["mrcnn_class_bn1/beta"],
                            w["mrcnn_class_bn1/moving_mean"], w["mrcnn_class_bn1/moving_variance"])
    a = np.maximum(a, 0)
    k2 = w["mrcnn_class_conv2/kernel"].reshape(-1, w["mrcnn_class_conv2/kernel"].shape[-1])
    a = a @ k2.astype(dtype) + w["mrcnn_class_conv2/bias"].astype(dtype)
    a = batchnorm_inference(a, w["mrcnn_class_bn2/gamma"], w["mrcnn_class_bn2/beta"],
                            w["mrcnn_class_bn2/moving_mean"], w["mrcnn_class_bn2/moving_variance"])
    return np.maximum(a, 0)


# ----------------------------------------------------------------------------------------
# v1 word model (a8), greedy (a9), training (a10)
# ----------------------------------------------------------------------------------------

def word_model_literal(f, prev_words, w, dtype=F32, return_logits=False):
    """word_generation_model on rows [feat ; prev_words] (text_generation_model.py:130-156).
    f [B,F]; prev_words [B,P] float token ids (0 = masked)."""
    f = f.astype(dtype)
    ids = prev_words.astype(np.int32)
    mask = ids != 0
    emb = w["imgcap_embedding_layer/embeddings"].astype(dtype)[ids]           # [B,P,E]
    P = ids.shape[1]
    ctx = np.concatenate([emb, np.repeat(f[:, None, :], P, 1)], -1)           # [emb ; feat]
    s1 = lstm_masked(ctx, mask, w["imgcap_lstm1/kernel"].astype(dtype),
                     w["imgcap_lstm1/recurrent_kernel"].astype(dtype),
                     w["imgcap_lstm1/bias"].astype(dtype), True)
    h2 = lstm_masked(s1, mask, w["imgcap_lstm2/kernel"].astype(dtype),
                     w["imgcap_lstm2/recurrent_kernel"].astype(dtype),
                     w["imgcap_lstm2/bias"].astype(dtype), False)
    d = np.maximum(np.concatenate([h2, f], -1) @ w["imgcap_lstm_d1/kernel"].astype(dtype)
                   + w["imgcap_lstm_d1/bias"].astype(dtype), 0)
    z = d @ w["imgcap_lstm_d2/kernel"].astype(dtype) + w["imgcap_lstm_d2/bias"].astype(dtype)
    return z if return_logits else softmax(z)


class V1State:
    def __init__(self, B, U, dtype=F32):
        self.h1 = np.zeros((B, U), dtype); self.c1 = np.zeros((B, U), dtype)
        self.h2 = np.zeros((B, U), dtype); self.c2 = np.zeros((B, U), dtype)

    def gather(self, idx):
        s = V1State(0, 0)
        s.h1, s.c1, s.h2, s.c2 = self.h1[idx], self.c1[idx], self.h2[idx], self.c2[idx]
        return s


def v1_step(f, tok, st, w, dtype=F32, return_logits=False):
    """Incremental form: consume token ``tok`` [B] with carried state."""
    f = f.astype(dtype)
    tok = np.asarray(tok).astype(np.int32)
    m = (tok != 0)[:, None]
    x1 = np.concatenate([w["imgcap_embedding_layer/embeddings"].astype(dtype)[tok], f], -1)
    h1n, c1n = lstm_cell(x1, st.h1, st.c1, w["imgcap_lstm1/kernel"].astype(dtype),
                         w["imgcap_lstm1/recurrent_kernel"].astype(dtype),
                         w["imgcap_lstm1/bias"].astype(dtype))
    st.h1 = np.where(m, h1n, st.h1); st.c1 = np.where(m, c1n, st.c1)
    h2n, c2n = lstm_cell(st.h1, st.h2, st.c2, w["imgcap_lstm2/kernel"].astype(dtype),
                         w["imgcap_lstm2/recurrent_kernel"].astype(dtype),
                         w["imgcap_lstm2/bias"].astype(dtype))
    st.h2 = np.where(m, h2n, st.h2); st.c2 = np.where(m, c2n, st.c2)
    d = np.maximum(np.concatenate([st.h2, f], -1) @ w["imgcap_lstm_d1/kernel"].astype(dtype)
                   + w["imgcap_lstm_d1/bias"].astype(dtype), 0)
    z = d @ w["imgcap_lstm_d2/kernel"].astype(dtype) + w["imgcap_lstm_d2/bias"].astype(dtype)
    return z if return_logits else softmax(z)


def greedy_v1_literal(f, w, P, dtype=F32):
    """ROICaptionInferenceLayer.call: P word-model runs over the growing post-padded prefix
    (O(P^2) LSTM steps), start id 1, argmax fed back as float, never stops at <end>.
    Returns probs [B,P,V]."""
    B = f.shape[0]
    prev = np.ones((B, 1), F32)
    outs = []
    for j in range(P):
        ctx = np.concatenate([prev, np.zeros((B, P - j - 1), F32)], -1)
        p = word_model_literal(f, ctx, w, dtype)
        outs.append(p)
        prev = np.concatenate([prev, p.argmax(-1).astype(F32)[:, None]], -1)
    return np.stack(outs, 1)


def greedy_v1(f, w, P, dtype=F32, return_logits=False):
    """Incremental masked scan; equals greedy_v1_literal.  Returns (tokens [B,P] int32,
    probs-or-logits [B,P,V])."""
    B = f.shape[0]
    U = w["imgcap_lstm1/recurrent_kernel"].shape[0]
    st = V1State(B, U, dtype)
    tok = np.ones(B, np.int32)
    toks, outs = [], []
    for _ in range(P):
        p = v1_step(f, tok, st, w, dtype, return_logits)
        tok = p.argmax(-1).astype(np.int32)
        toks.append(tok); outs.append(p)
    return np.stack(toks, 1), np.stack(outs, 1)


def train_forward_v1_literal(f, gt, w, dtype=F32):
    """build_roi_caption_model_training: position j sees gt[:, :j+1] zero-padded."""
    B, P = gt.shape
    outs = []
    for j in range(1, P + 1):
        ctx = np.concatenate([gt[:, :j], np.zeros((B, P - j), gt.dtype)], -1)
        outs.append(word_model_literal(f, ctx, w, dtype))
    return np.stack(outs, 1)


def train_forward_v1(f, gt, w, dtype=F32, return_logits=False):
    """Teacher-forced single scan; equals the literal form."""
    B, P = gt.shape
    U = w["imgcap_lstm1/recurrent_kernel"].shape[0]
    st = V1State(B, U, dtype)
    outs = [v1_step(f, gt[:, j], st, w, dtype, return_logits) for j in range(P)]
    return np.stack(outs, 1)


def targets_from_captions(gt):
    """data_generator (text_generation_model.py:352-358): targets = shift-left(gt) ++ [0]."""
    gt = np.asarray(gt).astype(np.int32)
    return np.concatenate([gt[:, 1:], np.zeros((gt.shape[0], 1), np.int32)], 1)


def roi_caption_loss(y_true_ids, probs, valid=None):
    """roi_caption_loss on one-hot targets given as ids.  K.categorical_crossentropy:
    p /= sum(p); p = clip(p, 1e-7, 1-1e-7); -sum(y*log p); mean over positions with sum(y)>0
    (all positions when targets are one-hots; ``valid`` lets a caller mask positions)."""
    p = probs / probs.sum(-1, keepdims=True)
    dt = p.dtype.type
    p = np.clip(p, dt(1e-7), dt(1 - 1e-7))
    picked = np.take_along_axis(p, y_true_ids[..., None].astype(np.int64), -1)[..., 0]
    ce = -np.log(picked)
    if valid is None:
        valid = np.ones(ce.shape, bool)
    if valid.sum() == 0:
        return dt(0)
    return ce[valid].mean(dtype=p.dtype)


def keras_adam_amsgrad(p, g, m, v, vhat, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7):
    """keras.optimizers.Adam(amsgrad=True).get_updates, iteration t (1-based)."""
    dt = p.dtype.type
    lr_t = dt(lr) * dt(np.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t))
    m = dt(b1) * m + dt(1 - b1) * g
    v = dt(b2) * v + dt(1 - b2) * g * g
    vhat = np.maximum(vhat, v)
    p = p - lr_t * m / (np.sqrt(vhat) + dt(eps))
    return p, m, v, vhat


# ----------------------------------------------------------------------------------------
# recurrent dropout masks (a10): KL.LSTM(..., recurrent_dropout=0.2), text_generation_model.py:141-142
# ----------------------------------------------------------------------------------------
# Keras 2.1 LSTMCell (implementation=1) draws FOUR masks per call, K.dropout(ones([B, units]), rate) -- one per gate
# i, f, c, o, values {0, 1/(1-rate)}, constant over the time steps of the call -- and multiplies h_{t-1} by mask g
# before the recurrent product of gate g.  TF's RNG stream cannot be reproduced, so the product and this oracle
# agree on a counter-based generator instead: Philox-4x32-10 (Salmon et al., SC'11; Random123 known-answer vectors in
# tests/test_oracle_decoder.py), key = the 64-bit seed, counter = (global row, unit, layer, step); the four output
# words decide the four gates: keep iff word >= rate * 2^32.

PHILOX_M0, PHILOX_M1 = 0xD2511F53, 0xCD9E8D57
PHILOX_W0, PHILOX_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(counter, key):
    """counter [..., 4] uint32, key (k0, k1) -> [..., 4] uint32."""
    c = [np.asarray(counter[..., i], np.uint64) for i in range(4)]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(PHILOX_M0) * c[0]
        p1 = np.uint64(PHILOX_M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = [(hi1 ^ c[1] ^ k0) & mask, lo1, (hi0 ^ c[3] ^ k1) & mask, lo0]
        k0 = (k0 + np.uint64(PHILOX_W0)) & mask
        k1 = (k1 + np.uint64(PHILOX_W1)) & mask
    return np.stack(c, -1).astype(np.uint32)


def philox_masks(rate, seed, step, layer, rows, units, row_offset=0, dtype=np.float64):
    """[rows, 4, units] recurrent-dropout masks of one LSTM (gate order i, f, c, o) for global rows
    row_offset .. row_offset + rows - 1."""
    r = (np.arange(rows, dtype=np.uint64) + np.uint64(row_offset))[:, None]
    u = np.arange(units, dtype=np.uint64)[None, :]
    ctr = np.stack(np.broadcast_arrays(r & np.uint64(0xFFFFFFFF), u, (r >> np.uint64(32)) | (np.uint64(layer) << np.uint64(16)),
                                       np.uint64(step & 0xFFFFFFFF) + 0 * r), -1)
    w = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))            # [rows, units, 4]
    thr = np.uint64(int(rate * 4294967296.0))
    keep = w.astype(np.uint64) >= thr
    return np.where(keep, 1.0 / (1.0 - rate), 0.0).astype(dtype).transpose(0, 2, 1)


# ----------------------------------------------------------------------------------------
# fp64 reference gradients of the v1 training graph (used to check the CUDA backward)
# ----------------------------------------------------------------------------------------

def train_loss_and_grads_v1(feat, gt, w, train_head=True, return_dfeat=False, rec_masks=None):
    """Loss (mean CE over all B*P positions, clipping ignored -- probabilities of random models
    stay far from 1e-7) and analytic gradients in fp64 by manual BPTT.  Returns (loss, grads), plus
    dL/d(feat) [B,pool,pool,C] with ``return_dfeat`` (the gradient the joint model sends back into
    PyramidROIAlign, dense_img_cap/dense_model.py:738-755).  ``rec_masks`` = (m1, m2), each [B, 4, U]: the recurrent
    dropout masks of the two LSTMs (philox_masks); None = dropout off."""
    D = np.float64
    B, P = gt.shape
    W = {k: v.astype(D) for k, v in w.items()}
    ids = gt.astype(np.int32)
    tgt = targets_from_captions(gt)
    x0 = feat.reshape(B, -1).astype(D)
    # head forward
    k1 = W["mrcnn_class_conv1/kernel"].reshape(-1, 1024)
    s1 = W["mrcnn_class_bn1/gamma"] / np.sqrt(W["mrcnn_class_bn1/moving_variance"] + BN_EPS)
    t1 = W["mrcnn_class_bn1/beta"] - W["mrcnn_class_bn1/moving_mean"] * s1
    z1 = x0 @ k1 + W["mrcnn_class_conv1/bias"]
    a1 = np.maximum(z1 * s1 + t1, 0)
    k2 = W["mrcnn_class_conv2/kernel"].reshape(-1, 1024)
    s2 = W["mrcnn_class_bn2/gamma"] / np.sqrt(W["mrcnn_class_bn2/moving_variance"] + BN_EPS)
    t2 = W["mrcnn_class_bn2/beta"] - W["mrcnn_class_bn2/moving_mean"] * s2
    z2 = a1 @ k2 + W["mrcnn_class_conv2/bias"]
    f = np.maximum(z2 * s2 + t2, 0)
    U = W["imgcap_lstm1/recurrent_kernel"].shape[0]
    E = W["imgcap_embedding_layer/embeddings"].shape[1]
    K1, R1, b1 = W["imgcap_lstm1/kernel"], W["imgcap_lstm1/recurrent_kernel"], W["imgcap_lstm1/bias"]
    K2, R2, b2 = W["imgcap_lstm2/kernel"], W["imgcap_lstm2/recurrent_kernel"], W["imgcap_lstm2/bias"]
    Kd1, bd1 = W["imgcap_lstm_d1/kernel"], W["imgcap_lstm_d1/bias"]
    Kd2, bd2 = W["imgcap_lstm_d2/kernel"], W["imgcap_lstm_d2/bias"]
    emb = W["imgcap_embedding_layer/embeddings"]

    def cell_fwd(x, h, c, K, R, b, msk=None):
        if msk is None:
            z = x @ K + b + h @ R
        else:
            z = x @ K + b + np.concatenate([(h * msk[:, g]) @ R[:, g * U:(g + 1) * U] for g in range(4)], -1)
        zi, zf, zg, zo = z[:, :U], z[:, U:2 * U], z[:, 2 * U:3 * U], z[:, 3 * U:]
        i, fg, g, o = hard_sigmoid(zi), hard_sigmoid(zf), np.tanh(zg), hard_sigmoid(zo)
        cn = fg * c + i * g
        tc = np.tanh(cn)
        return o * tc, cn, (x, h, c, z, i, fg, g, o, tc)

    h1 = np.zeros((B, U)); c1 = np.zeros((B, U)); h2 = np.zeros((B, U)); c2 = np.zeros((B, U))
    cache = []
    loss = 0.0
    for t in range(P):
        m = (ids[:, t] != 0)[:, None]
        x1 = np.concatenate([emb[ids[:, t]], f], -1)
        h1n, c1n, k1c = cell_fwd(x1, h1, c1, K1, R1, b1, rec_masks[0] if rec_masks else None)
        h1 = np.where(m, h1n, h1); c1 = np.where(m, c1n, c1)
        h2n, c2n, k2c = cell_fwd(h1, h2, c2, K2, R2, b2, rec_masks[1] if rec_masks else None)
        h2 = np.where(m, h2n, h2); c2 = np.where(m, c2n, c2)
        din = np.concatenate([h2, f], -1)
        zd = din @ Kd1 + bd1
        d = np.maximum(zd, 0)
        z = d @ Kd2 + bd2
        p = softmax(z)
        loss += -np.log(p[np.arange(B), tgt[:, t]]).sum()
        cache.append((m, k1c, k2c, din, zd, d, p))
    n = B * P
    loss /= n
    G = {k: np.zeros_like(v) for k, v in W.items()}
    df = np.zeros_like(f)
    dh1 = np.zeros((B, U)); dc1 = np.zeros((B, U)); dh2 = np.zeros((B, U)); dc2 = np.zeros((B, U))

    def hs_grad(z):
        return np.where((z > -2.5) & (z < 2.5), 0.2, 0.0)

    def cell_bwd(dh, dc, kc, K, R, msk=None):
        x, h, c, z, i, fg, g, o, tc = kc
        do = dh * tc
        dcn = dc + dh * o * (1 - tc * tc)
        di = dcn * g; dfg = dcn * c; dg = dcn * i
        dc_prev = dcn * fg
        dz = np.concatenate([di * hs_grad(z[:, :U]), dfg * hs_grad(z[:, U:2 * U]),
                             dg * (1 - g * g), do * hs_grad(z[:, 3 * U:])], -1)
        if msk is None:
            return dz, dz @ K.T, dz @ R.T, dc_prev, x, h, h.T @ dz
        dhp = sum(msk[:, q] * (dz[:, q * U:(q + 1) * U] @ R[:, q * U:(q + 1) * U].T) for q in range(4))
        dR = np.concatenate([(h * msk[:, q]).T @ dz[:, q * U:(q + 1) * U] for q in range(4)], -1)
        return dz, dz @ K.T, dhp, dc_prev, x, h, dR

    for t in reversed(range(P)):
        m, k1c, k2c, din, zd, d, p = cache[t]
        dz = p.copy(); dz[np.arange(B), tgt[:, t]] -= 1.0; dz /= n
        G["imgcap_lstm_d2/kernel"] += d.T @ dz; G["imgcap_lstm_d2/bias"] += dz.sum(0)
        dd = (dz @ Kd2.T) * (zd > 0)
        G["imgcap_lstm_d1/kernel"] += din.T @ dd; G["imgcap_lstm_d1/bias"] += dd.sum(0)
        ddin = dd @ Kd1.T
        dh2 = dh2 + ddin[:, :U]; df += ddin[:, U:]
        # layer 2 (masked rows carry their state gradient through untouched)
        dzz, dx, dhp, dcp, x, h, dR = cell_bwd(dh2 * m, dc2 * m, k2c, K2, R2, rec_masks[1] if rec_masks else None)
        G["imgcap_lstm2/kernel"] += x.T @ dzz; G["imgcap_lstm2/recurrent_kernel"] += dR
        G["imgcap_lstm2/bias"] += dzz.sum(0)
        dh1 = dh1 + dx
        dh2 = np.where(m, dhp, dh2); dc2 = np.where(m, dcp, dc2)
        dzz, dx, dhp, dcp, x, h, dR = cell_bwd(dh1 * m, dc1 * m, k1c, K1, R1, rec_masks[0] if rec_masks else None)
        G["imgcap_lstm1/kernel"] += x.T @ dzz; G["imgcap_lstm1/recurrent_kernel"] += dR
        G["imgcap_lstm1/bias"] += dzz.sum(0)
        df += dx[:, E:]
        dh1 = np.where(m, dhp, dh1); dc1 = np.where(m, dcp, dc1)
    if train_head:
        # BatchNorm runs with training=False but gamma / beta stay trainable
        # (modified_dense_model.py:52-62): y = (z - mean) * gamma * r + beta, r = rsqrt(var + eps)
        dy2 = df * (f > 0)
        r2 = 1.0 / np.sqrt(W["mrcnn_class_bn2/moving_variance"] + BN_EPS)
        G["mrcnn_class_bn2/gamma"] = (dy2 * (z2 - W["mrcnn_class_bn2/moving_mean"]) * r2).sum(0)
        G["mrcnn_class_bn2/beta"] = dy2.sum(0)
        dz2 = dy2 * s2
        G["mrcnn_class_conv2/kernel"] = (a1.T @ dz2).reshape(W["mrcnn_class_conv2/kernel"].shape)
        G["mrcnn_class_conv2/bias"] = dz2.sum(0)
        da1 = dz2 @ k2.T
        dy1 = da1 * (a1 > 0)
        r1 = 1.0 / np.sqrt(W["mrcnn_class_bn1/moving_variance"] + BN_EPS)
        G["mrcnn_class_bn1/gamma"] = (dy1 * (z1 - W["mrcnn_class_bn1/moving_mean"]) * r1).sum(0)
        G["mrcnn_class_bn1/beta"] = dy1.sum(0)
        dz1 = dy1 * s1
        G["mrcnn_class_conv1/kernel"] = (x0.T @ dz1).reshape(W["mrcnn_class_conv1/kernel"].shape)
        G["mrcnn_class_conv1/bias"] = dz1.sum(0)
        if return_dfeat:
            return loss, G, (dz1 @ k1.T).reshape(feat.shape)
    return loss, G


# ----------------------------------------------------------------------------------------
# v2 "inject" model (a13) and its greedy loop (a14)
# ----------------------------------------------------------------------------------------

def pad_sequences_pre(seqs, maxlen):
    """keras.preprocessing.sequence.pad_sequences defaults: pre-pad, pre-truncate."""
    out = np.zeros((len(seqs), maxlen), np.int32)
    for i, s in enumerate(seqs):
        s = list(s)[-maxlen:]
        if s:
            out[i, -len(s):] = s
    return out


def v2_inject_predict(feat, words, w, dtype=F32, return_logits=False):
    """build_model(inject=True): feat [B,7,7,256], words [B,L] pre-padded ids -> probs [B,V]."""
    fh = head(feat, w, dtype)                                                  # [B,1024]
    ids = np.asarray(words).astype(np.int32)
    emb = w["imgcap_embedding_layer/embeddings"].astype(dtype)[ids]
    wv = lstm_masked(emb, ids != 0, w["lstm_1/kernel"].astype(dtype),
                     w["lstm_1/recurrent_kernel"].astype(dtype), w["lstm_1/bias"].astype(dtype), False)
    x = np.concatenate([fh, wv], -1)
    u = w["imgcap_lstm/recurrent_kernel"].shape[0]
    z0 = np.zeros((x.shape[0], u), dtype)
    h, _ = lstm_cell(x, z0, z0, w["imgcap_lstm/kernel"].astype(dtype),
                     w["imgcap_lstm/recurrent_kernel"].astype(dtype), w["imgcap_lstm/bias"].astype(dtype))
    z = h @ w["imgcap_d1/kernel"].astype(dtype) + w["imgcap_d1/bias"].astype(dtype)
    return z if return_logits else softmax(z)


def train_loss_and_grads_v2(feat, words, y, w):
    """v2 inject model training step in fp64 (text_generation_model_v2.py:140-166, 263-267): loss = mean over the
    batch of keras.losses.categorical_crossentropy(one_hot(y), model([feat, words])) and its gradients w.r.t. the
    TRAINABLE tensors -- lstm_1, imgcap_lstm, imgcap_d1; the RoI head (trainable=False) and the embedding are
    frozen.  words [B,L] pre-padded ids (0 = masked), y [B] next-word ids.  Clipping ignored (see v1)."""
    D = np.float64
    W = {k: v.astype(D) for k, v in w.items()}
    ids = np.asarray(words).astype(np.int32)
    B, L = ids.shape
    f = head(feat, w, D)
    Wu = W["lstm_1/recurrent_kernel"].shape[0]
    U = W["imgcap_lstm/recurrent_kernel"].shape[0]
    Kw, Rw, bw = W["lstm_1/kernel"], W["lstm_1/recurrent_kernel"], W["lstm_1/bias"]
    Ki, bi = W["imgcap_lstm/kernel"], W["imgcap_lstm/bias"]
    Kd, bd = W["imgcap_d1/kernel"], W["imgcap_d1/bias"]
    emb = W["imgcap_embedding_layer/embeddings"]

    def cell_fwd(x, h, c, K, R, b, u):
        z = x @ K + b + (h @ R if R is not None else 0.0)
        i, fg, g, o = hard_sigmoid(z[:, :u]), hard_sigmoid(z[:, u:2 * u]), np.tanh(z[:, 2 * u:3 * u]), hard_sigmoid(z[:, 3 * u:])
        cn = fg * c + i * g
        tc = np.tanh(cn)
        return o * tc, cn, (x, h, c, z, i, fg, g, o, tc)

    def cell_bwd(dh, dc, kc, u):
        x, h, c, z, i, fg, g, o, tc = kc
        lin = lambda zz: np.where((zz > -2.5) & (zz < 2.5), 0.2, 0.0)
        do = dh * tc
        dcn = dc + dh * o * (1 - tc * tc)
        dz = np.concatenate([dcn * g * lin(z[:, :u]), dcn * c * lin(z[:, u:2 * u]), dcn * i * (1 - g * g),
                             do * lin(z[:, 3 * u:])], -1)
        return dz, dcn * fg

    h = np.zeros((B, Wu)); c = np.zeros((B, Wu))
    cache = []
    for t in range(L):
        m = (ids[:, t] != 0)[:, None]
        hn, cn, kc = cell_fwd(emb[ids[:, t]], h, c, Kw, Rw, bw, Wu)
        cache.append((m, kc))
        h = np.where(m, hn, h); c = np.where(m, cn, c)
    x = np.concatenate([f, h], -1)
    z0 = np.zeros((B, U))
    hi, _, kci = cell_fwd(x, z0, z0, Ki, None, bi, U)              # zero state: the recurrent kernel never contributes
    p = softmax(hi @ Kd + bd)
    yy = np.asarray(y).astype(np.int64)
    loss = -np.log(p[np.arange(B), yy]).mean()
    G = {k: np.zeros_like(W[k]) for k in ("lstm_1/kernel", "lstm_1/recurrent_kernel", "lstm_1/bias", "imgcap_lstm/kernel",
                                          "imgcap_lstm/recurrent_kernel", "imgcap_lstm/bias", "imgcap_d1/kernel", "imgcap_d1/bias")}
    dz = p.copy(); dz[np.arange(B), yy] -= 1.0; dz /= B
    G["imgcap_d1/kernel"] = hi.T @ dz; G["imgcap_d1/bias"] = dz.sum(0)
    dzi, _ = cell_bwd(dz @ Kd.T, np.zeros((B, U)), kci, U)
    G["imgcap_lstm/kernel"] = x.T @ dzi; G["imgcap_lstm/bias"] = dzi.sum(0)
    dh = (dzi @ Ki.T)[:, f.shape[1]:]
    dc = np.zeros((B, Wu))
    for t in reversed(range(L)):
        m, kc = cache[t]
        dzw, dcp = cell_bwd(dh * m, dc * m, kc, Wu)
        G["lstm_1/kernel"] += kc[0].T @ dzw; G["lstm_1/recurrent_kernel"] += kc[1].T @ dzw; G["lstm_1/bias"] += dzw.sum(0)
        dh = np.where(m, dzw @ Rw.T, dh); dc = np.where(m, dcp, dc)
    return loss, G


def greedy_v2(feat, w, P, dtype=F32, start=None):
    """test_score_dense_captions.py:216-225: start [argmax(zeros)] = [0]; P-1 predicts; the
    padded window keeps the LAST P ids.  ``start`` [B]: first word per RoI instead of 0
    (eval_text_generation_model_v2.py:176-186: prev = [gt[0]]).
    Returns (tokens [B,P-1], probs [B,P-1,V])."""
    B = feat.shape[0]
    seqs = [[0 if start is None else int(start[i])] for i in range(B)]
    toks, outs = [], []
    for _ in range(P - 1):
        p = v2_inject_predict(feat, pad_sequences_pre(seqs, P), w, dtype)
        t = p.argmax(-1)
        for i in range(B):
            seqs[i].append(int(t[i]))
        toks.append(t.astype(np.int32)); outs.append(p)
    return np.stack(toks, 1), np.stack(outs, 1)


# ----------------------------------------------------------------------------------------
# beam search (a15) applied to the v1 inject decoder
# ----------------------------------------------------------------------------------------

def beam_v1(f, w, P, k, dtype=F32, start=1):
    """gen_captions semantics (image captioning/test.py:23-64) per RoI: scores are SUMS OF
    PROBABILITIES accumulated in float64 (Python float + np.float32 under numpy 1.x),
    candidates per beam = ascending top-k of p, pool in generation order, stable ascending
    sort, keep the last k; no end-token handling.  Returns (tokens [B,k,P] int32 ascending by
    score, scores [B,k] float64).  Top-k ties inside p follow a stable ascending argsort
    (the reference's np.argsort quicksort order is unpinned)."""
    B = f.shape[0]
    U = w["imgcap_lstm1/recurrent_kernel"].shape[0]
    out_tok = np.zeros((B, k, P), np.int32)
    out_sc = np.zeros((B, k), np.float64)
    for r in range(B):
        fr = f[r:r + 1]
        caps = [[start]]
        scores = [0.0]
        states = [V1State(1, U, dtype)]
        while len(caps[0]) < P:
            new_caps, new_scores, new_states = [], [], []
            for cap, sc, st in zip(caps, scores, states):
                st2 = st.gather(slice(None))
                st2.h1, st2.c1, st2.h2, st2.c2 = (st.h1.copy(), st.c1.copy(), st.h2.copy(), st.c2.copy())
                p = v1_step(fr, np.array([cap[-1]]), st2, w, dtype)[0]
                cand = np.argsort(p, kind="stable")[-k:]
                for nw in cand:
                    new_caps.append(cap + [int(nw)])
                    new_scores.append(sc + float(p[nw]))
                    new_states.append(st2)
            order = sorted(range(len(new_caps)), key=lambda i: new_scores[i])[-k:]
            caps = [new_caps[i] for i in order]
            scores = [new_scores[i] for i in order]
            states = [new_states[i] for i in order]
        nb = len(caps)
        out_tok[r, k - nb:] = np.array(caps, np.int32)
        out_sc[r, k - nb:] = scores
    return out_tok, out_sc
