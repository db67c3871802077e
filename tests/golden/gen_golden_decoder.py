"""Generates tests/golden/decoder_small.npz from the numpy oracle (fixed seed).

The reference cannot be imported in this container (no TensorFlow/Keras) and ships no weights, so
these vectors pin the ORACLE, not the reference ("parity unpinned"): they freeze
today's oracle outputs -- head features, v1 greedy ids/probabilities, beam search, teacher-forced
loss, fp64 gradients (as per-tensor checksums), one Keras AMSGrad update, v2 inject greedy -- so that
later edits to the oracle or the CUDA kernels are checked against the same numbers.  Run from the
repo root:
    python tests/golden/gen_golden_decoder.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from image_captioning_b200 import synth  # noqa: E402  (pure numpy input generation)
from oracle import decoder as dec  # noqa: E402

SHAPE = dict(V=96, E=16, U=64, C=8)
P, B, K = 6, 12, 3


def build():
    rng = np.random.default_rng(20261018)
    w = synth.synth_weights_v1(rng, **SHAPE)
    feat = rng.standard_normal((B, 7, 7, SHAPE["C"])).astype(np.float32)
    gt = synth.synth_captions(rng, B, P, SHAPE["V"])
    gt[1, 2] = 0
    f = dec.head(feat, w)
    tok, probs = dec.greedy_v1(f, w, P)
    bt, bs = dec.beam_v1(f, w, P, K)
    tf_probs = dec.train_forward_v1(f, gt, w)
    loss = dec.roi_caption_loss(dec.targets_from_captions(gt), tf_probs)
    loss64, G = dec.train_loss_and_grads_v1(feat, gt, w)
    names = sorted(G)
    p0 = w["imgcap_lstm_d1/bias"] + np.float32(0.01)
    g0 = G["imgcap_lstm_d1/bias"].astype(np.float32)
    z = np.zeros_like(p0)
    p1, m1, v1, vh1 = dec.keras_adam_amsgrad(p0, g0, z, z, z, 1)
    rng2 = np.random.default_rng(7)
    w2 = synth.synth_weights_v2(rng2, V=SHAPE["V"], E=SHAPE["E"], units=32, C=SHAPE["C"])
    tok2, probs2 = dec.greedy_v2(feat[:5], w2, P)
    out = dict(feat=feat, gt=gt, head=f, greedy_tok=tok, greedy_probs=probs, beam_tok=bt, beam_scores=bs,
               teacher_forced_probs=tf_probs, loss=np.float32(loss), loss64=np.float64(loss64),
               grad_names=np.array(names), grad_l2=np.array([np.linalg.norm(G[n]) for n in names]),
               grad_sum=np.array([G[n].sum() for n in names]), adam_p0=p0, adam_g=g0, adam_p1=p1, adam_m1=m1,
               adam_v1=v1, v2_tok=tok2, v2_probs_last=probs2[:, -1])
    return out


if __name__ == "__main__":
    out = build()
    path = os.path.join(ROOT, "tests", "golden", "decoder_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
