"""bf16-vs-fp32 precision envelope of the decoder at the BASELINE shapes (informs the test bars).

For the trained-like synthetic model prints, against the fp32 numpy oracle:
  * free-running greedy-token agreement (a flipped token changes everything after it),
  * prefix-conditioned agreement (bf16 model teacher-forced on the oracle's own tokens: per-decision flip rate),
  * the oracle's own top-1/top-2 gap at every first divergence (disagreements must be near-ties),
  * |delta log p| rms / p99.9 / max where the fp32 model has mass.
Run on a GPU box:  python tools/bf16_error_study.py [B]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import image_captioning_b200 as pkg
from image_captioning_b200 import synth
from oracle import decoder as dec


def logp(z):
    m = z.max(-1, keepdims=True)
    return z - (np.log(np.exp(z - m).sum(-1, keepdims=True)) + m)


def study_v1(seed, B, V=10000, E=300, U=512, C=256, P=15):
    rng = np.random.default_rng(seed)
    w = synth.synth_weights_v1(rng, V=V, E=E, U=U, C=C)
    feat = rng.standard_normal((B, 7, 7, C)).astype(np.float32)
    tok_want, z = dec.greedy_v1(dec.head(feat, w), w, P, return_logits=True)
    lp_want = logp(z)
    cfg = pkg.DenseCapConfig(V, w["imgcap_embedding_layer/embeddings"], 1, P)
    m = pkg.build_lstm_model([7, 7, C], cfg, U, "inference", dtype="bfloat16")
    m.set_weights(w)
    tok = m.generate(feat)
    m32 = pkg.build_lstm_model([7, 7, C], cfg, U, "inference", dtype="float32")
    m32.set_weights(w)
    tok32 = m32.generate(feat)
    cfg_t = pkg.DenseCapConfig(V, w["imgcap_embedding_layer/embeddings"], B, P)
    mt = pkg.build_lstm_model([7, 7, C], cfg_t, U, "training", dtype="bfloat16")
    mt.set_weights(w)
    gt = np.concatenate([np.ones((B, 1), np.float32), tok_want[:, :-1].astype(np.float32)], 1)
    probs = mt.predict_teacher_forced([feat, gt])
    lp = np.log(np.maximum(probs, 1e-38))
    sel = lp_want > -12
    err = np.abs(lp - lp_want)[sel]
    agree = tok == tok_want
    first = np.where(agree.all(1), P, (~agree).argmax(1))
    rows = np.nonzero(first < P)[0]
    gaps = np.array([z[r, first[r], tok_want[r, first[r]]] - z[r, first[r], tok[r, first[r]]] for r in rows])
    top2 = np.sort(z, -1)[..., -2:]
    print("v1 seed %d B %d: distinct %d | fp32 CUDA ids == oracle %.5f | free-running agreement %.4f | captions identical %.4f | "
          "prefix-conditioned agreement %.4f | oracle gap at first divergence: max %.4f median %.4f (median top-2 gap overall %.3f) | "
          "|dlogp| rms %.4f p99.9 %.4f max %.4f | logit std %.2f"
          % (seed, B, len(np.unique(tok_want)), (tok32 == tok_want).mean(), agree.mean(), agree.all(1).mean(),
             (probs.argmax(-1) == tok_want).mean(), gaps.max() if len(gaps) else 0, np.median(gaps) if len(gaps) else 0,
             np.median(top2[..., 1] - top2[..., 0]), np.sqrt((err ** 2).mean()), np.quantile(err, 0.999), err.max(), z.std()))


def study_v2(seed, B, V=10000, E=300, units=256, C=256, P=10):
    rng = np.random.default_rng(seed)
    w = synth.synth_weights_v2(rng, V=V, E=E, units=units, C=C)
    feat = rng.standard_normal((B, 7, 7, C)).astype(np.float32)
    tok_want, p_want = dec.greedy_v2(feat, w, P)
    cfg = pkg.DenseCapConfig(V, w["imgcap_embedding_layer/embeddings"], 1, P)
    out = {}
    for dt in ("float32", "bfloat16"):
        m = pkg.build_model((7, 7, C), (P,), cfg, units, inject=True, dtype=dt)
        m.set_weights(w)
        out[dt] = m.generate(feat, return_probs=True)
    tok, probs = out["bfloat16"]
    agree = tok == tok_want
    prefix_ok = np.concatenate([np.ones((B, 1), bool), np.cumprod(agree[:, :-1], 1).astype(bool)], 1)
    err = np.abs(np.log(np.maximum(probs, 1e-38)) - np.log(np.maximum(p_want, 1e-38)))[prefix_ok[:, :, None] & (p_want > np.exp(-12.0))]
    print("v2 seed %d B %d: distinct %d | fp32 CUDA ids == oracle %.5f | free-running agreement %.4f | decisions on an agreeing prefix %.4f | "
          "|dlogp| rms %.4f p99.9 %.4f max %.4f"
          % (seed, B, len(np.unique(tok_want)), (out["float32"][0] == tok_want).mean(), agree.mean(),
             agree[prefix_ok].mean(), np.sqrt((err ** 2).mean()), np.quantile(err, 0.999), err.max()))


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    for seed in (1001, 1002):
        study_v1(seed, B)
    study_v2(1006, 96)
