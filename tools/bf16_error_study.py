"""bf16-vs-fp32 error study on the cfg1 decoder shapes (informs the test thresholds)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import image_captioning_b200 as pkg
from image_captioning_b200 import synth
from oracle import decoder as dec

V, E, U, C, P, B = 10000, 300, 512, 256, 15, 256
for scale, zipf in [(3.0, True)]:
    rng = np.random.default_rng(1001)
    w = synth.synth_weights_v1(rng, V=V, E=E, U=U, C=C, trained_like=False)
    w["imgcap_lstm_d2/kernel"] *= np.float32(scale)
    if zipf:
        w["imgcap_lstm_d2/bias"] = (-np.log(1.0 + np.arange(V))).astype(np.float32)
    feat = rng.standard_normal((B, 7, 7, C)).astype(np.float32)
    tok_want, z = dec.greedy_v1(dec.head(feat, w), w, P, return_logits=True)
    cfg = pkg.DenseCapConfig(V, w["imgcap_embedding_layer/embeddings"], 1, P)
    outs = {}
    for dt in ("float32", "bfloat16"):
        m = pkg.build_lstm_model([7, 7, C], cfg, U, "inference", dtype=dt)
        m.set_weights(w)
        tok, probs = m.generate(feat, return_probs=True)
        outs[dt] = (tok, probs)
    tok, probs = outs["bfloat16"]
    agree = tok == tok_want
    prefix_ok = np.concatenate([np.ones((B, 1), bool), np.cumprod(agree[:, :-1], 1).astype(bool)], 1)
    lse = np.log(np.exp(z - z.max(-1, keepdims=True)).sum(-1, keepdims=True)) + z.max(-1, keepdims=True)
    lp_want = z - lse
    lp = np.log(np.maximum(probs, 1e-38))
    sel = prefix_ok[:, :, None] & (lp_want > -15)
    err = np.abs(lp - lp_want)[sel]
    step0 = np.abs(lp[:, 0] - lp_want[:, 0])[lp_want[:, 0] > -15]
    top2 = np.sort(z, -1)[..., -2:]
    print("scale %.0f zipf %d | logit std %.3f | agree %.4f (fp32 path %.4f) | |dlogp| max %.4f p99.9 %.4f rms %.5f | step0 max %.4f | median top1-top2 gap %.4f"
          % (scale, zipf, z.std(), agree.mean(), (outs["float32"][0] == tok_want).mean(), err.max(),
             np.quantile(err, 0.999), np.sqrt((err ** 2).mean()), step0.max(), np.median(top2[..., 1] - top2[..., 0])))
