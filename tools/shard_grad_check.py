"""Full-batch gradient against the sum of two half-batch shard gradients (one GPU), per tensor, with and without recurrent
dropout and with the fused / two-kernel soft-max: isolates what differs between a single-process and a 2-rank step."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import image_captioning_b200 as pkg           # noqa: E402
from image_captioning_b200 import synth       # noqa: E402

SHAPE = dict(V=1000, E=48, U=128, C=64)
P, B = 6, 64
rng = np.random.default_rng(77)
w = synth.synth_weights_v1(rng, trained_like=False, **SHAPE)
feat = rng.standard_normal((B, 7, 7, SHAPE["C"])).astype(np.float32)
gt = synth.synth_captions(rng, B, P, SHAPE["V"])


def model(batch):
    cfg = pkg.DenseCapConfig(SHAPE["V"], w["imgcap_embedding_layer/embeddings"], batch, P)
    m = pkg.build_lstm_model([7, 7, SHAPE["C"]], cfg, SHAPE["U"], "training", dtype="bfloat16")
    m.set_weights(w)
    m.compile(optimizer=pkg.Adam(amsgrad=True), loss=pkg.roi_caption_loss)
    return m


for fused in ("1", "0"):
    os.environ["DCAP_XENT_FUSED"] = fused
    for dropout in (0.0, 0.2):
        opts = lambda lo: dict(recurrent_dropout=dropout, dropout_seed=99, dropout_step=0, row_offset=lo) if dropout else {}
        m = model(B)
        l_full = float(m.train_step_device(feat, gt, None, 1.0 / (B * P), **opts(0)).item())
        g_full = m.get_gradients()
        parts, l_sum = None, 0.0
        for lo, hi in ((0, B // 2), (B // 2, B)):
            ms = model(hi - lo)
            l_sum += float(ms.train_step_device(feat[lo:hi], gt[lo:hi], None, 1.0 / (B * P), **opts(lo)).item())
            g = ms.get_gradients()
            parts = g if parts is None else {k: parts[k] + g[k] for k in g}
        worst = sorted(((float(np.linalg.norm(parts[k] - g_full[k]) / max(np.linalg.norm(g_full[k]), 1e-30)), k) for k in g_full), reverse=True)
        print("fused=%s dropout=%.1f loss full %.7f shards %.7f; worst rel-L2 shard-sum vs full: %s" %
              (fused, dropout, l_full, l_sum, ", ".join("%s %.2e" % (k, v) for v, k in worst[:4])))
