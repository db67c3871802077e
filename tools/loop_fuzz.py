"""Randomised A/B of the persistent greedy-loop kernel against the launch-per-GEMM path: model shapes (U, V, E, P),
batch sizes and input kinds drawn at random; token ids must be identical, scores equal to fp32 rounding, repeated calls
identical.  Run on a GPU box: python tools/loop_fuzz.py [--cases 60] [--seed 1]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import image_captioning_b200 as pkg          # noqa: E402
from image_captioning_b200 import synth      # noqa: E402


def build(w, P, V, U, C):
    cfg = pkg.DenseCapConfig(V, w["imgcap_embedding_layer/embeddings"], 1, P)
    m = pkg.build_lstm_model([7, 7, C], cfg, U, "inference", dtype="bfloat16")
    m.set_weights(w)
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=60)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    bad = 0
    t0 = time.time()
    for case in range(a.cases):
        U = int(rng.choice([64, 128, 256, 512]))
        V = int(rng.choice([256, 300, 1000, 2000, 5000, 10000]))
        E = int(rng.choice([16, 48, 100, 300]))
        C = int(rng.choice([64, 128]))
        P = int(rng.choice([1, 2, 3, 6, 15, 20]))
        w = synth.synth_weights_v1(np.random.default_rng(1000 + case), V=V, E=E, U=U, C=C)
        sizes = [int(x) for x in rng.choice([1, 2, 31, 32, 33, 127, 128, 129, 255, 256, 257, 300, 511, 513, 777, 1000, 1281, 1537, 2049, 3001], size=3, replace=False)]
        for B in sizes:
            roi_kind = bool(rng.integers(0, 2)) and B <= 600
            g = torch.Generator(device="cuda").manual_seed(case * 100 + B)
            feats = (torch.randn((B, 7, 7, C), device="cuda", generator=g) if roi_kind
                     else torch.randn((B, 1024), device="cuda", generator=g).relu())
            out = {}
            for mode in ("0", "2"):
                os.environ["DCAP_GREEDY_LOOP"] = mode
                m = build(w, P, V, U, C)
                toks = [m.generate(feats).cpu().numpy() for _ in range(3)]
                ts, sc = m.generate(feats, return_scores=True)
                torch.cuda.synchronize()
                out[mode] = (toks, ts.cpu().numpy(), sc.cpu().numpy())
            o, n = out["0"], out["2"]
            ok = all(np.array_equal(n[0][i], o[0][0]) for i in range(3)) and np.array_equal(n[1], o[1]) and \
                np.allclose(n[2], o[2], rtol=0, atol=5e-5)
            if not ok:
                bad += 1
                agree = float((n[0][0] == o[0][0]).mean())
                print("MISMATCH case %d U=%d V=%d E=%d C=%d P=%d B=%d roi=%s: agreement %.5f, repeat %s, max dscore %.3g"
                      % (case, U, V, E, C, P, B, roi_kind, agree, [bool(np.array_equal(n[0][i], n[0][0])) for i in range(3)],
                         float(np.abs(n[2] - o[2]).max())), flush=True)
        if case % 10 == 9:
            print("case %d done, %d mismatches, %.0f s" % (case + 1, bad, time.time() - t0), flush=True)
    print("LOOP_FUZZ", "OK" if bad == 0 else "FAIL (%d)" % bad)
    return 0 if bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
