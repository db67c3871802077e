"""Experiment: does running two half-batches on two streams (two decoder handles) beat one full batch?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import image_captioning_b200 as pkg
from image_captioning_b200 import synth

V, E, U, C, P, R = 10000, 300, 512, 256, 15, 8000
w = synth.synth_weights_v1(np.random.default_rng(1005), V=V, E=E, U=U, C=C)
cfg = pkg.DenseCapConfig(V, w["imgcap_embedding_layer/embeddings"], 1, P)
models = []
for _ in range(2):
    m = pkg.build_lstm_model([7, 7, C], cfg, U, "inference", dtype="bfloat16")
    m.set_weights(w)
    models.append(m)
feats = torch.randn((R, 7, 7, C), device="cuda").to(torch.bfloat16)
halves = [feats[:R // 2].contiguous(), feats[R // 2:].contiguous()]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]


def one():
    return models[0].generate(feats)


def two():
    cur = torch.cuda.current_stream()
    outs = []
    for m, h, s in zip(models, halves, streams):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            outs.append(m.generate(h))
    for s in streams:
        cur.wait_stream(s)
    return outs


for name, fn in (("one lane", one), ("two lanes", two), ("one lane", one), ("two lanes", two)):
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%-10s %.3f ms per 8000 RoIs" % (name, e0.elapsed_time(e1) / 20), flush=True)
a = one()
b = torch.cat(two(), 0)
print("same tokens:", bool(torch.equal(a, b)))
