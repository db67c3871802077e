import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import image_captioning_b200 as pkg
from image_captioning_b200 import synth
V, E, U, C, P, B = 10000, 300, 512, 256, 15, 8000
w = synth.synth_weights_v1(np.random.default_rng(1005), V=V, E=E, U=U, C=C)
cfg = pkg.DenseCapConfig(V, w["imgcap_embedding_layer/embeddings"], 1, P)
m = pkg.build_lstm_model([7, 7, C], cfg, U, "inference", dtype="bfloat16")
m.set_weights(w)
feats = torch.randn((B, 1024), device="cuda").relu()
m.generate(feats); torch.cuda.synchronize()
os.environ["DCAP_LOOP_TRACE"] = sys.argv[1]
os.environ["DCAP_NO_GRAPHS"] = "1"
m.generate(feats); torch.cuda.synchronize()
