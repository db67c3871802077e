"""Debug: per-tensor / per-block gradient error of the CUDA training step vs the fp64 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import image_captioning_b200 as pkg
from image_captioning_b200 import synth
from oracle import decoder as dec

SHAPE = dict(V=1000, E=48, U=128, C=64)
P, B = 6, int(os.environ.get("B", 96))
rng = np.random.default_rng(32)
w = synth.synth_weights_v1(rng, trained_like=False, **SHAPE)
feat = rng.standard_normal((B, 7, 7, SHAPE["C"])).astype(np.float32)
gt = synth.synth_captions(rng, B, P, SHAPE["V"])
cfg = pkg.DenseCapConfig(SHAPE["V"], w["imgcap_embedding_layer/embeddings"], B, P)
m = pkg.build_lstm_model([7, 7, SHAPE["C"]], cfg, SHAPE["U"], "training", dtype="bfloat16")
m.set_weights(w)
loss_want, G = dec.train_loss_and_grads_v1(feat, gt, w)
# oracle on bf16-rounded weights + inputs (isolates operand quantisation from activation rounding)
r16 = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).float().numpy()
wq = {k: (r16(v) if ("kernel" in k or "embeddings" in k) else v) for k, v in w.items()}
_, Gq = dec.train_loss_and_grads_v1(r16(feat), gt, wq)
loss = float(m.train_step_device(feat, gt).item())
print("loss", loss, loss_want)
got = m.get_gradients()
rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))
for n, g in got.items():
    print("%-36s vs fp64 %.4f   vs bf16-weight oracle %.4f   oracle-vs-oracle %.4f" % (n, rel(g, G[n]), rel(g, Gq[n]), rel(Gq[n], G[n])))
g = got["mrcnn_class_conv1/kernel"].reshape(-1, 1024)
o = G["mrcnn_class_conv1/kernel"].reshape(-1, 1024)
print("conv1 kernel per 448-row block:", [round(rel(g[i:i + 448], o[i:i + 448]), 3) for i in range(0, g.shape[0], 448)])
print("conv1 kernel per 128-col block:", [round(rel(g[:, i:i + 128], o[:, i:i + 128]), 3) for i in range(0, 1024, 128)])
