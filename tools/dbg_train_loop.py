import os, sys
import numpy as np, torch
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import test_train_gpu as T
from oracle import decoder as dec
P = T.P
for merged in ("1", "0"):
    os.environ["DCAP_HOIST_MERGED"] = merged
    os.environ["DCAP_GREEDY_LOOP"] = "2"
    for B in (16, 64, 40):
        pkg, rng, w, feat, gt, m = T._setup(36, 64)
        tok0 = m.generate(feat[:B]); want0, _ = dec.greedy_v1(dec.head(feat[:B], w), w, P)
        print("merged", merged, "B", B, "agreement", (tok0 == want0).mean(), "rows wrong", np.nonzero((tok0 != want0).any(1))[0].tolist()[:20])
        h = m.head_features(feat[:B]) if hasattr(m, "head_features") else None
        tok1 = m.generate(torch.from_numpy(np.ascontiguousarray(h)).cuda()) if h is not None else None
        if tok1 is not None:
            print("   from head features (cast path): agreement", (tok1.cpu().numpy() == want0).mean())
