"""Re-entrancy (threading): the reference calls the feature model -- hence
PyramidROIAlign -- from Keras' generator-enqueuer thread while the main thread trains / decodes
(text_generation_model.py:332-371, 470-472).  The drop-in must give the same results from a
non-main thread, concurrently with decoder work on the main thread."""
import threading

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_roi_align_from_worker_thread_while_main_thread_decodes():
    import image_captioning_b200 as pkg
    from image_captioning_b200 import synth
    rng = np.random.default_rng(77)
    boxes = synth.synth_boxes(rng, 2, 300, 1024.0)
    fms = [rng.standard_normal((2, 64 >> i, 64 >> i, 64), dtype=np.float32) for i in range(4)]
    layer = pkg.PyramidROIAlign([7, 7], (1024, 1024, 3))
    want = layer([boxes] + fms)
    V, E, U, C, P = 1000, 300, 512, 64, 5
    w = synth.synth_weights_v1(rng, V=V, E=E, U=U, C=C)
    cfg = pkg.DenseCapConfig(V, w["imgcap_embedding_layer/embeddings"], 1, P)
    m = pkg.build_lstm_model([7, 7, C], cfg, U, "inference", dtype="bfloat16")
    m.set_weights(w)
    feats = torch.from_numpy(want[0]).cuda()
    tok_want = m.generate(feats).clone()
    results, errors = [], []

    def worker():
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(20):
                    results.append(layer([boxes] + fms))
        except Exception as e:          # pragma: no cover
            errors.append(e)

    th = threading.Thread(target=worker)
    th.start()
    toks = [m.generate(feats).clone() for _ in range(20)]
    th.join()
    torch.cuda.synchronize()
    assert not errors, errors
    assert all(np.array_equal(r, want) for r in results)
    assert all(torch.equal(t, tok_want) for t in toks)
