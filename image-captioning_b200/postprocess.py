"""Caption post-processing of the dense-captioning evaluation path:
drop-in for ``refine_generations`` (/root/reference/evaluate_models/test_score_dense_captions.py:
245-283: caption score = sum of log max p, non-max suppression with this copy's 2*I/(A+B) overlap,
``evaluate_models/utils.py:30-104``, top DETECTION_MAX_INSTANCES) and the caption text rule
(:229-232).  The suppression runs on the device (``dc_refine_generations``); no CPU fallback."""
import ctypes

import numpy as np
import torch

from . import _lib


def refine_generations(rois, scores, nms_threshold=0.7, max_instances=100):
    """rois [N,4] or [B,N,4] (y1,x1,y2,x2), scores [N] / [B,N] caption scores (e.g. from
    ``model.generate(features, return_scores=True)``).  Returns the kept RoI indices in descending
    score order: an int32 array [n_keep] (single image) or a list of them (batched).
    numpy in -> numpy out, CUDA tensors in -> CUDA tensors out."""
    lib = _lib.load()
    was_numpy = not isinstance(rois, torch.Tensor)
    r = torch.as_tensor(np.ascontiguousarray(rois, np.float32)) if was_numpy else rois
    s = torch.as_tensor(np.ascontiguousarray(scores, np.float32)) if not isinstance(scores, torch.Tensor) else scores
    single = r.dim() == 2
    if single:
        r, s = r[None], s[None]
    if r.dim() != 3 or r.shape[2] != 4 or tuple(s.shape) != tuple(r.shape[:2]):
        raise ValueError("rois must be [N,4] / [B,N,4] with scores [N] / [B,N]")
    dev = r.device if r.is_cuda else torch.device("cuda", torch.cuda.current_device())
    r = r.to(dev, torch.float32).contiguous()
    s = s.to(dev, torch.float32).contiguous()
    B, N = r.shape[:2]
    keep = torch.empty((B, int(max_instances)), dtype=torch.int32, device=dev)
    n_keep = torch.empty((B,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.dc_refine_generations(
            ctypes.c_void_p(r.data_ptr()), ctypes.c_void_p(s.data_ptr()), B, N, ctypes.c_float(nms_threshold),
            int(max_instances), ctypes.c_void_p(keep.data_ptr()), ctypes.c_void_p(n_keep.data_ptr()),
            ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    counts = n_keep.cpu().tolist()
    out = [keep[i, :c] for i, c in enumerate(counts)]
    if was_numpy:
        out = [o.cpu().numpy() for o in out]
    return out[0] if single else out


def caption_text(token_ids, id_to_word, stop=" ."):
    """' '.join(words) cut at the first ``stop`` (' .' in test_score_dense_captions.py:230-231,
    ' <end>' in eval_text_generation_model.py:146)."""
    return " ".join(id_to_word[int(t)] for t in token_ids).split(stop, 1)[0]
