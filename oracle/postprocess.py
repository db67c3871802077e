"""CPU oracle for the caption post-processing that follows the decoder in the dense-captioning
evaluation path (TEST INFRASTRUCTURE ONLY).

numpy restatement of (paths relative to /root/reference/evaluate_models):

  * compute_iou           utils.py:30-48   -- in THIS copy of the Mask R-CNN utils the overlap is the
                                              Dice coefficient 2*I/(A+B), not IoU
  * non_max_suppression   utils.py:69-104
  * refine_generations    test_score_dense_captions.py:245-283   (score = sum_t log max_v p; NMS;
                                                                  keep the DETECTION_MAX_INSTANCES best)
  * caption text          test_score_dense_captions.py:229-232 / eval_text_generation_model.py:146

PARITY UNPINNED (no tests or fixtures in the reference).  Tie order: the reference sorts with
``scores.argsort()[::-1]`` (numpy quicksort, unspecified among equal scores); the oracle and the
CUDA kernel both define it as a STABLE ascending argsort reversed (among equal scores the larger
index comes first).
"""
import numpy as np

F32 = np.float32


def compute_overlap(box, boxes, box_area, boxes_area):
    """utils.py:30-48 in fp32: 2 * intersection / (area + areas)."""
    y1 = np.maximum(box[0], boxes[:, 0])
    y2 = np.minimum(box[2], boxes[:, 2])
    x1 = np.maximum(box[1], boxes[:, 1])
    x2 = np.minimum(box[3], boxes[:, 3])
    inter = np.maximum(x2 - x1, F32(0)) * np.maximum(y2 - y1, F32(0))
    union = box_area + boxes_area
    with np.errstate(divide="ignore", invalid="ignore"):
        return F32(2) * inter / union


def non_max_suppression(boxes, scores, threshold):
    """utils.py:69-104.  Returns the kept indices in descending score order."""
    boxes = np.asarray(boxes, F32)
    scores = np.asarray(scores, F32)
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    ixs = np.argsort(scores, kind="stable")[::-1]
    pick = []
    while len(ixs) > 0:
        i = ixs[0]
        pick.append(i)
        ov = compute_overlap(boxes[i], boxes[ixs[1:]], area[i], area[ixs[1:]])
        remove = np.where(ov > F32(threshold))[0] + 1
        ixs = np.delete(ixs, remove)
        ixs = np.delete(ixs, 0)
    return np.array(pick, dtype=np.int32)


def caption_scores(probs):
    """test_score_dense_captions.py:256-258: sum over steps of log(max over the vocabulary)."""
    return np.log(np.max(probs, axis=2)).sum(axis=1, dtype=probs.dtype)


def refine_generations(rois, scores, nms_threshold=0.7, max_instances=100):
    """test_score_dense_captions.py:245-283 on precomputed caption scores: indices of the kept RoIs
    (NMS survivors, best ``max_instances`` by score, descending)."""
    keep = non_max_suppression(rois, scores, nms_threshold)
    top = np.argsort(np.asarray(scores, F32)[keep], kind="stable")[::-1][:max_instances]
    # `keep` is already in descending score order; a stable sort reversed flips equal-score neighbours,
    # exactly as the reference's second argsort would with a stable kind
    return keep[top]


def caption_text(token_ids, id_to_word, stop=" ."):
    """' '.join(words) cut at the first ' .' (v2, test_score_dense_captions.py:230-231) or ' <end>' (v1)."""
    cap = " ".join(id_to_word[int(t)] for t in token_ids)
    return cap.split(stop, 1)[0]
