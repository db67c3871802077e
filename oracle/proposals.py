"""CPU oracle for the box front-end that feeds PyramidROIAlign (TEST INFRASTRUCTURE ONLY; SURVEY.md
section 8f rank 3).  Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this.

numpy restatement of (paths relative to /root/reference/dense_img_cap_separate_models):

  * generate_anchors / generate_pyramid_anchors   utils.py:347-403  (fp64 numpy, cast to fp32 by
                                                   ProposalLayer.__init__, modified_dense_model.py:245)
  * apply_box_deltas_graph                        modified_dense_model.py:179-200
  * clip_boxes_graph                              modified_dense_model.py:203-218
  * ProposalLayer.call                            modified_dense_model.py:247-303
  * GT-box normalisation (boxes / [h, w, h, w])   modified_dense_model.py:1523-1526

Third-party pieces (TensorFlow 1.x, not vendored and not installed here; the reference pins no version):
  * tf.nn.top_k(sorted=True): descending values; among equal values the LOWER index first (documented).
  * tf.image.non_max_suppression (core/kernels/non_max_suppression_op.cc): candidates in descending
    score order; a candidate is selected unless its IoU with an already selected box is > threshold;
    IoU uses min/max-normalised corners and is 0 when either area is <= 0; stops at max_output_size.
    Equal scores: restated as a STABLE order (the order top_k delivered).
  * tf.exp: restated as the correctly rounded fp32 exponential (computed in fp64, rounded once);
    TF's Eigen expf may differ from that in the last ulp (unknowable without TF).

PINNED PARTS (tests/golden/gen_golden_reference_numpy.py -> tests/golden/reference_numpy.npz): the anchor
generator and the numpy twin of apply_box_deltas (utils.py:107-128) against the reference's own numpy code;
and the LAYER'S GLUE -- ProposalLayer.call with apply_box_deltas_graph, clip_boxes_graph and utils.batch_slice
is EXECUTED from /root/reference over a numpy stand-in for its TF ops (tests/golden/tf_numpy_shim.py, with
top_k_indices / exp_f32 / tf_non_max_suppression / std_min / std_max below as the primitives) and
proposal_layer() reproduces its output bit for bit.  UNPINNED (TensorFlow absent): the arithmetic inside
tf.nn.top_k (tie order), tf.exp and tf.image.non_max_suppression themselves.
"""
import numpy as np

F32 = np.float32


def generate_anchors(scales, ratios, shape, feature_stride, anchor_stride):
    """utils.py:347-383 (one pyramid level; fp64).  Order: cell row, cell column, ratio."""
    scales = np.atleast_1d(np.asarray(scales, np.float64))
    ratios = np.asarray(ratios, np.float64)
    out = []
    for s in scales:                                        # meshgrid(scales, ratios).flatten(): ratio-major
        out.append((s / np.sqrt(ratios), s * np.sqrt(ratios)))
    # meshgrid(scales, ratios) flattens as ratio outer, scale inner
    hs = np.stack([o[0] for o in out], 1).reshape(-1)
    ws = np.stack([o[1] for o in out], 1).reshape(-1)
    ys = np.arange(0, shape[0], anchor_stride) * feature_stride
    xs = np.arange(0, shape[1], anchor_stride) * feature_stride
    cy = np.repeat(ys, len(xs))[:, None].astype(np.float64)
    cx = np.tile(xs, len(ys))[:, None].astype(np.float64)
    n, k = cy.shape[0], hs.shape[0]
    cy, cx = np.broadcast_to(cy, (n, k)), np.broadcast_to(cx, (n, k))
    h, w = np.broadcast_to(hs, (n, k)), np.broadcast_to(ws, (n, k))
    boxes = np.stack([cy - 0.5 * h, cx - 0.5 * w, cy + 0.5 * h, cx + 0.5 * w], 2)
    return boxes.reshape(-1, 4)


def generate_pyramid_anchors(scales, ratios, feature_shapes, feature_strides, anchor_stride):
    """utils.py:386-403: the levels' anchors concatenated in the order of ``scales``."""
    return np.concatenate([generate_anchors(scales[i], ratios, feature_shapes[i], feature_strides[i], anchor_stride)
                           for i in range(len(scales))], 0)


def exp_f32(x):
    """Correctly rounded fp32 exponential (see the header)."""
    with np.errstate(over="ignore"):
        return np.exp(np.asarray(x, F32).astype(np.float64)).astype(F32)


def apply_box_deltas(boxes, deltas, exp=exp_f32):
    """modified_dense_model.py:179-200, fp32 op by op (no fused multiply-add).  ``exp``: the fp32
    exponential to use (numpy's own SIMD expf reproduces the reference's numpy twin bit for bit)."""
    boxes, deltas = np.asarray(boxes, F32), np.asarray(deltas, F32)
    height = boxes[:, 2] - boxes[:, 0]
    width = boxes[:, 3] - boxes[:, 1]
    center_y = boxes[:, 0] + F32(0.5) * height
    center_x = boxes[:, 1] + F32(0.5) * width
    center_y = center_y + deltas[:, 0] * height
    center_x = center_x + deltas[:, 1] * width
    height = height * exp(deltas[:, 2])
    width = width * exp(deltas[:, 3])
    y1 = center_y - F32(0.5) * height
    x1 = center_x - F32(0.5) * width
    return np.stack([y1, x1, y1 + height, x1 + width], 1)


def clip_boxes(boxes, window):
    """modified_dense_model.py:203-218: max(min(v, hi), lo) per corner."""
    wy1, wx1, wy2, wx2 = [F32(v) for v in window]
    b = np.asarray(boxes, F32)
    return np.stack([std_max(std_min(b[:, 0], wy2), wy1), std_max(std_min(b[:, 1], wx2), wx1),
                     std_max(std_min(b[:, 2], wy2), wy1), std_max(std_min(b[:, 3], wx2), wx1)], 1)


def std_min(a, b):
    """std::min(a, b) / Eigen's scalar min as TF's C++ evaluates it: ``b < a ? b : a`` (a NaN first argument
    propagates; equal operands return the first)."""
    return np.where(np.less(b, a), b, a).astype(F32)


def std_max(a, b):
    """std::max(a, b): ``a < b ? b : a``."""
    return np.where(np.less(a, b), b, a).astype(F32)


def top_k_indices(scores, k):
    """tf.nn.top_k(sorted=True).indices: descending, lower index first among equals."""
    s = np.asarray(scores, F32)
    return np.argsort(-s, kind="stable")[:k].astype(np.int32)


def tf_iou(a, b):
    """non_max_suppression_op.cc IOU() in fp32 (scalar form; ``a`` is the earlier / selected box)."""
    a, b = np.asarray(a, F32), np.asarray(b, F32)
    ymin_a, ymax_a, xmin_a, xmax_a = std_min(a[0], a[2]), std_max(a[0], a[2]), std_min(a[1], a[3]), std_max(a[1], a[3])
    ymin_b, ymax_b, xmin_b, xmax_b = std_min(b[0], b[2]), std_max(b[0], b[2]), std_min(b[1], b[3]), std_max(b[1], b[3])
    area_a = (ymax_a - ymin_a) * (xmax_a - xmin_a)
    area_b = (ymax_b - ymin_b) * (xmax_b - xmin_b)
    if area_a <= 0 or area_b <= 0:
        return F32(0)
    ih = std_max(std_min(ymax_a, ymax_b) - std_max(ymin_a, ymin_b), F32(0))
    iw = std_max(std_min(xmax_a, xmax_b) - std_max(xmin_a, xmin_b), F32(0))
    inter = ih * iw
    with np.errstate(divide="ignore", invalid="ignore"):
        return F32(inter / ((area_a + area_b) - inter))


def tf_non_max_suppression(boxes, scores, max_output_size, iou_threshold):
    """tf.image.non_max_suppression: selected indices in descending score order (vectorised over the
    remaining candidates; same decisions as the candidate-vs-selected loop)."""
    b = np.asarray(boxes, F32)
    order = np.argsort(-np.asarray(scores, F32), kind="stable")
    ymin, ymax = std_min(b[:, 0], b[:, 2]), std_max(b[:, 0], b[:, 2])
    xmin, xmax = std_min(b[:, 1], b[:, 3]), std_max(b[:, 1], b[:, 3])
    area = (ymax - ymin) * (xmax - xmin)
    alive = np.ones(len(b), bool)
    thr = F32(iou_threshold)
    picked = []
    for pos, i in enumerate(order):
        if not alive[i]:
            continue
        picked.append(i)
        if len(picked) >= max_output_size:
            break
        rest = order[pos + 1:]
        rest = rest[alive[rest]]
        if area[i] <= 0 or rest.size == 0:
            continue
        with np.errstate(divide="ignore", invalid="ignore"):
            ih = std_max(std_min(ymax[i], ymax[rest]) - std_max(ymin[i], ymin[rest]), F32(0))
            iw = std_max(std_min(xmax[i], xmax[rest]) - std_max(xmin[i], xmin[rest]), F32(0))
            inter = ih * iw
            iou = inter / ((area[i] + area[rest]) - inter)
        iou = np.where(area[rest] <= 0, F32(0), iou)
        alive[rest[iou > thr]] = False
    return np.array(picked, np.int32)


def normalize_boxes(boxes, height, width):
    """modified_dense_model.py:1523-1526 (and :287): fp32 division by [h, w, h, w]."""
    return np.asarray(boxes, F32) / np.array([height, width, height, width], F32)


def proposal_layer(rpn_probs, rpn_bbox, anchors, proposal_count, nms_threshold, image_shape,
                   bbox_std_dev=(0.1, 0.1, 0.2, 0.2), pre_nms_limit=6000, return_indices=False):
    """ProposalLayer.call (modified_dense_model.py:247-303).  rpn_probs [B, A, 2], rpn_bbox [B, A, 4],
    anchors [A, 4] pixels.  Returns proposals [B, proposal_count, 4] (normalised, zero padded)."""
    rpn_probs, rpn_bbox = np.asarray(rpn_probs, F32), np.asarray(rpn_bbox, F32)
    anchors = np.asarray(anchors).astype(F32)
    B, A = rpn_probs.shape[:2]
    k = min(pre_nms_limit, A)
    h, w = F32(image_shape[0]), F32(image_shape[1])
    std = np.asarray(bbox_std_dev, np.float64).astype(F32)
    out = np.zeros((B, proposal_count, 4), F32)
    picked = []
    for b in range(B):
        scores = rpn_probs[b, :, 1]
        deltas = rpn_bbox[b] * std
        ix = top_k_indices(scores, k)
        boxes = apply_box_deltas(anchors[ix], deltas[ix])
        boxes = clip_boxes(boxes, (0, 0, h, w))
        nb = normalize_boxes(boxes, h, w)
        sel = tf_non_max_suppression(nb, scores[ix], proposal_count, nms_threshold)
        out[b, :len(sel)] = nb[sel]
        picked.append(ix[sel])
    return (out, picked) if return_indices else out
