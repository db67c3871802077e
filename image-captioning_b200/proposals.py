"""Box front-end of the RoI path: drop-in for the reference's
``ProposalLayer`` (/root/reference/dense_img_cap_separate_models/modified_dense_model.py:221-306),
the anchor generator it is constructed with (``utils.generate_pyramid_anchors``, utils.py:347-403,
called at modified_dense_model.py:1436-1440) and the GT-box normalisation Lambda (:1523-1526).

The anchors are a constant of the configuration and are generated on the host with numpy, exactly as
the reference does at model-construction time; everything per batch (top-k, box refinement, clipping,
NMS, padding) runs on the device through ``dc_proposal_layer`` -- there is no CPU fallback."""
import ctypes

import numpy as np
import torch

from . import _lib


def generate_anchors(scales, ratios, shape, feature_stride, anchor_stride):
    """One pyramid level (utils.py:347-383): [cells * len(scales) * len(ratios), 4] fp64 (y1,x1,y2,x2) in
    pixels, ordered by cell row, cell column, then (ratio, scale)."""
    s = np.atleast_1d(np.asarray(scales, np.float64))
    root = np.sqrt(np.asarray(ratios, np.float64))
    half = 0.5 * np.stack([(s[None, :] / root[:, None]).ravel(), (s[None, :] * root[:, None]).ravel()], 1)   # [K, (h, w)]
    cy = np.arange(0, shape[0], anchor_stride, dtype=np.float64) * feature_stride
    cx = np.arange(0, shape[1], anchor_stride, dtype=np.float64) * feature_stride
    centers = np.stack(np.meshgrid(cy, cx, indexing="ij"), -1).reshape(-1, 1, 2)                                # [cells, 1, (y, x)]
    return np.concatenate([centers - half[None], centers + half[None]], -1).reshape(-1, 4)


def generate_pyramid_anchors(scales, ratios, feature_shapes, feature_strides, anchor_stride):
    """utils.py:386-403: the anchors of all levels, level i using scales[i]."""
    return np.concatenate([generate_anchors(scales[i], ratios, feature_shapes[i], feature_strides[i], anchor_stride)
                           for i in range(len(scales))], axis=0)


class ProposalConfig(object):
    """The fields of the reference's Config that the box front-end reads (config.py:53-76, 109, 157-164)."""
    BACKBONE_STRIDES = [4, 8, 16, 32, 64]
    RPN_ANCHOR_SCALES = (32, 64, 128, 256, 512)
    RPN_ANCHOR_RATIOS = [0.5, 1, 2]
    RPN_ANCHOR_STRIDE = 1
    RPN_NMS_THRESHOLD = 0.7
    POST_NMS_ROIS_TRAINING = 2000
    POST_NMS_ROIS_INFERENCE = 1000
    RPN_BBOX_STD_DEV = np.array([0.1, 0.1, 0.2, 0.2])
    IMAGE_MAX_DIM = 1024
    IMAGES_PER_GPU = 1

    def __init__(self, **overrides):
        for k, v in overrides.items():
            setattr(self, k, v)
        self.IMAGE_SHAPE = np.array([self.IMAGE_MAX_DIM, self.IMAGE_MAX_DIM, 3])
        self.BACKBONE_SHAPES = np.array([[int(np.ceil(self.IMAGE_SHAPE[0] / s)), int(np.ceil(self.IMAGE_SHAPE[1] / s))]
                                         for s in self.BACKBONE_STRIDES])

    def anchors(self):
        return generate_pyramid_anchors(self.RPN_ANCHOR_SCALES, self.RPN_ANCHOR_RATIOS, self.BACKBONE_SHAPES,
                                        self.BACKBONE_STRIDES, self.RPN_ANCHOR_STRIDE)


def _as_cuda_f32(x, dev):
    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, np.float32))
    return t.to(dev, torch.float32).contiguous()


class ProposalLayer(object):
    """``ProposalLayer(proposal_count, nms_threshold, anchors, config)([rpn_class, rpn_bbox])`` ->
    proposals [batch, proposal_count, 4] in normalised coordinates, zero padded.  numpy in -> numpy out,
    CUDA tensors in -> CUDA tensors out (no host synchronisation in that case)."""
    PRE_NMS_LIMIT = 6000                                    # modified_dense_model.py:258

    def __init__(self, proposal_count, nms_threshold, anchors, config=None, name="ROI"):
        self.config = config if config is not None else ProposalConfig()
        self.proposal_count = int(proposal_count)
        self.nms_threshold = float(nms_threshold)
        self.anchors = np.ascontiguousarray(np.asarray(anchors).astype(np.float32))
        if self.anchors.ndim != 2 or self.anchors.shape[1] != 4:
            raise ValueError("anchors must be [N, 4]")
        self.name = name
        self._dev_anchors, self._ws = {}, {}
        self._std = (ctypes.c_float * 4)(*[float(np.float32(v)) for v in np.asarray(self.config.RPN_BBOX_STD_DEV).ravel()])

    def compute_output_shape(self, input_shape=None):
        return None, self.proposal_count, 4

    def __call__(self, inputs, return_details=False):
        lib = _lib.load()
        rpn_class, rpn_bbox = inputs
        was_numpy = not isinstance(rpn_class, torch.Tensor)
        dev = rpn_class.device if (not was_numpy and rpn_class.is_cuda) else torch.device("cuda", torch.cuda.current_device())
        probs, bbox = _as_cuda_f32(rpn_class, dev), _as_cuda_f32(rpn_bbox, dev)
        A = self.anchors.shape[0]
        if probs.dim() != 3 or probs.shape[1:] != (A, 2) or tuple(bbox.shape) != (probs.shape[0], A, 4):
            raise ValueError("expected rpn_class [B, %d, 2] and rpn_bbox [B, %d, 4], got %s and %s"
                             % (A, A, tuple(probs.shape), tuple(bbox.shape)))
        B = probs.shape[0]
        if dev not in self._dev_anchors:
            self._dev_anchors[dev] = torch.from_numpy(self.anchors).to(dev)
        need = int(lib.dc_proposal_workspace_bytes(B, A, self.PRE_NMS_LIMIT, self.proposal_count))
        ws = self._ws.get(dev)
        if ws is None or ws.numel() < need:
            ws = self._ws[dev] = torch.empty((need,), dtype=torch.uint8, device=dev)
        out = torch.empty((B, self.proposal_count, 4), dtype=torch.float32, device=dev)
        n_valid = torch.empty((B,), dtype=torch.int32, device=dev)
        index = torch.empty((B, self.proposal_count), dtype=torch.int32, device=dev)
        h, w = float(self.config.IMAGE_SHAPE[0]), float(self.config.IMAGE_SHAPE[1])
        with torch.cuda.device(dev):
            _lib.check(lib.dc_proposal_layer(
                ctypes.c_void_p(probs.data_ptr()), ctypes.c_void_p(bbox.data_ptr()),
                ctypes.c_void_p(self._dev_anchors[dev].data_ptr()), B, A, ctypes.cast(self._std, ctypes.c_void_p),
                ctypes.c_float(h), ctypes.c_float(w), self.PRE_NMS_LIMIT, self.proposal_count,
                ctypes.c_float(self.nms_threshold), ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(n_valid.data_ptr()),
                ctypes.c_void_p(index.data_ptr()), ctypes.c_void_p(ws.data_ptr()), ctypes.c_size_t(ws.numel()),
                ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        if was_numpy:
            out, n_valid, index = out.cpu().numpy(), n_valid.cpu().numpy(), index.cpu().numpy()
        return (out, n_valid, index) if return_details else out

    call = __call__


def normalize_boxes(boxes, image_shape):
    """GT boxes in pixels -> normalised (y1,x1,y2,x2): ``boxes / [h, w, h, w]`` in fp32
    (modified_dense_model.py:1523-1526).  numpy in -> numpy out, CUDA tensors in -> CUDA tensors out."""
    lib = _lib.load()
    was_numpy = not isinstance(boxes, torch.Tensor)
    dev = boxes.device if (not was_numpy and boxes.is_cuda) else torch.device("cuda", torch.cuda.current_device())
    b = _as_cuda_f32(boxes, dev)
    if b.shape[-1] != 4:
        raise ValueError("boxes must be [..., 4]")
    out = torch.empty_like(b)
    with torch.cuda.device(dev):
        _lib.check(lib.dc_normalize_boxes(ctypes.c_void_p(b.data_ptr()), b.numel() // 4, ctypes.c_float(float(image_shape[0])),
                                          ctypes.c_float(float(image_shape[1])), ctypes.c_void_p(out.data_ptr()),
                                          ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out.cpu().numpy() if was_numpy else out
