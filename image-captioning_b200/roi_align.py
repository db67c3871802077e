"""Drop-in for the reference's ``PyramidROIAlign`` Keras layer.

Mirrors /root/reference/evaluate_models/modified_dense_model.py:318-419 (byte-identical copies in
mask_rcnn/mask_rcnn_model.py:321-423 and the other Mask R-CNN forks): same constructor
(``pool_shape, image_shape, **kwargs``), same call convention (``layer([boxes, P2, P3, P4, P5])``
with normalised ``(y1, x1, y2, x2)`` boxes ``[batch, num_boxes, 4]`` and NHWC maps), same literal
output ``[1, batch*num_boxes, pool_h, pool_w, C]`` and the same ``compute_output_shape``.

All arithmetic runs in the sm_100a kernel behind ``dc_pyramid_roi_align_*`` (include/dcap.h);
this file only validates shapes and moves pointers.  numpy inputs go through the host-buffer
C entry point (as a Keras ``predict`` would); torch CUDA tensors stay on the device.
"""
import ctypes

import numpy as np
import torch

from . import _lib

_BOX_LIMIT = 100000      # reference sort key batch*100000 + box (modified_dense_model.py:408)


def _as_ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check_inputs(boxes_shape, fm_shapes):
    if len(boxes_shape) != 3 or boxes_shape[2] != 4:
        raise ValueError("boxes must be [batch, num_boxes, 4], got %s" % (tuple(boxes_shape),))
    if len(fm_shapes) != 4:
        raise ValueError("PyramidROIAlign expects 4 feature maps (P2..P5), got %d" % len(fm_shapes))
    C = fm_shapes[0][-1]
    for i, s in enumerate(fm_shapes):
        if len(s) != 4:
            raise ValueError("feature map %d must be [batch, h, w, channels], got %s" % (i, tuple(s)))
        if s[0] != boxes_shape[0]:
            raise ValueError("feature map %d batch %d != boxes batch %d" % (i, s[0], boxes_shape[0]))
        if s[-1] != C:
            raise ValueError("feature maps disagree on channel count")
    if C % 4 != 0:
        raise ValueError("channels must be a multiple of 4, got %d" % C)
    if boxes_shape[1] > _BOX_LIMIT:
        raise ValueError("num_boxes %d > %d: the reference's re-sort key collides beyond that"
                         % (boxes_shape[1], _BOX_LIMIT))


def fpn_levels(boxes, image_shape):
    """FPN level (2..5) of every box, int32, same leading shape as ``boxes[..., 0]``.
    boxes: torch CUDA tensor [..., 4] fp32 normalised (y1, x1, y2, x2)."""
    if not (isinstance(boxes, torch.Tensor) and boxes.is_cuda):
        raise TypeError("fpn_levels expects a CUDA tensor (no CPU fallback)")
    lib = _lib.load()
    b = boxes.detach().to(torch.float32).contiguous()
    if b.shape[-1] != 4:
        raise ValueError("boxes must end in 4 coordinates")
    out = torch.empty(b.shape[:-1], dtype=torch.int32, device=b.device)
    with torch.cuda.device(b.device):
        _lib.check(lib.dc_fpn_levels_f32(_as_ptr(b), b.numel() // 4, int(image_shape[0]),
                                         int(image_shape[1]), _as_ptr(out), _stream_ptr(b.device)))
    return out


def pyramid_roi_align(boxes, feature_maps, pool_shape, image_shape, out_dtype=torch.float32,
                      return_levels=False, out=None):
    """Device path.  boxes [B,N,4] and 4 NHWC maps as torch CUDA fp32 tensors ->
    ``[B*N, ph, pw, C]`` (fp32 or bf16) in (image, box) order, optionally with the levels."""
    lib = _lib.load()
    if not all(isinstance(t, torch.Tensor) and t.is_cuda for t in [boxes] + list(feature_maps)):
        raise TypeError("pyramid_roi_align expects CUDA tensors (no CPU fallback)")
    _check_inputs(boxes.shape, [f.shape for f in feature_maps])
    dev = boxes.device
    boxes = boxes.detach().to(torch.float32).contiguous()
    fms = [f.detach().to(torch.float32).contiguous() for f in feature_maps]
    B, N = boxes.shape[:2]
    C = fms[0].shape[-1]
    ph, pw = int(pool_shape[0]), int(pool_shape[1])
    if out is None:
        out = torch.empty((B * N, ph, pw, C), dtype=out_dtype, device=dev)
    elif out.dtype != out_dtype or out.numel() != B * N * ph * pw * C or not out.is_contiguous():
        raise ValueError("out buffer has the wrong dtype/size")
    levels = torch.empty((B, N), dtype=torch.int32, device=dev) if return_levels else None
    ptrs = (ctypes.c_void_p * 4)(*[f.data_ptr() for f in fms])
    hs = (ctypes.c_int * 4)(*[f.shape[1] for f in fms])
    ws = (ctypes.c_int * 4)(*[f.shape[2] for f in fms])
    if out_dtype == torch.float32:
        fn = lib.dc_pyramid_roi_align_f32
    elif out_dtype == torch.bfloat16:
        fn = lib.dc_pyramid_roi_align_bf16out
    else:
        raise ValueError("out_dtype must be torch.float32 or torch.bfloat16")
    with torch.cuda.device(dev):
        _lib.check(fn(_as_ptr(boxes), ptrs, hs, ws, B, N, C, ph, pw, int(image_shape[0]),
                      int(image_shape[1]), _as_ptr(out),
                      _as_ptr(levels) if levels is not None else None, _stream_ptr(dev)))
    return (out, levels) if return_levels else out


def pyramid_roi_align_backward(boxes, grad_out, fm_shapes, pool_shape, image_shape, grads=None):
    """Gradient of the layer with respect to its four feature maps (the boxes get none: the reference
    wraps them in tf.stop_gradient, modified_dense_model.py:379-380).  boxes [B,N,4], grad_out
    [B*N, ph, pw, C] CUDA fp32; fm_shapes = [(H_l, W_l)] * 4.  Returns 4 tensors [B, H_l, W_l, C]
    (``grads``: existing tensors to ACCUMULATE into)."""
    lib = _lib.load()
    if not all(isinstance(t, torch.Tensor) and t.is_cuda for t in (boxes, grad_out)):
        raise TypeError("pyramid_roi_align_backward expects CUDA tensors (no CPU fallback)")
    dev = boxes.device
    boxes = boxes.detach().to(torch.float32).contiguous()
    B, N = boxes.shape[:2]
    ph, pw = int(pool_shape[0]), int(pool_shape[1])
    g = grad_out.detach().to(torch.float32).contiguous()
    C = g.shape[-1]
    if g.numel() != B * N * ph * pw * C or len(fm_shapes) != 4:
        raise ValueError("grad_out must be [B*N, ph, pw, C] and fm_shapes four (H, W) pairs")
    if grads is None:
        grads = [torch.zeros((B, int(h), int(w), C), dtype=torch.float32, device=dev) for h, w in fm_shapes]
    ptrs = (ctypes.c_void_p * 4)(*[t.data_ptr() for t in grads])
    hs = (ctypes.c_int * 4)(*[int(h) for h, _ in fm_shapes])
    ws = (ctypes.c_int * 4)(*[int(w) for _, w in fm_shapes])
    with torch.cuda.device(dev):
        _lib.check(lib.dc_pyramid_roi_align_backward_f32(_as_ptr(boxes), _as_ptr(g), ptrs, hs, ws, B, N, C, ph, pw,
                                                         int(image_shape[0]), int(image_shape[1]), _stream_ptr(dev)))
    return grads


class _PyramidROIAlignFn(torch.autograd.Function):
    """autograd bridge: forward = dc_pyramid_roi_align_f32, backward = dc_pyramid_roi_align_backward_f32."""

    @staticmethod
    def forward(ctx, boxes, pool_shape, image_shape, *fms):
        ctx.save_for_backward(boxes)
        ctx.meta = (tuple(pool_shape), tuple(image_shape), [tuple(f.shape[1:3]) for f in fms])
        return pyramid_roi_align(boxes, list(fms), pool_shape, image_shape)

    @staticmethod
    def backward(ctx, grad_out):
        (boxes,) = ctx.saved_tensors
        pool, ishape, shapes = ctx.meta
        grads = pyramid_roi_align_backward(boxes, grad_out, shapes, pool, ishape)
        return (None, None, None) + tuple(grads)


def pyramid_roi_align_autograd(boxes, feature_maps, pool_shape, image_shape):
    """Differentiable (w.r.t. the feature maps) form of pyramid_roi_align for torch users -- what the joint
    model of dense_img_cap/dense_model.py:738-755 needs to train through the layer."""
    return _PyramidROIAlignFn.apply(boxes, tuple(pool_shape), tuple(image_shape), *feature_maps)


class PyramidROIAlign(object):
    """Implements ROI Pooling on multiple levels of the feature pyramid.

    Params (as the reference layer):
    - pool_shape: [height, width] of the output pooled regions. Usually [7, 7]
    - image_shape: [height, width, channels]. Shape of input image in pixels

    Inputs: ``[boxes, P2, P3, P4, P5]`` -- boxes [batch, num_boxes, (y1, x1, y2, x2)] normalised,
    possibly zero padded; maps [batch, height, width, channels].

    Output: ``[1, batch*num_boxes, pool_h, pool_w, channels]`` -- the reference's literal return
    shape (modified_dense_model.py:415-416); ``compute_output_shape`` advertises
    ``(batch, num_boxes, pool_h, pool_w, channels)`` exactly as the reference does (:418-419).
    numpy in -> numpy out (host-buffer C entry point); torch CUDA in -> torch CUDA out.
    """

    def __init__(self, pool_shape, image_shape, **kwargs):
        self.pool_shape = tuple(pool_shape)
        self.image_shape = tuple(image_shape)
        self.name = kwargs.pop("name", "roi_align")
        self.out_dtype = kwargs.pop("out_dtype", torch.float32)
        if len(self.pool_shape) != 2 or len(self.image_shape) < 2:
            raise ValueError("pool_shape must be (h, w) and image_shape (h, w[, c])")

    def __call__(self, inputs):
        return self.call(inputs)

    def call(self, inputs):
        boxes, feature_maps = inputs[0], list(inputs[1:])
        if all(isinstance(t, torch.Tensor) and t.is_cuda for t in [boxes] + feature_maps):
            pooled = pyramid_roi_align(boxes, feature_maps, self.pool_shape, self.image_shape,
                                       out_dtype=self.out_dtype)
            return pooled.unsqueeze(0)
        return self._call_host(boxes, feature_maps)

    def _call_host(self, boxes, feature_maps):
        lib = _lib.load()
        to_np = lambda a: a.numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
        boxes = np.ascontiguousarray(to_np(boxes), dtype=np.float32)
        fms = [np.ascontiguousarray(to_np(f), dtype=np.float32) for f in feature_maps]
        _check_inputs(boxes.shape, [f.shape for f in fms])
        B, N = boxes.shape[:2]
        C = fms[0].shape[-1]
        ph, pw = self.pool_shape
        out = np.empty((1, B * N, ph, pw, C), dtype=np.float32)
        ptrs = (ctypes.c_void_p * 4)(*[f.ctypes.data for f in fms])
        hs = (ctypes.c_int * 4)(*[f.shape[1] for f in fms])
        ws = (ctypes.c_int * 4)(*[f.shape[2] for f in fms])
        _lib.check(lib.dc_pyramid_roi_align_host_f32(
            ctypes.c_void_p(boxes.ctypes.data), ptrs, hs, ws, B, N, C, int(ph), int(pw),
            int(self.image_shape[0]), int(self.image_shape[1]), ctypes.c_void_p(out.ctypes.data),
            None))
        return out

    def compute_output_shape(self, input_shape):
        return tuple(input_shape[0][:2]) + self.pool_shape + (input_shape[1][-1],)
