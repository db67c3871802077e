"""CPU oracle for the PyramidROIAlign feature stage (TEST INFRASTRUCTURE ONLY).

This module is a numpy fp32 restatement of the reference's RoI feature stage.  It is
imported only by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs -- never by the product path in ``image-captioning_b200/`` (which has no CPU
fallback and fails loudly when the CUDA library is missing).

PARITY UNPINNED for the op arithmetic: the reference has no tests / golden vectors for this
path and its arithmetic lives in TensorFlow 1.x (``tf.image.crop_and_resize``, ``tf.log``,
``tf.round``), which is neither vendored under /root/reference nor installable here.  The
restatement follows the reference call sites line by line and TF's published CropAndResize CPU
semantics; it is cross-checked in tests against two independent transcriptions
(TVM's ``crop_and_resize_python`` and ``torch.nn.functional.grid_sample``).
PINNED: the layer's glue.  tests/golden/gen_golden_reference_numpy.py EXECUTES the reference's
``PyramidROIAlign.call`` from /root/reference over a numpy stand-in for its TF ops
(tests/golden/tf_numpy_shim.py, with ``crop_and_resize`` below as the primitive), and
``pyramid_roi_align_literal`` / ``pyramid_roi_align`` reproduce its output bit for bit on the
golden inputs (level formula, per-level dispatch, concat, top_k re-sort; degenerate boxes too).

Reference lines followed (paths relative to /root/reference):
  * ``log2_graph``                      evaluate_models/modified_dense_model.py:313-315
  * ``PyramidROIAlign.call`` (levels)   evaluate_models/modified_dense_model.py:351-363
  * per-level crop                      evaluate_models/modified_dense_model.py:366-393
  * re-ordering / output contract       evaluate_models/modified_dense_model.py:395-419

Arithmetic conventions (so that the C oracle, this module and the CUDA kernel can agree
bit for bit):
  * every operation is an individually rounded IEEE fp32 op (no FMA contraction) --
    numpy never fuses, the C oracle is compiled with -ffp-contract=off and the CUDA
    kernel uses __fmul_rn/__fadd_rn/__fsub_rn;
  * ``tf.log`` is restated as the correctly rounded fp32 logarithm (computed in fp64 and
    rounded once).  TF's CPU kernel is an Eigen polynomial whose last-ulp behaviour is
    unknowable without TF; it can only matter for boxes whose pre-round log2 value lies
    within a few ulps of a half-integer, which ``level_ambiguity`` reports.
"""
import numpy as np

F32 = np.float32
INT_MIN = np.int32(-2147483648)


def _f32_log(x):
    """Correctly rounded fp32 natural log (fp64 log, one rounding)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.log(np.asarray(x, dtype=np.float64)).astype(F32)


def _x86_f32_to_i32(r):
    """float -> int32 with x86 cvttss2si semantics: NaN / inf / out-of-range -> INT_MIN."""
    r = np.asarray(r, dtype=F32)
    ok = np.isfinite(r) & (r >= F32(-2147483648.0)) & (r < F32(2147483648.0))
    out = np.full(r.shape, INT_MIN, dtype=np.int32)
    out[ok] = r[ok].astype(np.int32)
    return out


def roi_level_raw(boxes, image_shape):
    """fp32 value fed to tf.round (modified_dense_model.py:353-361)."""
    boxes = np.asarray(boxes, dtype=F32)
    y1, x1, y2, x2 = boxes[..., 0], boxes[..., 1], boxes[..., 2], boxes[..., 3]
    h = y2 - y1
    w = x2 - x1
    image_area = F32(float(image_shape[0]) * float(image_shape[1]))
    denom = F32(224.0) / np.sqrt(image_area, dtype=F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.sqrt(h * w, dtype=F32)
        q = (s / denom).astype(F32)
        r = (_f32_log(q) / _f32_log(F32(2.0))).astype(F32)
    return r


def fpn_level(boxes, image_shape):
    """FPN level per box: min(5, max(2, 4 + int32(round_half_even(log2(...))))).

    Reference: evaluate_models/modified_dense_model.py:351-363.  Zero-area (padded) boxes
    give log(0) = -inf -> int32 min -> clamp to 2; negative area gives NaN -> int32 min -> 2.
    """
    r = roi_level_raw(boxes, image_shape)
    with np.errstate(invalid="ignore"):
        rounded = np.rint(r).astype(F32)           # tf.round == round-half-to-even
    iv = _x86_f32_to_i32(rounded).astype(np.int64)
    lvl = np.minimum(5, np.maximum(2, 4 + iv))
    return lvl.astype(np.int32)


def level_ambiguity(boxes, image_shape, tol=2.0 ** -18):
    """True for boxes whose pre-round log2 value is within ``tol`` of a half-integer AND whose
    two candidate levels differ after clamping (the only boxes where TF's own ``log`` could
    give a different level than this oracle)."""
    r = roi_level_raw(boxes, image_shape).astype(np.float64)
    with np.errstate(invalid="ignore"):
        frac = np.abs(r - np.floor(r) - 0.5)
        near = np.isfinite(r) & (frac < tol)
        lo = np.clip(4 + np.floor(r), 2, 5)
        hi = np.clip(4 + np.floor(r) + 1, 2, 5)
    return near & (lo != hi)


def crop_and_resize(image, boxes, box_indices, crop_size, extrapolation_value=0.0):
    """TF ``tf.image.crop_and_resize(method='bilinear')`` CPU semantics, vectorised over
    boxes.  image [B,H,W,C] f32 NHWC; boxes [n,4] normalised (y1,x1,y2,x2); -> [n,ch,cw,C].

    Call site: evaluate_models/modified_dense_model.py:391-393.
    """
    image = np.asarray(image, dtype=F32)
    boxes = np.asarray(boxes, dtype=F32).reshape(-1, 4)
    box_indices = np.asarray(box_indices, dtype=np.int64).reshape(-1)
    ch, cw = int(crop_size[0]), int(crop_size[1])
    _, H, W, C = image.shape
    n = boxes.shape[0]
    out = np.empty((n, ch, cw, C), dtype=F32)
    if n == 0:
        return out
    y1, x1, y2, x2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    Hm1, Wm1 = F32(H - 1), F32(W - 1)
    if ch > 1:
        hs = ((y2 - y1) * Hm1) / F32(ch - 1)
        in_y = (y1 * Hm1)[:, None] + np.arange(ch, dtype=F32)[None, :] * hs[:, None]
    else:
        in_y = (F32(0.5) * (y1 + y2) * Hm1)[:, None]
    if cw > 1:
        ws = ((x2 - x1) * Wm1) / F32(cw - 1)
        in_x = (x1 * Wm1)[:, None] + np.arange(cw, dtype=F32)[None, :] * ws[:, None]
    else:
        in_x = (F32(0.5) * (x1 + x2) * Wm1)[:, None]
    in_y = in_y.astype(F32)
    in_x = in_x.astype(F32)
    with np.errstate(invalid="ignore"):
        y_ok = ~((in_y < 0) | (in_y > Hm1)) & ~np.isnan(in_y)
        x_ok = ~((in_x < 0) | (in_x > Wm1)) & ~np.isnan(in_x)
    # NaN coordinates: TF's comparisons are both false -> it would index with garbage; we
    # treat NaN as out of range (documented deviation; normalised boxes never produce NaN).
    iy = np.where(y_ok, in_y, F32(0))
    ix = np.where(x_ok, in_x, F32(0))
    top = np.floor(iy).astype(np.int64)
    bot = np.ceil(iy).astype(np.int64)
    ly = (iy - np.floor(iy)).astype(F32)
    left = np.floor(ix).astype(np.int64)
    right = np.ceil(ix).astype(np.int64)
    lx = (ix - np.floor(ix)).astype(F32)
    b = box_indices[:, None, None]
    tl = image[b, top[:, :, None], left[:, None, :]]      # [n,ch,cw,C]
    tr = image[b, top[:, :, None], right[:, None, :]]
    bl = image[b, bot[:, :, None], left[:, None, :]]
    br = image[b, bot[:, :, None], right[:, None, :]]
    lxb = lx[:, None, :, None]
    lyb = ly[:, :, None, None]
    t = tl + (tr - tl) * lxb
    bt = bl + (br - bl) * lxb
    val = t + (bt - t) * lyb
    ok = (y_ok[:, :, None] & x_ok[:, None, :])[..., None]
    out[...] = np.where(ok, val, F32(extrapolation_value))
    return out


def crop_and_resize_loops(image, boxes, box_indices, crop_size, extrapolation_value=0.0):
    """Scalar-loop transcription of the same kernel (small cases only); used to check the
    vectorised form above."""
    image = np.asarray(image, dtype=F32)
    boxes = np.asarray(boxes, dtype=F32).reshape(-1, 4)
    ch, cw = crop_size
    _, H, W, C = image.shape
    out = np.empty((boxes.shape[0], ch, cw, C), dtype=F32)
    for n in range(boxes.shape[0]):
        y1, x1, y2, x2 = boxes[n]
        b = int(box_indices[n])
        hs = F32(F32(y2 - y1) * F32(H - 1)) / F32(ch - 1) if ch > 1 else F32(0)
        ws = F32(F32(x2 - x1) * F32(W - 1)) / F32(cw - 1) if cw > 1 else F32(0)
        for y in range(ch):
            in_y = F32(F32(y1 * F32(H - 1)) + F32(F32(y) * hs)) if ch > 1 \
                else F32(F32(0.5) * F32(y1 + y2) * F32(H - 1))
            if in_y < 0 or in_y > H - 1:
                out[n, y] = extrapolation_value
                continue
            t, bo = int(np.floor(in_y)), int(np.ceil(in_y))
            ly = F32(in_y - F32(t))
            for x in range(cw):
                in_x = F32(F32(x1 * F32(W - 1)) + F32(F32(x) * ws)) if cw > 1 \
                    else F32(F32(0.5) * F32(x1 + x2) * F32(W - 1))
                if in_x < 0 or in_x > W - 1:
                    out[n, y, x] = extrapolation_value
                    continue
                l, r = int(np.floor(in_x)), int(np.ceil(in_x))
                lx = F32(in_x - F32(l))
                tl, tr = image[b, t, l], image[b, t, r]
                bl, br = image[b, bo, l], image[b, bo, r]
                top = tl + (tr - tl) * lx
                bot = bl + (br - bl) * lx
                out[n, y, x] = top + (bot - top) * ly
    return out


def pyramid_roi_align_literal(boxes, feature_maps, pool_shape, image_shape):
    """Literal PyramidROIAlign.call: 4 per-level crops, concat, re-sort by
    ``batch*100000 + box`` (modified_dense_model.py:343-416).  Returns
    ``(pooled [1, B*N, ph, pw, C], levels [B, N])``."""
    boxes = np.asarray(boxes, dtype=F32)
    assert boxes.ndim == 3 and boxes.shape[2] == 4
    roi_level = fpn_level(boxes, image_shape)
    pooled, box_to_level = [], []
    for i, level in enumerate(range(2, 6)):
        ix = np.argwhere(roi_level == level)               # tf.where: row-major order
        level_boxes = boxes[ix[:, 0], ix[:, 1]]
        box_indices = ix[:, 0].astype(np.int32)
        box_to_level.append(ix)
        pooled.append(crop_and_resize(feature_maps[i], level_boxes, box_indices, pool_shape))
    pooled = np.concatenate(pooled, axis=0)
    box_to_level = np.concatenate(box_to_level, axis=0)
    box_range = np.arange(box_to_level.shape[0])[:, None]
    box_to_level = np.concatenate([box_to_level.astype(np.int64), box_range], axis=1)
    sorting_tensor = box_to_level[:, 0] * 100000 + box_to_level[:, 1]
    # tf.nn.top_k(k=all).indices[::-1]  == ascending order of the (unique) keys
    ix = np.argsort(-sorting_tensor, kind="stable")[::-1]
    ix = box_to_level[:, 2][ix]
    pooled = pooled[ix]
    return pooled[None], roi_level


def pyramid_roi_align(boxes, feature_maps, pool_shape, image_shape):
    """Direct (single pass, original order) form; must equal the literal form bit for bit
    whenever num_boxes <= 100000 (the reference's sort key limit, :408)."""
    boxes = np.asarray(boxes, dtype=F32)
    B, N = boxes.shape[:2]
    lv = fpn_level(boxes, image_shape)
    C = feature_maps[0].shape[-1]
    out = np.empty((B * N, pool_shape[0], pool_shape[1], C), dtype=F32)
    flat_boxes = boxes.reshape(-1, 4)
    flat_lv = lv.reshape(-1)
    flat_b = np.repeat(np.arange(B), N)
    for i, level in enumerate(range(2, 6)):
        sel = np.nonzero(flat_lv == level)[0]
        out[sel] = crop_and_resize(feature_maps[i], flat_boxes[sel], flat_b[sel], pool_shape)
    return out[None], lv


def unique_tap_pixels(boxes, fm_shapes, pool_shape, image_shape):
    """T_unique of SURVEY.md section 8(d): the number of DISTINCT (image, level, y, x) feature-map
    pixels read by any bilinear tap of any RoI in the call.  Used only to compute the
    algorithmic (compulsory) bytes of the roofline: 4*C*(ph*pw*R + T_unique)."""
    boxes = np.asarray(boxes, dtype=F32)
    B, N = boxes.shape[:2]
    lv = fpn_level(boxes, image_shape).reshape(-1)
    fb = boxes.reshape(-1, 4)
    img = np.repeat(np.arange(B), N)
    ph, pw = pool_shape
    total = 0
    for i, level in enumerate(range(2, 6)):
        sel = np.nonzero(lv == level)[0]
        if sel.size == 0:
            continue
        H, W = fm_shapes[i]
        bx = fb[sel]
        Hm1, Wm1 = F32(H - 1), F32(W - 1)
        hs = ((bx[:, 2] - bx[:, 0]) * Hm1) / F32(ph - 1)
        ws = ((bx[:, 3] - bx[:, 1]) * Wm1) / F32(pw - 1)
        in_y = ((bx[:, 0] * Hm1)[:, None] + np.arange(ph, dtype=F32)[None] * hs[:, None]).astype(F32)
        in_x = ((bx[:, 1] * Wm1)[:, None] + np.arange(pw, dtype=F32)[None] * ws[:, None]).astype(F32)
        y_ok = (in_y >= 0) & (in_y <= Hm1)
        x_ok = (in_x >= 0) & (in_x <= Wm1)
        ys = np.stack([np.floor(in_y), np.ceil(in_y)], -1).astype(np.int64)     # [n,ph,2]
        xs = np.stack([np.floor(in_x), np.ceil(in_x)], -1).astype(np.int64)     # [n,pw,2]
        ok = (y_ok[:, :, None, None, None] & x_ok[:, None, None, :, None])
        ok = np.broadcast_to(ok, (len(sel), ph, 2, pw, 2))
        yy = np.broadcast_to(ys[:, :, :, None, None], ok.shape)
        xx = np.broadcast_to(xs[:, None, None, :, :], ok.shape)
        ii = np.broadcast_to(img[sel][:, None, None, None, None], ok.shape)
        key = (ii[ok] * H + yy[ok]) * W + xx[ok]
        total += np.unique(key).size
    return int(total)


# ----------------------------------------------------------------------------------------
# backward: gradient of PyramidROIAlign w.r.t. the feature maps
# ----------------------------------------------------------------------------------------

def pyramid_roi_align_backward(boxes, grad_out, fm_shapes, pool_shape, image_shape, dtype=np.float64):
    """Restates TF's CropAndResizeGradImage for the layer's 4 per-level crops (the boxes get no gradient:
    tf.stop_gradient, modified_dense_model.py:379-380).  For every in-range sample
        dtop = (1 - y_lerp) * g,  dbottom = y_lerp * g
        d[top, left] += (1 - x_lerp) * dtop,  d[top, right] += x_lerp * dtop   (same for bottom)
    boxes [B,N,4]; grad_out [B*N, ph, pw, C] in (image, box) order; fm_shapes = [(H_l, W_l)] * 4.
    Returns 4 arrays [B, H_l, W_l, C] (accumulated in `dtype`; the CUDA kernel accumulates with fp32
    atomics, so parity is to ~1e-5 relative, not bit-exact)."""
    B, N = boxes.shape[:2]
    ph, pw = pool_shape
    C = grad_out.shape[-1]
    levels = fpn_level(boxes, image_shape)
    grads = [np.zeros((B, h, w, C), dtype) for h, w in fm_shapes]
    g = grad_out.reshape(B, N, ph, pw, C)
    for b in range(B):
        for n in range(N):
            li = int(levels[b, n]) - 2
            H, W = fm_shapes[li]
            y1, x1, y2, x2 = [F32(v) for v in boxes[b, n]]
            Hm1, Wm1 = F32(H - 1), F32(W - 1)
            hs = ((y2 - y1) * Hm1) / F32(ph - 1) if ph > 1 else F32(0)
            ws = ((x2 - x1) * Wm1) / F32(pw - 1) if pw > 1 else F32(0)
            for iy in range(ph):
                in_y = y1 * Hm1 + F32(iy) * hs if ph > 1 else F32(0.5) * (y1 + y2) * Hm1
                if not (in_y >= 0 and in_y <= Hm1):
                    continue
                t, bo = int(np.floor(in_y)), int(np.ceil(in_y))
                ly = F32(in_y - F32(np.floor(in_y)))
                for ix in range(pw):
                    in_x = x1 * Wm1 + F32(ix) * ws if pw > 1 else F32(0.5) * (x1 + x2) * Wm1
                    if not (in_x >= 0 and in_x <= Wm1):
                        continue
                    l, r = int(np.floor(in_x)), int(np.ceil(in_x))
                    lx = F32(in_x - F32(np.floor(in_x)))
                    gv = g[b, n, iy, ix].astype(dtype)
                    dtop, dbot = (1 - dtype(ly)) * gv, dtype(ly) * gv
                    grads[li][b, t, l] += (1 - dtype(lx)) * dtop
                    grads[li][b, t, r] += dtype(lx) * dtop
                    grads[li][b, bo, l] += (1 - dtype(lx)) * dbot
                    grads[li][b, bo, r] += dtype(lx) * dbot
    return grads
