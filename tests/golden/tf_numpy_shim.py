"""A numpy stand-in for the ~30 TensorFlow-1.x graph ops that the reference's PyramidROIAlign.call and
ProposalLayer.call use, so that those two methods can be EXECUTED from /root/reference (by
gen_golden_reference_numpy.py, in this container only) and their glue -- level formula, per-level dispatch and
re-sort, delta / clip / normalise order, dtype coercions, padding -- checked against the oracle's restatement.

The heavy primitives are the oracle's own restatements (crop_and_resize, non_max_suppression, top_k order, the
correctly rounded log / exp, x86 float->int conversion, std::min/max NaN rules): what gets pinned is everything
AROUND them.  Test infrastructure; never imported by the product."""
import collections
import types

import numpy as np

from oracle import proposals as opr
from oracle import roi_align as ora


class T(np.ndarray):
    """ndarray with TF's coercion rule: Python scalars and plain ndarrays take the TENSOR's dtype (no promotion)."""
    __array_priority__ = 100

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kw):
        dts = {i.dtype for i in inputs if isinstance(i, T)}
        assert len(dts) == 1, "TF would reject mixed tensor dtypes: %s" % dts
        dt = dts.pop()
        conv = [i.view(np.ndarray) if isinstance(i, T) else np.asarray(i, dtype=dt) for i in inputs]
        if out is not None:
            kw["out"] = tuple(o.view(np.ndarray) if isinstance(o, T) else o for o in out)
        with np.errstate(all="ignore"):
            res = getattr(ufunc, method)(*conv, **kw)
        if out is not None:
            return out[0] if len(out) == 1 else out
        return tensor(res)


def tensor(x, dtype=None):
    return np.asarray(x, dtype=dtype).view(T)


def _plain(x):
    return np.asarray(x).view(np.ndarray)


def _minmax(fn_float, fn_int):
    def op(a, b, name=None):
        ts = [v for v in (a, b) if isinstance(v, T)]
        dt = ts[0].dtype if ts else np.result_type(a, b)      # (0-d indexing hands back numpy scalars)
        a, b = np.asarray(_plain(a), dt), np.asarray(_plain(b), dt)
        return tensor(fn_float(a, b) if dt.kind == "f" else fn_int(a, b))
    return op


def _cast(x, dtype):
    x = _plain(x)
    if dtype == np.int32 and x.dtype.kind == "f":
        return tensor(ora._x86_f32_to_i32(x))
    return tensor(x.astype(dtype))


def _top_k(x, k, sorted=True, name=None):
    x = _plain(x)
    k = int(k)
    order = np.argsort(-x.astype(np.float64) if x.dtype.kind == "f" else -x.astype(np.int64), axis=-1, kind="stable")[..., :k]
    return collections.namedtuple("TopKV2", "values indices")(tensor(np.take_along_axis(x, order, -1)), tensor(order.astype(np.int32)))


def _nms(boxes, scores, max_output_size, iou_threshold=0.5, name=None):
    return tensor(opr.tf_non_max_suppression(_plain(boxes), _plain(scores), int(max_output_size), float(iou_threshold)))


def _crop_and_resize(image, boxes, box_ind, crop_size, method="bilinear", extrapolation_value=0, name=None):
    assert method == "bilinear"
    return tensor(ora.crop_and_resize(_plain(image), _plain(boxes), _plain(box_ind), tuple(crop_size), extrapolation_value))


def make_tf():
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.int32 = np.float32, np.int32
    tf.split = lambda v, n, axis=0, name=None: [tensor(p) for p in np.split(_plain(v), n, axis=axis)]
    tf.sqrt = lambda x: tensor(np.sqrt(_plain(np.asarray(x, np.float32)))) if not isinstance(x, T) else np.sqrt(x)
    tf.log = lambda x: tensor(ora._f32_log(np.asarray(_plain(x), np.float32)))
    tf.exp = lambda x: tensor(opr.exp_f32(_plain(x)))
    tf.cast = _cast
    tf.round = lambda x: tensor(np.round(_plain(x)))
    tf.minimum = _minmax(opr.std_min, np.minimum)
    tf.maximum = _minmax(opr.std_max, np.maximum)
    tf.squeeze = lambda x, axis=None: tensor(np.squeeze(_plain(x), axis=axis))
    tf.equal = lambda a, b: tensor(_plain(a) == b)
    tf.where = lambda c: tensor(np.argwhere(_plain(c)).astype(np.int64))
    tf.gather_nd = lambda p, ix: tensor(_plain(p)[tuple(_plain(ix).T)])
    tf.gather = lambda p, ix, name=None: tensor(_plain(p)[_plain(ix)])
    tf.stop_gradient = lambda x: x
    tf.concat = lambda vs, axis=0, name=None: tensor(np.concatenate([_plain(v) for v in vs], axis=axis))
    tf.stack = lambda vs, axis=0, name=None: tensor(np.stack([_plain(v) for v in vs], axis=axis))
    tf.expand_dims = lambda x, axis: tensor(np.expand_dims(_plain(x), axis))
    tf.range = lambda n: tensor(np.arange(int(n), dtype=np.int32))
    tf.shape = lambda x: tensor(np.array(np.shape(x), np.int32))
    tf.pad = lambda x, paddings: tensor(np.pad(_plain(x), [(int(a), int(b)) for a, b in paddings]))
    tf.nn = types.SimpleNamespace(top_k=_top_k)
    tf.image = types.SimpleNamespace(crop_and_resize=_crop_and_resize, non_max_suppression=_nms)
    return tf


class Layer(object):
    """keras.engine.Layer as far as the two layers use it."""
    def __init__(self, **kwargs):
        self.name = kwargs.get("name")

    def __call__(self, inputs):
        return self.call(inputs)
