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


class _Shape(tuple):
    def as_list(self):
        return list(self)


class T(np.ndarray):
    """ndarray with TF's coercion rule: Python scalars and plain ndarrays take the TENSOR's dtype (no promotion)."""
    __array_priority__ = 100

    @property
    def shape(self):
        return _Shape(np.ndarray.shape.__get__(self))

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
    tf.squeeze = lambda x, axis=None: tensor(np.squeeze(_plain(x), axis=tuple(axis) if isinstance(axis, list) else axis))
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


# ---------------------------------------------------------------------------------------------------
# Eager stand-ins for the Keras layers the v1 caption models are wired from (text_generation_model.py:130-232).
# A "model" is built by RUNNING the reference's builder function on real arrays: KL.Input hands back the array
# registered under its name in FEEDS, every layer computes at once, and KM.Model just keeps the outputs.  Layer
# numerics (Embedding lookup + mask, masked LSTM scan, Dense + activation) are the oracle's restatements, with the
# trainable weights looked up by Keras name in WEIGHTS; what is pinned is the WIRING the reference code does.
# ---------------------------------------------------------------------------------------------------
FEEDS, WEIGHTS = {}, {}


def _with_mask(x, mask):
    x = tensor(x)
    x._keras_mask = mask
    return x


def _mask_of(x):
    return getattr(x, "_keras_mask", None)


def make_keras(tf):
    from oracle import decoder as dec

    counters = collections.Counter()

    def auto_name(kind, name):
        if name is not None:
            return name
        counters[kind] += 1
        return "%s_%d" % (kind, counters[kind])                # Keras' automatic layer names

    def Input(batch_shape=None, shape=None, name=None, **kw):
        x = tensor(FEEDS[name] if name is not None else FEEDS["__unnamed__"].pop(0))
        assert batch_shape is None or list(x.shape) == list(batch_shape), (name, x.shape, batch_shape)
        assert shape is None or list(x.shape[1:]) == list(shape), (name, x.shape, shape)
        return x

    class Conv2D(object):
        """Only what the RoI head needs: 'valid' convolution whose window covers the whole input (7x7 on 7x7, 1x1 on 1x1)."""
        def __init__(self, filters, kernel_size, padding="valid", trainable=True, name=None):
            self.filters, self.kernel_size, self.name = filters, tuple(kernel_size), name
            assert padding == "valid"

        def __call__(self, x):
            x = _plain(x)
            k, b = WEIGHTS[self.name + "/kernel"], WEIGHTS[self.name + "/bias"]
            assert x.shape[1:3] == self.kernel_size == k.shape[:2] and k.shape[3] == self.filters, (x.shape, k.shape)
            return tensor((x.reshape(x.shape[0], -1) @ k.reshape(-1, k.shape[-1]) + b)[:, None, None, :])

    class BatchNormalization(object):
        def __init__(self, axis=-1, trainable=True, name=None):
            self.name = name
            assert axis in (3, -1)

        def call(self, inputs, training=None):
            assert training is False                             # the reference's BatchNorm subclass hardcodes it
            g, b, m, v = (WEIGHTS["%s/%s" % (self.name, n)] for n in ("gamma", "beta", "moving_mean", "moving_variance"))
            return tensor(dec.batchnorm_inference(_plain(inputs), g, b, m, v))

        def __call__(self, x):
            return self.call(x)

    class Activation(object):
        def __init__(self, kind):
            assert kind == "relu"

        def __call__(self, x):
            return tensor(np.maximum(_plain(x), 0))

    class Lambda(object):
        def __init__(self, fn, name=None, **kw):
            self.fn = fn

        def __call__(self, x):
            return self.fn(x)

    class Embedding(object):
        def __init__(self, input_dim, output_dim, weights=None, trainable=True, mask_zero=False, name=None):
            self.table, self.mask_zero = np.asarray(weights[0], np.float32), mask_zero
            assert self.table.shape == (input_dim, output_dim)

        def __call__(self, ids):
            ids = _plain(ids).astype(np.int32)
            return _with_mask(self.table[ids], (ids != 0) if self.mask_zero else None)

    class LSTM(object):
        def __init__(self, units, recurrent_dropout=0.0, return_sequences=False, name=None):
            self.units, self.seq, self.name = units, return_sequences, auto_name("lstm", name)

        def __call__(self, x):
            mask = _mask_of(x)
            x = _plain(x)
            if mask is None:
                mask = np.ones(x.shape[:2], bool)
            k, r, b = (WEIGHTS["%s/%s" % (self.name, n)] for n in ("kernel", "recurrent_kernel", "bias"))
            assert r.shape[0] == self.units
            out = dec.lstm_masked(x, mask, k, r, b, self.seq)
            return _with_mask(out, mask if self.seq else None)

    class Dense(object):
        def __init__(self, units, activation=None, name=None):
            self.units, self.activation, self.name = units, activation, name

        def __call__(self, x):
            k, b = WEIGHTS[self.name + "/kernel"], WEIGHTS[self.name + "/bias"]
            assert k.shape[1] == self.units
            z = _plain(x) @ k + b
            return tensor({"relu": lambda v: np.maximum(v, 0), "softmax": dec.softmax, None: lambda v: v}[self.activation](z))

    class Concatenate(object):
        def __init__(self, axis=-1, name=None):
            self.axis = axis

        def __call__(self, xs):
            masks = [m for m in (_mask_of(x) for x in xs) if m is not None]
            return _with_mask(np.concatenate([_plain(x) for x in xs], self.axis), masks[0] if masks else None)

    class RepeatVector(object):
        def __init__(self, n):
            self.n = n

        def __call__(self, x):
            return tensor(np.repeat(_plain(x)[:, None, :], self.n, 1))

    class TimeDistributed(object):
        def __init__(self, layer, name=None):
            self.layer = layer
            if getattr(layer, "name", "") is None:              # weights are matched through the wrapper's name
                layer.name = name

        def __call__(self, x):
            return tensor(np.stack([_plain(self.layer(tensor(_plain(x)[:, t]))) for t in range(x.shape[1])], 1))

    class Model(object):
        def __init__(self, inputs, outputs, name=None):
            self.inputs, self.outputs, self.name = inputs, outputs, name

        def __call__(self, x):                                  # a built model re-applied (TimeDistributed): see recallable()
            return self.recall(x)

    KL = types.SimpleNamespace(Input=Input, Lambda=Lambda, Embedding=Embedding, LSTM=LSTM, Dense=Dense, Concatenate=Concatenate,
                               RepeatVector=RepeatVector, TimeDistributed=TimeDistributed, Layer=Layer, Conv2D=Conv2D,
                               BatchNormalization=BatchNormalization, Activation=Activation, _counters=counters)
    KM = types.SimpleNamespace(Model=Model)
    K = types.SimpleNamespace(
        squeeze=lambda x, axis: tensor(np.squeeze(_plain(x), axis)),
        switch=lambda c, a, b: a if bool(c) else b,
        mean=lambda x: tensor(np.mean(_plain(x), dtype=_plain(x).dtype)),
        categorical_crossentropy=lambda target, output: tensor(
            -np.sum(_plain(target) * np.log(np.clip(_plain(output) / _plain(output).sum(-1, keepdims=True),
                                                    np.float32(1e-7), np.float32(1 - 1e-7))), -1)))
    tf.ones = lambda shape: tensor(np.ones([int(s) for s in shape], np.float32))
    tf.zeros = lambda shape: tensor(np.zeros([int(s) for s in shape], np.float32))
    tf.argmax = lambda x, axis=None: tensor(np.argmax(_plain(x), axis=axis).astype(np.int64))
    tf.reduce_sum = lambda x, axis=None: tensor(np.sum(_plain(x), axis=axis, dtype=_plain(x).dtype))
    tf.size = lambda x: np.size(x)
    tf.constant = lambda v: tensor(np.float32(v))
    return KL, KM, K
