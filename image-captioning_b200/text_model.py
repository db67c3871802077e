"""Drop-in for the reference's text-generation Keras models.

Mirrors /root/reference/dense_img_cap_separate_models/text_generation_model.py:
``DenseCapConfig`` (:23-49), ``build_lstm_model(features_input, config, units, mode)`` (:235-283)
and text_generation_model_v2.py ``build_model(features_shape, word_shape, config, units, inject)``
(:140-166).  The returned objects expose the part of the Keras ``Model`` surface the reference's
callers use -- ``predict``, ``get_weights`` / ``set_weights`` / ``load_weights`` / ``save_weights``,
``compile`` / ``train_on_batch`` / ``fit_generator``, ``summary`` -- and run entirely in the sm_100a
kernels behind the C ABI (include/dcap.h).  No TensorFlow, no CPU fallback.
"""
import ctypes
import queue
import threading

import numpy as np
import torch

from . import _lib

ARCH_V1, ARCH_V2_INJECT = 1, 2
DTYPE_F32, DTYPE_BF16 = 0, 1
FEATS_ROI_F32, FEATS_HEAD_F32, FEATS_ROI_BF16 = 0, 1, 2
_HEAD = ["mrcnn_class_conv1/kernel", "mrcnn_class_conv1/bias",
         "mrcnn_class_bn1/gamma", "mrcnn_class_bn1/beta", "mrcnn_class_bn1/moving_mean",
         "mrcnn_class_bn1/moving_variance",
         "mrcnn_class_conv2/kernel", "mrcnn_class_conv2/bias",
         "mrcnn_class_bn2/gamma", "mrcnn_class_bn2/beta", "mrcnn_class_bn2/moving_mean",
         "mrcnn_class_bn2/moving_variance"]
# Keras get_weights() order: per layer trainable then non-trainable; the caption layer lists the
# word model's trainable weights first and the frozen embedding last (text_generation_model.py:205-206).
V1_WEIGHT_ORDER = _HEAD + [
    "imgcap_lstm1/kernel", "imgcap_lstm1/recurrent_kernel", "imgcap_lstm1/bias",
    "imgcap_lstm2/kernel", "imgcap_lstm2/recurrent_kernel", "imgcap_lstm2/bias",
    "imgcap_lstm_d1/kernel", "imgcap_lstm_d1/bias", "imgcap_lstm_d2/kernel", "imgcap_lstm_d2/bias",
    "imgcap_embedding_layer/embeddings"]
V2_WEIGHT_ORDER = _HEAD + [
    "imgcap_embedding_layer/embeddings",
    "lstm_1/kernel", "lstm_1/recurrent_kernel", "lstm_1/bias",
    "imgcap_lstm/kernel", "imgcap_lstm/recurrent_kernel", "imgcap_lstm/bias",
    "imgcap_d1/kernel", "imgcap_d1/bias"]


def trainable_weight_names(arch):
    """Names behind Keras' ``model.trainable_weights`` (the training scripts print it: text_generation_model.py:465,
    text_generation_model_v2.py:305).  v1: everything but the BatchNorm moving statistics and the frozen embedding
    (text_generation_model.py:130-139, trainable=False).  v2 inject: the RoI head is frozen as well
    (text_generation_model_v2.py:141-152), so only the two LSTMs and the output layer train."""
    if arch == ARCH_V1:
        return [n for n in V1_WEIGHT_ORDER if "/moving_" not in n and not n.endswith("/embeddings")]
    return [n for n in V2_WEIGHT_ORDER if n not in _HEAD and not n.endswith("/embeddings")]


class DenseCapConfig(object):
    """The fields of the reference's DenseCapConfig that the text models read
    (text_generation_model.py:23-49)."""
    NAME = "dense image captioning"
    GPU_COUNT = 1
    IMAGES_PER_GPU = 1
    BATCH_SIZE = 10
    STEPS_PER_EPOCH = 500
    VALIDATION_STEPS = 50
    PADDING_SIZE = 10
    POOL_SIZE = 7

    def __init__(self, vocab_size, embedding_weights, batch_size=None, padding_size=None):
        self.VOCABULARY_SIZE = int(vocab_size)
        self.EMBEDDING_WEIGHTS = np.asarray(embedding_weights, dtype=np.float32)
        self.EMBEDDING_SIZE = int(self.EMBEDDING_WEIGHTS.shape[1])
        if batch_size is not None:
            self.BATCH_SIZE = int(batch_size)
        if padding_size is not None:
            self.PADDING_SIZE = int(padding_size)

    def display(self):
        """Config.display() as the training scripts call it (config.py:166-172): one line per public attribute.  The
        embedding matrix is shown by shape, not dumped."""
        print("\nConfigurations:")
        for a in dir(self):
            if not a.startswith("__") and not callable(getattr(self, a)):
                v = getattr(self, a)
                print("{:30} {}".format(a, "array%s" % (tuple(v.shape),) if isinstance(v, np.ndarray) else v))
        print("\n")


class Adam(object):
    """keras.optimizers.Adam as the reference constructs it (text_generation_model.py:425,
    text_generation_model_v2.py:266): same arguments and defaults as Keras 2.1 (epsilon=None means
    K.epsilon() = 1e-7; `decay` rescales lr by 1/(1 + decay*iterations))."""

    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=None, decay=0.0, amsgrad=False):
        self.lr, self.beta_1, self.beta_2 = float(lr), float(beta_1), float(beta_2)
        self.epsilon = 1e-7 if epsilon is None else float(epsilon)
        self.decay, self.amsgrad = float(decay), bool(amsgrad)
        self.iterations = 0

    def current_lr(self):
        return self.lr / (1.0 + self.decay * self.iterations) if self.decay > 0 else self.lr


def roi_caption_loss(y_true=None, y_pred=None):
    """Marker for compile(loss=roi_caption_loss) (text_generation_model.py:286-294).  The loss is
    evaluated inside the fused training step (dc_decoder_train_step); it is not a host function."""
    raise NotImplementedError("roi_caption_loss is fused into the training step; use train_on_batch / evaluate")


class GeneratorQueue(object):
    """What Keras puts behind ``fit_generator(generator, max_queue_size=, workers=)`` (GeneratorEnqueuer; the reference
    trains with max_queue_size=100 and the default single worker, text_generation_model.py:470-472): a background thread
    that keeps up to ``max_queue_size`` batches of the generator ready while the previous batch trains.  One worker thread,
    so batches arrive in generator order; ``workers=0`` pulls from the generator on the calling thread, as Keras does.
    The generator's numpy work overlaps the training step because the step's host side is mostly a wait on the device
    (which releases the GIL).  Exceptions raised by the generator -- StopIteration included -- surface from ``get()``."""

    _END = object()

    def __init__(self, generator, max_queue_size=10, workers=1):
        self._gen = generator
        self._thread = None
        if workers and workers > 0:
            self._q = queue.Queue(maxsize=max(1, int(max_queue_size)))
            self._stop = threading.Event()
            self._thread = threading.Thread(target=self._fill, name="dcap-generator-queue", daemon=True)
            self._thread.start()

    def _fill(self):
        try:
            while not self._stop.is_set():
                try:
                    item = (True, next(self._gen))
                except BaseException as e:                  # StopIteration too: hand it to the consumer and end
                    item = (False, e)
                while not self._stop.is_set():
                    try:
                        self._q.put(item, timeout=0.05)
                        break
                    except queue.Full:
                        continue
                if not item[0]:
                    return
        finally:
            pass

    def get(self):
        if self._thread is None:
            return next(self._gen)
        ok, item = self._q.get()
        if not ok:
            self._q.put((False, item))                      # every later get() raises again
            raise item
        return item

    def close(self):
        """Stop the worker (the batches it had queued are dropped -- Keras' enqueuer.stop())."""
        if self._thread is not None:
            self._stop.set()
            try:
                while True:
                    self._q.get_nowait()
            except queue.Empty:
                pass
            self._thread.join(timeout=5.0)
            self._thread = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class History(object):
    """What fit_generator returns (keras.callbacks.History): .history = {'loss': [...], 'val_loss': [...]}."""

    def __init__(self):
        self.epoch, self.history = [], {}


class _DeviceArray(object):
    """A raw device range exposed through __cuda_array_interface__ (zero-copy torch view)."""

    def __init__(self, ptr, numel, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": (int(numel),), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2}


class _DcTrainOptions(ctypes.Structure):
    """include/dcap.h: DcTrainOptions"""
    _fields_ = [("d_feats", ctypes.c_void_p), ("recurrent_dropout", ctypes.c_float), ("dropout_seed", ctypes.c_uint64),
                ("dropout_step", ctypes.c_int64), ("row_offset", ctypes.c_int64)]


class _DcDecoderConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("arch", "dtype", "vocab", "embed", "feat", "units",
                                            "word_units", "pool", "channels", "padding")]


def _dtype_code(dtype):
    if dtype in ("float32", "fp32", torch.float32, np.float32, DTYPE_F32):
        return DTYPE_F32
    if dtype in ("bfloat16", "bf16", torch.bfloat16, DTYPE_BF16):
        return DTYPE_BF16
    raise ValueError("dtype must be 'float32' or 'bfloat16'")


class _ModelBase(object):
    """Handle owner + the weight half of the Keras Model surface."""

    def __init__(self, arch, config, units, features_input, dtype, word_units=0, device=None):
        if len(features_input) != 3 or features_input[0] != features_input[1]:
            raise ValueError("features_input must be [pool, pool, channels]")
        if features_input[0] != config.POOL_SIZE:
            raise ValueError("features_input pool %d != config.POOL_SIZE %d"
                             % (features_input[0], config.POOL_SIZE))
        self.config = config
        self.units = int(units)
        self.arch = arch
        self.dtype = _dtype_code(dtype)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        self._lib = _lib.load()
        cfg = _DcDecoderConfig(arch, self.dtype, config.VOCABULARY_SIZE, config.EMBEDDING_SIZE, 1024,
                               self.units, int(word_units), int(features_input[0]),
                               int(features_input[2]), config.PADDING_SIZE)
        self._cfg = cfg
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.dc_decoder_create(ctypes.byref(cfg), ctypes.byref(self._h)))
        self._dirty = True
        self._names = [self._lib.dc_decoder_weight_name(self._h, i).decode()
                       for i in range(self._lib.dc_decoder_weight_count(self._h))]
        self._numel = {n: self._lib.dc_decoder_weight_numel(self._h, i) for i, n in enumerate(self._names)}
        self._shapes = self._weight_shapes()
        self._set = set()
        # like the Keras models, the embedding layer is initialised from config.EMBEDDING_WEIGHTS
        self._set_one("imgcap_embedding_layer/embeddings", config.EMBEDDING_WEIGHTS)
        self.optimizer = None
        self.loss = None

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                self._lib.dc_decoder_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass

    # ---- weights ----
    def _weight_shapes(self):
        c, u = self._cfg, self.units
        F, E, V, p, C = c.feat, c.embed, c.vocab, c.pool, c.channels
        s = {"mrcnn_class_conv1/kernel": (p, p, C, F), "mrcnn_class_conv1/bias": (F,),
             "mrcnn_class_conv2/kernel": (1, 1, F, F), "mrcnn_class_conv2/bias": (F,),
             "imgcap_embedding_layer/embeddings": (V, E)}
        for bn in ("mrcnn_class_bn1", "mrcnn_class_bn2"):
            for n in ("gamma", "beta", "moving_mean", "moving_variance"):
                s["%s/%s" % (bn, n)] = (F,)
        if self.arch == ARCH_V1:
            s.update({"imgcap_lstm1/kernel": (E + F, 4 * u), "imgcap_lstm1/recurrent_kernel": (u, 4 * u),
                      "imgcap_lstm1/bias": (4 * u,), "imgcap_lstm2/kernel": (u, 4 * u),
                      "imgcap_lstm2/recurrent_kernel": (u, 4 * u), "imgcap_lstm2/bias": (4 * u,),
                      "imgcap_lstm_d1/kernel": (u + F, 1024), "imgcap_lstm_d1/bias": (1024,),
                      "imgcap_lstm_d2/kernel": (1024, V), "imgcap_lstm_d2/bias": (V,)})
        else:
            wu = c.word_units
            s.update({"lstm_1/kernel": (E, 4 * wu), "lstm_1/recurrent_kernel": (wu, 4 * wu),
                      "lstm_1/bias": (4 * wu,), "imgcap_lstm/kernel": (F + wu, 4 * u),
                      "imgcap_lstm/recurrent_kernel": (u, 4 * u), "imgcap_lstm/bias": (4 * u,),
                      "imgcap_d1/kernel": (u, V), "imgcap_d1/bias": (V,)})
        assert sorted(s) == sorted(self._names)
        return s

    @property
    def weight_names(self):
        return list(V1_WEIGHT_ORDER if self.arch == ARCH_V1 else V2_WEIGHT_ORDER)

    def _set_one(self, name, value):
        if name not in self._shapes:
            raise ValueError("unknown weight %r" % name)
        a = np.ascontiguousarray(np.asarray(value, dtype=np.float32))
        if tuple(a.shape) != tuple(self._shapes[name]):
            raise ValueError("weight %r: expected shape %s, got %s" % (name, self._shapes[name], a.shape))
        with torch.cuda.device(self.device):
            _lib.check(self._lib.dc_decoder_set_weight(self._h, name.encode(), ctypes.c_void_p(a.ctypes.data),
                                                       ctypes.c_int64(a.size)))
        self._set.add(name)
        self._dirty = True

    def set_weights(self, weights):
        """Keras-ordered list (see ``weight_names``) or a ``{name: array}`` dict."""
        if isinstance(weights, dict):
            for n, v in weights.items():
                self._set_one(n, v)
            return
        names = self.weight_names
        if len(weights) != len(names):
            raise ValueError("expected %d weight arrays, got %d" % (len(names), len(weights)))
        for n, v in zip(names, weights):
            self._set_one(n, v)

    def _get_one(self, name):
        out = np.empty(self._shapes[name], np.float32)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.dc_decoder_get_weight(self._h, name.encode(), ctypes.c_void_p(out.ctypes.data),
                                                       ctypes.c_int64(out.size)))
        return out

    def get_weights(self):
        return [self._get_one(n) for n in self.weight_names]

    def get_weights_dict(self):
        return {n: self._get_one(n) for n in self.weight_names}

    def save_weights(self, path, overwrite=True):
        """Weights as an .npz archive keyed by Keras weight name, written exactly at ``path`` (the reference names its
        checkpoints *.h5; h5py is not available here, and load_weights tells the formats apart by content)."""
        from . import data
        data.write_weight_file(path, self.get_weights_dict())

    def save(self, path, overwrite=True, include_optimizer=False):
        """keras Model.save as ModelCheckpoint(save_weights_only=False) calls it: these models have no state besides
        their weights (the optimiser slots are not written), so this is save_weights."""
        self.save_weights(path)

    def load_weights(self, path, by_name=False, skip_mismatch=False):
        """by_name=True loads only the tensors present in the file (as the reference does with
        mask_rcnn_coco.h5 to fill the head, text_generation_model.py:468)."""
        from . import data
        found = data.read_weight_file(path)                  # .npz archive, or Keras HDF5 (needs h5py) -- by content
        for n, v in found.items():
            if n not in self._shapes:
                if by_name:
                    continue
                raise ValueError("file holds unknown weight %r" % n)
            if tuple(v.shape) != tuple(self._shapes[n]):
                if skip_mismatch:
                    continue
                raise ValueError("weight %r: expected shape %s, got %s" % (n, self._shapes[n], v.shape))
            self._set_one(n, v)
        if not by_name:
            missing = [n for n in self._shapes if n not in found]
            if missing:
                raise ValueError("file lacks weights: %s" % ", ".join(missing))

    def _ready(self):
        if self._dirty:
            missing = [n for n in self._names if n not in self._set]
            if missing:
                raise RuntimeError("weights not set: %s" % ", ".join(missing))
            with torch.cuda.device(self.device):
                _lib.check(self._lib.dc_decoder_finalize(self._h, self._stream()))
            self._dirty = False

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def trainable_weights(self):
        """Keras attribute; here the NAMES of the trainable tensors (get_weights_dict()[name] is the value)."""
        return trainable_weight_names(self.arch)

    @property
    def non_trainable_weights(self):
        t = set(self.trainable_weights)
        return [n for n in self.weight_names if n not in t]

    @property
    def weights(self):
        return self.weight_names

    def count_params(self):
        return int(sum(int(np.prod(self._shapes[n])) for n in self.weight_names))

    def summary(self):
        total, trainable = 0, set(self.trainable_weights)
        print("%-40s %-24s %12s" % ("weight", "shape", "params"))
        for n in self.weight_names:
            k = int(np.prod(self._shapes[n]))
            total += k
            print("%-40s %-24s %12d" % (n, self._shapes[n], k))
        n_train = int(sum(int(np.prod(self._shapes[n])) for n in trainable))
        print("Total params: %d" % total)
        print("Trainable params: %d" % n_train)
        print("Non-trainable params: %d" % (total - n_train))

    # ---- feature plumbing ----
    def _feats_to_device(self, x):
        """Accepts [N,p,p,C] RoI features or [N,1024] head features (numpy or torch; fp32, or bf16
        RoI features on the device).  Returns (tensor, feats_kind, was_numpy)."""
        c = self._cfg
        was_numpy = not isinstance(x, torch.Tensor)
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)) if was_numpy else x
        if t.dim() == 4 and tuple(t.shape[1:]) == (c.pool, c.pool, c.channels):
            kind = FEATS_ROI_BF16 if t.dtype == torch.bfloat16 else FEATS_ROI_F32
        elif t.dim() == 2 and t.shape[1] == c.feat:
            kind = FEATS_HEAD_F32
        else:
            raise ValueError("features must be [N,%d,%d,%d] or [N,%d], got %s"
                             % (c.pool, c.pool, c.channels, c.feat, tuple(t.shape)))
        if kind != FEATS_ROI_BF16:
            t = t.to(torch.float32)
        return t.to(self.device).contiguous(), kind, was_numpy


class RoiCaptionModel(_ModelBase):
    """build_lstm_model(...) result: v1 "inject at every step" model."""

    def __init__(self, features_input, config, units, mode, dtype="float32", device=None):
        assert mode in ["training", "inference"]
        self.mode = mode
        super().__init__(ARCH_V1, config, units, list(features_input), dtype, device=device)

    def generate(self, features, return_probs=False, chunk=None, return_scores=False):
        """Greedy captions: token ids [N,P] (int32) and optionally the [N,P,V] probabilities or the
        caption scores [N] = sum_t log max_v p (the score refine_generations ranks by,
        evaluate_models/test_score_dense_captions.py:256-258).
        torch CUDA in -> torch CUDA out; numpy in -> numpy out."""
        self._ready()
        t, kind, was_numpy = self._feats_to_device(features)
        N, P, V = t.shape[0], self.config.PADDING_SIZE, self.config.VOCABULARY_SIZE
        tokens = torch.empty((N, P), dtype=torch.int32, device=self.device)
        probs = torch.empty((N, P, V), dtype=torch.float32, device=self.device) if return_probs else None
        scores = torch.empty((N,), dtype=torch.float32, device=self.device) if return_scores else None
        chunk = N if not chunk else int(chunk)
        with torch.cuda.device(self.device):
            for i in range(0, N, max(chunk, 1)):
                j = min(N, i + chunk)
                if scores is not None and probs is None:
                    _lib.check(self._lib.dc_decoder_greedy_scored(
                        self._h, ctypes.c_void_p(t[i:j].data_ptr()), kind, j - i,
                        ctypes.c_void_p(tokens[i:j].data_ptr()), ctypes.c_void_p(scores[i:j].data_ptr()), self._stream()))
                    continue
                _lib.check(self._lib.dc_decoder_greedy(
                    self._h, ctypes.c_void_p(t[i:j].data_ptr()), kind, j - i,
                    ctypes.c_void_p(tokens[i:j].data_ptr()),
                    ctypes.c_void_p(probs[i:j].data_ptr()) if probs is not None else None, self._stream()))
        if scores is not None and probs is not None:
            scores = torch.log(probs.max(-1).values).sum(-1)
        outs = [tokens] + ([probs] if return_probs else []) + ([scores] if return_scores else [])
        if was_numpy:
            outs = [o.cpu().numpy() for o in outs]
        return outs[0] if len(outs) == 1 else tuple(outs)

    def predict(self, x, batch_size=None, verbose=0):
        """Keras predict of the inference model: [N,P,V] word probabilities
        (evaluate_models/eval_text_generation_model.py:141).  As in Keras the sample count must be
        a multiple of config.BATCH_SIZE (the graph bakes the batch in, :211)."""
        if self.mode != "inference":
            # the training graph's inputs are [features, gt_captions] (text_generation_model.py:264-277)
            if not isinstance(x, (list, tuple)) or len(x) != 2:
                raise ValueError("the training graph predicts from [features, gt_captions]")
            if len(x[0]) % self.config.BATCH_SIZE != 0:
                raise ValueError("number of samples %d is not a multiple of config.BATCH_SIZE %d"
                                 % (len(x[0]), self.config.BATCH_SIZE))
            return self.predict_teacher_forced(x)
        n = len(x)
        if n % self.config.BATCH_SIZE != 0:
            raise ValueError("number of samples %d is not a multiple of config.BATCH_SIZE %d"
                             % (n, self.config.BATCH_SIZE))
        _, probs = self.generate(x, return_probs=True)
        return probs

    def beam_search(self, features, beam_width=3, chunk=None):
        """gen_captions semantics (image captioning/test.py:23-64) on the v1 decoder: returns
        (tokens [N,k,P] ascending by score -- best beam last --, scores [N,k] float64)."""
        self._ready()
        t, kind, was_numpy = self._feats_to_device(features)
        N, P, k = t.shape[0], self.config.PADDING_SIZE, int(beam_width)
        tokens = torch.empty((N, k, P), dtype=torch.int32, device=self.device)
        scores = torch.empty((N, k), dtype=torch.float64, device=self.device)
        chunk = N if not chunk else int(chunk)
        with torch.cuda.device(self.device):
            for i in range(0, N, max(chunk, 1)):
                j = min(N, i + chunk)
                _lib.check(self._lib.dc_decoder_beam(
                    self._h, ctypes.c_void_p(t[i:j].data_ptr()), kind, j - i, k,
                    ctypes.c_void_p(tokens[i:j].data_ptr()), ctypes.c_void_p(scores[i:j].data_ptr()),
                    self._stream()))
        if was_numpy:
            return tokens.cpu().numpy(), scores.cpu().numpy()
        return tokens, scores

    def caption_rois(self, boxes, feature_maps, image_shape, wait=True):
        """generate_features + predict in one call: ROIAlign -> head -> greedy ids [B*N, P].
        torch CUDA inputs run on the device (dc_caption_rois); numpy inputs go through the
        host-buffer pipeline (dc_caption_rois_host).  ``wait=False`` (host inputs): only enqueue the call
        (dc_caption_rois_host_submit) and return the token array it WILL fill; ``caption_rois_wait()`` blocks until the
        oldest outstanding call is complete.  Up to two calls may be outstanding: the upload of the second overlaps the
        decode tail of the first.  Keep the (pinned) inputs alive and unmodified until the wait."""
        self._ready()
        P = self.config.PADDING_SIZE
        on_dev = all(isinstance(t, torch.Tensor) and t.is_cuda for t in [boxes] + list(feature_maps))
        if on_dev:
            b = boxes.detach().to(torch.float32).contiguous()
            fms = [f.detach().to(torch.float32).contiguous() for f in feature_maps]
        else:
            to_np = lambda a: a.numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
            b = np.ascontiguousarray(to_np(boxes), dtype=np.float32)
            fms = [np.ascontiguousarray(to_np(f), dtype=np.float32) for f in feature_maps]
        if b.ndim != 3 or b.shape[2] != 4 or len(fms) != 4:
            raise ValueError("expected boxes [B,N,4] and 4 feature maps")
        for f in fms:
            if f.ndim != 4 or f.shape[0] != b.shape[0] or f.shape[3] != self._cfg.channels:
                raise ValueError("feature maps must be [B,h,w,%d]" % self._cfg.channels)
        B, N = b.shape[:2]
        hs = (ctypes.c_int * 4)(*[f.shape[1] for f in fms])
        ws = (ctypes.c_int * 4)(*[f.shape[2] for f in fms])
        if on_dev:
            tokens = torch.empty((B * N, P), dtype=torch.int32, device=self.device)
            ptrs = (ctypes.c_void_p * 4)(*[f.data_ptr() for f in fms])
            with torch.cuda.device(self.device):
                _lib.check(self._lib.dc_caption_rois(self._h, ctypes.c_void_p(b.data_ptr()), ptrs, hs, ws, B, N,
                                                     int(image_shape[0]), int(image_shape[1]),
                                                     ctypes.c_void_p(tokens.data_ptr()), self._stream()))
            return tokens
        tokens = np.empty((B * N, P), np.int32)
        ptrs = (ctypes.c_void_p * 4)(*[f.ctypes.data for f in fms])
        fn = self._lib.dc_caption_rois_host if wait else self._lib.dc_caption_rois_host_submit
        with torch.cuda.device(self.device):
            _lib.check(fn(self._h, ctypes.c_void_p(b.ctypes.data), ptrs, hs, ws, B, N, int(image_shape[0]),
                          int(image_shape[1]), ctypes.c_void_p(tokens.ctypes.data)))
        if not wait:
            self._inflight = getattr(self, "_inflight", []) + [(b, fms, tokens)]      # keep the host buffers alive
        return tokens

    def caption_rois_wait(self):
        """Block until the oldest caption_rois(..., wait=False) call has delivered its tokens."""
        with torch.cuda.device(self.device):
            _lib.check(self._lib.dc_caption_rois_host_wait(self._h))
        if getattr(self, "_inflight", None):
            self._inflight.pop(0)

    # ---- training surface (text_generation_model.py:424-426, 470-472) ----
    def compile(self, optimizer="adam", loss=None, metrics=None, **kwargs):
        """model.compile(optimizer=Adam(amsgrad=True), loss=roi_caption_loss)."""
        if isinstance(optimizer, str):
            if optimizer.lower() != "adam":
                raise ValueError("only the Adam optimiser of the reference is implemented")
            optimizer = Adam()
        if not isinstance(optimizer, Adam):
            raise ValueError("optimizer must be image_captioning_b200.Adam (mirror of keras.optimizers.Adam)")
        ok = loss is None or loss is roi_caption_loss or loss in ("categorical_crossentropy", "roi_caption_loss") \
            or getattr(loss, "__name__", "") in ("roi_caption_loss", "categorical_crossentropy")
        if not ok:
            raise ValueError("loss must be roi_caption_loss / categorical_crossentropy (fused in the training step)")
        if self.dtype != DTYPE_BF16:
            raise ValueError("the training step runs on the bf16 tensor-core path: build the model with dtype='bfloat16'")
        self.optimizer, self.loss = optimizer, loss

    def _ids_to_device(self, a, name):
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
        if t.dim() != 2 or t.shape[1] != self.config.PADDING_SIZE:
            raise ValueError("%s must be [N, PADDING_SIZE=%d], got %s" % (name, self.config.PADDING_SIZE, tuple(t.shape)))
        return t.to(self.device).to(torch.int32).contiguous()

    def _targets_to_device(self, y):
        """One-hot [N,P,V] (what the reference's generator yields, :354-358; an all-zero row means
        'no target', :287) or class ids [N,P] (negative = no target)."""
        if y is None:
            return None
        t = y if isinstance(y, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(y))
        if t.dim() == 3:
            t = t.to(self.device)
            ids = t.argmax(-1).to(torch.int32)
            ids = torch.where(t.sum(-1) > 0, ids, torch.full_like(ids, -1))
            return ids.contiguous()
        return self._ids_to_device(t, "targets")

    def grad_buffer(self):
        """Zero-copy torch view of the flat fp32 gradient buffer over all trainable tensors (what a
        data-parallel host all-reduces between train_step_device and apply_gradients)."""
        p, n = ctypes.c_void_p(), ctypes.c_int64()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.dc_decoder_grad_buffer(self._h, ctypes.byref(p), ctypes.byref(n)))
            return torch.as_tensor(_DeviceArray(p.value, n.value, self), device=self.device)

    def grad_buckets(self):
        """[(offset, numel)] of the gradient buckets in the order the backward pass completes them."""
        out = []
        for i in range(4):
            o, n = ctypes.c_int64(), ctypes.c_int64()
            _lib.check(self._lib.dc_decoder_grad_bucket(self._h, i, ctypes.byref(o), ctypes.byref(n)))
            out.append((o.value, n.value))
        return out

    def wait_grad_bucket(self, index, stream):
        """Make torch stream `stream` wait until bucket `index` of the last training step is complete."""
        with torch.cuda.device(self.device):
            _lib.check(self._lib.dc_decoder_wait_grad_bucket(self._h, int(index), ctypes.c_void_p(stream.cuda_stream)))

    def param_buffer(self):
        p, n = ctypes.c_void_p(), ctypes.c_int64()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.dc_decoder_param_buffer(self._h, ctypes.byref(p), ctypes.byref(n)))
            return torch.as_tensor(_DeviceArray(p.value, n.value, self), device=self.device)

    def get_gradients(self):
        """{weight name: gradient array} of the last training step (trainable tensors, Keras layout)."""
        out = {}
        for n in self.weight_names:
            if "/moving_" in n or n.endswith("/embeddings"):
                continue
            g = np.empty(self._shapes[n], np.float32)
            with torch.cuda.device(self.device):
                _lib.check(self._lib.dc_decoder_get_grad(self._h, n.encode(), ctypes.c_void_p(g.ctypes.data),
                                                         ctypes.c_int64(g.size)))
            out[n] = g
        return out

    def train_step_device(self, features, gt_captions, targets=None, inv_count=0.0, loss_out=None, d_features=None,
                          recurrent_dropout=0.0, dropout_seed=0, dropout_step=0, row_offset=0):
        """Forward + loss + backward on device tensors; gradients stay in grad_buffer().  Returns the
        device scalar  sum_positions(-log p_y) * inv_count  (inv_count <= 0: 1/(N*P)).  No host sync.
        ``d_features`` (fp32 CUDA tensor shaped like the RoI features) receives dL/d(features): the gradient the
        joint model chains into PyramidROIAlign's backward (dense_img_cap/dense_model.py:738-755).
        ``recurrent_dropout`` > 0 applies KL.LSTM(recurrent_dropout=...) masks (text_generation_model.py:141-142),
        drawn from (dropout_seed, dropout_step, row_offset + row): see include/dcap.h DcTrainOptions."""
        self._ready()
        t, kind, _ = self._feats_to_device(features)
        gt = self._ids_to_device(gt_captions, "gt_captions")
        tg = self._targets_to_device(targets)
        if gt.shape[0] != t.shape[0] or (tg is not None and tg.shape != gt.shape):
            raise ValueError("features, gt_captions and targets disagree on the batch size")
        loss = torch.empty((), dtype=torch.float32, device=self.device) if loss_out is None else loss_out
        opts = None
        if d_features is not None or recurrent_dropout:
            if d_features is not None and (not d_features.is_cuda or d_features.dtype != torch.float32
                                           or not d_features.is_contiguous() or d_features.numel() != t.numel()):
                raise ValueError("d_features must be a contiguous fp32 CUDA tensor shaped like the RoI features")
            opts = _DcTrainOptions(d_features.data_ptr() if d_features is not None else None, float(recurrent_dropout),
                                   int(dropout_seed), int(dropout_step), int(row_offset))
        with torch.cuda.device(self.device):
            _lib.check(self._lib.dc_decoder_train_step_ex(
                self._h, ctypes.c_void_p(t.data_ptr()), kind, t.shape[0], ctypes.c_void_p(gt.data_ptr()),
                ctypes.c_void_p(tg.data_ptr()) if tg is not None else None, ctypes.c_float(inv_count),
                ctypes.c_void_p(loss.data_ptr()), ctypes.byref(opts) if opts is not None else None, self._stream()))
        return loss

    def apply_gradients(self, grad_scale=1.0):
        """One optimiser update from grad_buffer() (Keras Adam / AMSGrad formula)."""
        if self.optimizer is None:
            raise RuntimeError("compile() the model first")
        o = self.optimizer
        lr = o.current_lr()
        o.iterations += 1
        with torch.cuda.device(self.device):
            _lib.check(self._lib.dc_adam_step(self._h, ctypes.c_float(lr), ctypes.c_float(o.beta_1),
                                              ctypes.c_float(o.beta_2), ctypes.c_float(o.epsilon), int(o.amsgrad),
                                              ctypes.c_int64(o.iterations), ctypes.c_float(grad_scale), self._stream()))

    def apply_gradients_ranges(self, ranges, grad_scale=1.0):
        """ONE optimiser update restricted to the ``(offset, numel)`` ranges of the flat buffers (sharded optimiser:
        this rank's ranges after the gradient reduce-scatter).  Advances the iteration count once; call
        ``params_updated()`` when the ranks have exchanged their ranges."""
        if self.optimizer is None:
            raise RuntimeError("compile() the model first")
        o = self.optimizer
        lr = o.current_lr()
        o.iterations += 1
        with torch.cuda.device(self.device):
            for offset, numel in ranges:
                if numel <= 0:
                    continue
                _lib.check(self._lib.dc_adam_step_range(self._h, ctypes.c_float(lr), ctypes.c_float(o.beta_1),
                                                        ctypes.c_float(o.beta_2), ctypes.c_float(o.epsilon), int(o.amsgrad),
                                                        ctypes.c_int64(o.iterations), ctypes.c_float(grad_scale),
                                                        ctypes.c_int64(int(offset)), ctypes.c_int64(int(numel)), self._stream()))

    def params_updated(self):
        """The flat parameter buffer was written from outside (all-gather): refresh every derived copy."""
        with torch.cuda.device(self.device):
            _lib.check(self._lib.dc_decoder_params_updated(self._h, self._stream()))

    def train_on_batch(self, x, y=None, sample_weight=None, class_weight=None):
        """Keras train_on_batch([features, gt_captions], one_hot_targets) -> scalar loss."""
        if self.optimizer is None:
            raise RuntimeError("compile() the model first")
        if sample_weight is not None or class_weight is not None:
            raise NotImplementedError("sample/class weights are not used by the reference")
        feats, gt = x
        tg = self._targets_to_device(y)
        inv = 0.0
        if tg is not None:
            cnt = int((tg >= 0).sum().item())                       # roi_caption_loss averages over rows with a target
            if cnt == 0:
                return 0.0
            inv = 1.0 / cnt
        loss = self.train_step_device(feats, gt, tg, inv)
        self.apply_gradients()
        return float(loss.item())

    def test_on_batch(self, x, y=None, sample_weight=None):
        feats, gt = x
        tg = self._targets_to_device(y)
        inv = 0.0
        if tg is not None:
            cnt = int((tg >= 0).sum().item())
            if cnt == 0:
                return 0.0
            inv = 1.0 / cnt
        return float(self.train_step_device(feats, gt, tg, inv).item())

    def predict_teacher_forced(self, x):
        """predict([features, gt_captions]) of the training graph: [N,P,V] word probabilities."""
        self._ready()
        feats, gt = x
        t, kind, was_numpy = self._feats_to_device(feats)
        g = self._ids_to_device(gt, "gt_captions")
        probs = torch.empty((t.shape[0], self.config.PADDING_SIZE, self.config.VOCABULARY_SIZE), dtype=torch.float32,
                            device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.dc_decoder_teacher_forced(self._h, ctypes.c_void_p(t.data_ptr()), kind, t.shape[0],
                                                           ctypes.c_void_p(g.data_ptr()),
                                                           ctypes.c_void_p(probs.data_ptr()), self._stream()))
        return probs.cpu().numpy() if was_numpy else probs

    def fit_generator(self, generator, steps_per_epoch=None, epochs=1, verbose=1, callbacks=None,
                      validation_data=None, validation_steps=None, max_queue_size=10, workers=1,
                      use_multiprocessing=False, shuffle=True, initial_epoch=0, **kwargs):
        """Keras fit_generator as the reference calls it (text_generation_model.py:470-472): the
        generator yields ([features, gt_captions], one_hot_targets); validation_data is one such
        batch (next(val_generator)) or a generator.  Callbacks receive Keras' on_epoch_end(epoch,
        logs) if they define it (ModelCheckpoint / CSVLogger stand-ins)."""
        if steps_per_epoch is None:
            raise ValueError("steps_per_epoch is required for a generator")
        hist = History()
        for cb in callbacks or []:
            if hasattr(cb, "set_model"):
                cb.set_model(self)
        for cb in callbacks or []:
            if hasattr(cb, "on_train_begin"):
                cb.on_train_begin()
        # max_queue_size / workers as in Keras: a worker thread keeps the next batches of the generator ready while the
        # current one trains (workers=0: the generator runs on this thread); use_multiprocessing is accepted and ignored
        with GeneratorQueue(generator, max_queue_size, workers) as batches:
            for epoch in range(initial_epoch, epochs):
                losses = []
                for _ in range(steps_per_epoch):
                    x, y = batches.get()[:2]
                    losses.append(self.train_on_batch(x, y))
                logs = {"loss": float(np.mean(losses))}
                if validation_data is not None:
                    if isinstance(validation_data, (tuple, list)):
                        logs["val_loss"] = self.test_on_batch(validation_data[0], validation_data[1])
                    else:
                        vs = [self.test_on_batch(*next(validation_data)[:2]) for _ in range(validation_steps or 1)]
                        logs["val_loss"] = float(np.mean(vs))
                hist.epoch.append(epoch)
                for k, v in logs.items():
                    hist.history.setdefault(k, []).append(v)
                if verbose:
                    print("Epoch %d/%d - " % (epoch + 1, epochs) + " - ".join("%s: %.4f" % kv for kv in logs.items()))
                for cb in callbacks or []:
                    if hasattr(cb, "on_epoch_end"):
                        cb.on_epoch_end(epoch, logs)
        for cb in callbacks or []:
            if hasattr(cb, "on_train_end"):
                cb.on_train_end()
        return hist

    def head_features(self, features):
        """`features_new` of build_lstm_model (:249-262): [N,1024]."""
        self._ready()
        t, kind, was_numpy = self._feats_to_device(features)
        out = torch.empty((t.shape[0], self._cfg.feat), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.dc_head_forward(self._h, ctypes.c_void_p(t.data_ptr()), kind, t.shape[0],
                                                 ctypes.c_void_p(out.data_ptr()), self._stream()))
        return out.cpu().numpy() if was_numpy else out


class InjectModelV2(_ModelBase):
    """build_model(features_shape, word_shape, config, units, inject=True) result."""

    def __init__(self, features_shape, word_shape, config, units, device=None, dtype="float32"):
        self.word_shape = tuple(word_shape)
        super().__init__(ARCH_V2_INJECT, config, units, list(features_shape), dtype, word_units=1024,
                         device=device)

    def predict(self, x, batch_size=None, verbose=0):
        """model.predict([features, words]) -> [N,V] next-word probabilities
        (evaluate_models/test_score_dense_captions.py:221)."""
        self._ready()
        feats, words = x
        t, kind, was_numpy = self._feats_to_device(feats)
        w = torch.as_tensor(np.asarray(words) if not isinstance(words, torch.Tensor) else words)
        if w.dim() != 2 or w.shape[0] != t.shape[0]:
            raise ValueError("words must be [N, L] with the same N as the features")
        w = w.to(torch.int32).to(self.device).contiguous()
        probs = torch.empty((t.shape[0], self.config.VOCABULARY_SIZE), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.dc_decoder_v2_predict(
                self._h, ctypes.c_void_p(t.data_ptr()), kind, ctypes.c_void_p(w.data_ptr()), t.shape[0],
                w.shape[1], ctypes.c_void_p(probs.data_ptr()), self._stream()))
        return probs.cpu().numpy() if was_numpy else probs

    def generate(self, features, return_probs=False, start_tokens=None, return_scores=False):
        """The reference's greedy loop (test_score_dense_captions.py:216-225): P-1 ids per RoI, started from
        [0] (argmax of the all-zero start vector) or, with ``start_tokens`` [N], from a given first word
        (eval_text_generation_model_v2.py:176-186 starts from the ground-truth first word)."""
        self._ready()
        t, kind, was_numpy = self._feats_to_device(features)
        N, P, V = t.shape[0], self.config.PADDING_SIZE, self.config.VOCABULARY_SIZE
        tokens = torch.empty((N, P - 1), dtype=torch.int32, device=self.device)
        probs = torch.empty((N, P - 1, V), dtype=torch.float32, device=self.device) if return_probs else None
        st = None
        if start_tokens is not None:
            st = torch.as_tensor(np.asarray(start_tokens) if not isinstance(start_tokens, torch.Tensor) else start_tokens)
            if st.dim() != 1 or st.shape[0] != N:
                raise ValueError("start_tokens must be [N]")
            st = st.to(self.device).to(torch.int32).contiguous()
        scores = torch.empty((N,), dtype=torch.float32, device=self.device) if return_scores else None
        with torch.cuda.device(self.device):
            _lib.check(self._lib.dc_decoder_v2_greedy_from(
                self._h, ctypes.c_void_p(t.data_ptr()), kind, N, ctypes.c_void_p(st.data_ptr()) if st is not None else None,
                ctypes.c_void_p(tokens.data_ptr()), ctypes.c_void_p(probs.data_ptr()) if probs is not None else None,
                ctypes.c_void_p(scores.data_ptr()) if scores is not None else None, self._stream()))
        outs = [tokens] + ([probs] if return_probs else []) + ([scores] if return_scores else [])
        if was_numpy:
            outs = [o.cpu().numpy() for o in outs]
        return outs[0] if len(outs) == 1 else tuple(outs)


def _v2_training_surface():
    """compile / train_on_batch / test_on_batch / fit_generator of the v2 inject model
    (text_generation_model_v2.py:263-267, 312-330): Adam(amsgrad=True) + keras.losses.categorical_crossentropy on
    [N, V] next-word targets; the generator yields ([features, words], one_hot_next_word)."""

    def _words_to_device(self, words, n):
        w = torch.as_tensor(np.asarray(words) if not isinstance(words, torch.Tensor) else words)
        if w.dim() != 2 or w.shape[0] != n:
            raise ValueError("words must be [N, L] with the same N as the features")
        return w.to(self.device).to(torch.int32).contiguous()

    def _next_word_ids(self, y, n):
        """one-hot [N, V] (what data_generator yields) or class ids [N]; an all-zero row = no target"""
        t = y if isinstance(y, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(y))
        t = t.to(self.device)
        if t.dim() == 2:
            ids = t.argmax(-1).to(torch.int32)
            ids = torch.where(t.sum(-1) > 0, ids, torch.full_like(ids, -1))
        else:
            ids = t.to(torch.int32)
        if ids.dim() != 1 or ids.shape[0] != n:
            raise ValueError("targets must be [N, V] one-hot rows or [N] class ids")
        return ids.contiguous()

    def train_step_device(self, features, words, targets, inv_count=0.0, loss_out=None):
        """Forward + categorical cross-entropy + backward; gradients stay in grad_buffer().  Returns the device scalar
        sum_rows(-log p_y) * inv_count (inv_count <= 0: 1/N).  No host sync."""
        self._ready()
        t, kind, _ = self._feats_to_device(features)
        w = self._words_to_device(words, t.shape[0])
        y = self._next_word_ids(targets, t.shape[0])
        loss = torch.empty((), dtype=torch.float32, device=self.device) if loss_out is None else loss_out
        with torch.cuda.device(self.device):
            _lib.check(self._lib.dc_decoder_v2_train_step(
                self._h, ctypes.c_void_p(t.data_ptr()), kind, t.shape[0], ctypes.c_void_p(w.data_ptr()), w.shape[1],
                ctypes.c_void_p(y.data_ptr()), ctypes.c_float(inv_count), ctypes.c_void_p(loss.data_ptr()), self._stream()))
        return loss

    def _inv(self, y_ids):
        cnt = int((y_ids >= 0).sum().item())
        return 1.0 / cnt if cnt else 0.0

    def train_on_batch(self, x, y=None, sample_weight=None, class_weight=None):
        if self.optimizer is None:
            raise RuntimeError("compile() the model first")
        if sample_weight is not None or class_weight is not None:
            raise NotImplementedError("sample/class weights are not used by the reference")
        feats, words = x
        ids = self._next_word_ids(y, len(feats))
        loss = self.train_step_device(feats, words, ids, self._inv(ids))
        self.apply_gradients()
        return float(loss.item())

    def test_on_batch(self, x, y=None, sample_weight=None):
        feats, words = x
        ids = self._next_word_ids(y, len(feats))
        return float(self.train_step_device(feats, words, ids, self._inv(ids)).item())

    return dict(_words_to_device=_words_to_device, _next_word_ids=_next_word_ids, train_step_device=train_step_device, _inv=_inv,
                train_on_batch=train_on_batch, test_on_batch=test_on_batch)


for _k, _v in _v2_training_surface().items():
    setattr(InjectModelV2, _k, _v)
for _k in ("compile", "grad_buffer", "param_buffer", "get_gradients", "apply_gradients", "fit_generator"):
    setattr(InjectModelV2, _k, getattr(RoiCaptionModel, _k))


def build_lstm_model(features_input, config, units, mode, dtype="float32", device=None):
    """Same signature as the reference (text_generation_model.py:235) plus `dtype`/`device`."""
    return RoiCaptionModel(features_input, config, units, mode, dtype=dtype, device=device)


def build_model(features_shape, word_shape, config, units, inject=True, device=None, dtype="float32"):
    """Same signature as the reference (text_generation_model_v2.py:140) plus `dtype`/`device`.  Only the
    inject variant ("m1") is on the hot path; the merge variant is out of scope."""
    if not inject:
        raise NotImplementedError("merge model (inject=False) is outside the hot path")
    return InjectModelV2(features_shape, word_shape, config, units, device=device, dtype=dtype)
