"""ModelCheckpoint / CSVLogger with Keras 2 semantics -- the callbacks the reference's training scripts hand to
fit_generator (text_generation_model.py:461-462, text_generation_model_v2.py:303-304) -- and the hooks fit_generator drives."""
import warnings

import numpy as np
import pytest
import torch

from image_captioning_b200 import callbacks, parallel


class _Saver(object):
    def __init__(self):
        self.saved, self.full = [], []

    def save_weights(self, path):
        self.saved.append(path)

    def save(self, path):
        self.full.append(path)


def test_model_checkpoint_every_epoch_with_formatted_path():
    m = _Saver()
    cb = callbacks.ModelCheckpoint("w-{epoch:02d}-{val_loss:.2f}.h5", verbose=0, save_weights_only=True, mode="min")
    cb.set_model(m)
    for e, v in enumerate([3.0, 2.5, 2.75]):
        cb.on_epoch_end(e, {"loss": v + 1, "val_loss": v})
    assert m.saved == ["w-01-3.00.h5", "w-02-2.50.h5", "w-03-2.75.h5"] and m.full == []
    full = callbacks.ModelCheckpoint("m.h5")                 # save_weights_only=False -> model.save
    full.set_model(m)
    full.on_epoch_end(0, {"val_loss": 1.0})
    assert m.full == ["m.h5"]


@pytest.mark.parametrize("mode,monitor,values,kept", [("min", "val_loss", [3.0, 2.5, 2.75, 2.0], [0, 1, 3]),
                                                      ("max", "val_loss", [3.0, 2.5, 3.5], [0, 2]),
                                                      ("auto", "val_acc", [0.1, 0.3, 0.2], [0, 1]),
                                                      ("auto", "loss", [1.0, 2.0, 0.5], [0, 2])])
def test_model_checkpoint_save_best_only(mode, monitor, values, kept):
    m = _Saver()
    cb = callbacks.ModelCheckpoint("w{epoch}.h5", monitor=monitor, save_best_only=True, save_weights_only=True, mode=mode)
    cb.set_model(m)
    for e, v in enumerate(values):
        cb.on_epoch_end(e, {monitor: v})
    assert m.saved == ["w%d.h5" % (e + 1) for e in kept]
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        cb.on_epoch_end(9, {"other": 1.0})                   # monitored quantity missing: skip with a warning
    assert len(w) == 1 and len(m.saved) == len(kept)


def test_model_checkpoint_period_and_unknown_mode():
    m = _Saver()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        cb = callbacks.ModelCheckpoint("w{epoch}.h5", save_weights_only=True, mode="median", period=2)
    assert len(w) == 1
    cb.set_model(m)
    for e in range(5):
        cb.on_epoch_end(e, {"val_loss": 1.0})
    assert m.saved == ["w2.h5", "w4.h5"]


def test_csv_logger_rows_header_and_append(tmp_path):
    path = str(tmp_path / "log.csv")
    cb = callbacks.CSVLogger(path)
    cb.on_train_begin()
    cb.on_epoch_end(0, {"val_loss": 2.5, "loss": 3.0})
    cb.on_epoch_end(1, {"val_loss": 2.25, "loss": 2.75})
    cb.on_train_end()
    assert open(path).read().splitlines() == ["epoch,loss,val_loss", "0,3.0,2.5", "1,2.75,2.25"]
    more = callbacks.CSVLogger(path, append=True)
    more.on_train_begin()
    more.on_epoch_end(2, {"val_loss": 2.0, "loss": 2.5})
    more.on_train_end()
    assert open(path).read().splitlines()[-2:] == ["1,2.75,2.25", "2,2.5,2.0"] and open(path).read().count("epoch") == 1
    fresh = callbacks.CSVLogger(path, separator=";")
    fresh.on_epoch_end(0, {"loss": 1.0})                     # driven without on_train_begin: opens on first use
    fresh.on_train_end()
    assert open(path).read().splitlines() == ["epoch;loss", "0;1.0"]


class _Stub(object):
    """Training surface of RoiCaptionModel on the CPU (as tests/test_parallel.py) + save_weights."""

    def __init__(self):
        self.w = torch.zeros(3, 2, dtype=torch.float64)
        self.g = torch.zeros(6, dtype=torch.float64)
        self.saved = []

    def train_step_device(self, features, gt, targets, inv_count):
        f, y = torch.as_tensor(features, dtype=torch.float64), torch.as_tensor(gt, dtype=torch.float64)
        r = f @ self.w - y
        self.g.copy_((2.0 * inv_count * (f.t() @ r)).reshape(-1))
        return (r * r).sum() * inv_count

    def grad_buffer(self):
        return self.g

    def param_buffer(self):
        return self.w.view(-1)

    def apply_gradients(self, grad_scale=1.0):
        self.w -= 0.1 * grad_scale * self.g.view_as(self.w)

    def save_weights(self, path):
        self.saved.append(path)


def test_fit_generator_drives_the_callbacks(tmp_path):
    def gen():
        g = torch.Generator().manual_seed(0)
        f, y = torch.randn(8, 3, generator=g, dtype=torch.float64), torch.randn(8, 2, generator=g, dtype=torch.float64)
        while True:
            yield ([f, y], None)
    m = _Stub()
    log = str(tmp_path / "train.csv")
    hist = parallel.DataParallelTrainer(m).fit_generator(
        gen(), steps_per_epoch=2, epochs=3, verbose=0,
        callbacks=[callbacks.ModelCheckpoint(str(tmp_path / "w-{epoch:02d}.h5"), save_weights_only=True, mode="min"),
                   callbacks.CSVLogger(log)])
    assert [p[-7:] for p in m.saved] == ["w-01.h5", "w-02.h5", "w-03.h5"]
    rows = open(log).read().splitlines()
    assert rows[0] == "epoch,loss" and len(rows) == 4
    np.testing.assert_allclose([float(r.split(",")[1]) for r in rows[1:]], hist.history["loss"])
    assert hist.history["loss"][-1] < hist.history["loss"][0]


def test_dense_cap_config_surface(capsys):
    """DenseCapConfig as the scripts use it (text_generation_model.py:23-49, config.display() at :475)."""
    import image_captioning_b200 as pkg
    c = pkg.DenseCapConfig(100, np.zeros((100, 8), np.float32), 4)
    assert (c.VOCABULARY_SIZE, c.EMBEDDING_SIZE, c.BATCH_SIZE, c.PADDING_SIZE) == (100, 8, 4, 10)
    assert (c.STEPS_PER_EPOCH, c.VALIDATION_STEPS, c.GPU_COUNT, c.IMAGES_PER_GPU) == (500, 50, 1, 1)
    c.display()
    out = capsys.readouterr().out
    assert "Configurations:" in out and "VOCABULARY_SIZE" in out and "array(100, 8)" in out


def test_trainable_weight_names_follow_the_reference_graphs():
    from image_captioning_b200 import text_model as tm
    v1 = tm.trainable_weight_names(tm.ARCH_V1)
    assert len(v1) == 18 and "imgcap_embedding_layer/embeddings" not in v1 and not any("/moving_" in n for n in v1)
    assert "mrcnn_class_conv1/kernel" in v1 and "mrcnn_class_bn2/gamma" in v1          # the head trains in the v1 graph
    v2 = tm.trainable_weight_names(tm.ARCH_V2_INJECT)
    assert v2 == ["lstm_1/kernel", "lstm_1/recurrent_kernel", "lstm_1/bias", "imgcap_lstm/kernel",
                  "imgcap_lstm/recurrent_kernel", "imgcap_lstm/bias", "imgcap_d1/kernel", "imgcap_d1/bias"]
