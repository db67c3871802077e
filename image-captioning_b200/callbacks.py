"""The two Keras callbacks the reference's training scripts pass to ``fit_generator``
(/root/reference/dense_img_cap_separate_models/text_generation_model.py:461-462 and text_generation_model_v2.py:303-304:
``keras.callbacks.ModelCheckpoint(model_filepath, verbose=1, save_weights_only=True, mode='min')`` and
``keras.callbacks.CSVLogger(logs_filepath)``), with Keras 2 semantics, for a training script that no longer imports Keras.
``fit_generator`` of RoiCaptionModel / InjectModelV2 / parallel.DataParallelTrainer calls ``set_model``,
``on_train_begin``, ``on_epoch_end(epoch, logs)`` and ``on_train_end`` on whatever it is given, so Keras' own callback
objects work as well wherever Keras is installed."""
import csv
import os
import warnings

import numpy as np


class Callback(object):
    """keras.callbacks.Callback: the hooks fit_generator drives."""

    def __init__(self):
        self.model = None

    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None):
        pass

    def on_epoch_end(self, epoch, logs=None):
        pass

    def on_train_end(self, logs=None):
        pass


class ModelCheckpoint(Callback):
    """Save the model after every ``period`` epochs.  ``filepath`` may contain ``{epoch:02d}`` (1-based, as in Keras) and
    any key of ``logs`` (``{val_loss:.2f}``).  ``save_best_only``: only when ``monitor`` improved (``mode`` 'min' / 'max' /
    'auto' -- 'auto' maximises a monitor that contains 'acc' or starts with 'fmeasure', minimises everything else).
    ``save_weights_only`` chooses ``model.save_weights(filepath)`` over ``model.save(filepath)``; the models of this
    package store weights as an .npz keyed by Keras weight name and have no other state, so both write the same file."""

    def __init__(self, filepath, monitor="val_loss", verbose=0, save_best_only=False, save_weights_only=False, mode="auto",
                 period=1):
        super(ModelCheckpoint, self).__init__()
        self.filepath, self.monitor, self.verbose = filepath, monitor, verbose
        self.save_best_only, self.save_weights_only, self.period = save_best_only, save_weights_only, period
        self.epochs_since_last_save = 0
        if mode not in ("auto", "min", "max"):
            warnings.warn("ModelCheckpoint mode %s is unknown, fallback to auto mode." % mode, RuntimeWarning)
            mode = "auto"
        if mode == "max" or (mode == "auto" and ("acc" in monitor or monitor.startswith("fmeasure"))):
            self.monitor_op, self.best = np.greater, -np.inf
        else:
            self.monitor_op, self.best = np.less, np.inf

    def _save(self, path):
        if self.save_weights_only or not hasattr(self.model, "save"):
            self.model.save_weights(path)
        else:
            self.model.save(path)

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        self.epochs_since_last_save += 1
        if self.epochs_since_last_save < self.period:
            return
        self.epochs_since_last_save = 0
        path = self.filepath.format(epoch=epoch + 1, **logs)
        if self.save_best_only:
            current = logs.get(self.monitor)
            if current is None:
                warnings.warn("Can save best model only with %s available, skipping." % self.monitor, RuntimeWarning)
            elif self.monitor_op(current, self.best):
                if self.verbose > 0:
                    print("\nEpoch %05d: %s improved from %0.5f to %0.5f, saving model to %s"
                          % (epoch + 1, self.monitor, self.best, current, path))
                self.best = current
                self._save(path)
            elif self.verbose > 0:
                print("\nEpoch %05d: %s did not improve from %0.5f" % (epoch + 1, self.monitor, self.best))
        else:
            if self.verbose > 0:
                print("\nEpoch %05d: saving model to %s" % (epoch + 1, path))
            self._save(path)


class CSVLogger(Callback):
    """One row per epoch: ``epoch`` (0-based, as Keras writes it) and every key of ``logs`` in sorted order; the header is
    written once (not again when ``append=True`` finds a file that already has content)."""

    def __init__(self, filename, separator=",", append=False):
        super(CSVLogger, self).__init__()
        self.filename, self.sep, self.append = filename, separator, append
        self.writer, self.keys, self.csv_file, self.append_header = None, None, None, True

    def on_train_begin(self, logs=None):
        if self.append:
            if os.path.exists(self.filename):
                with open(self.filename, "r") as f:
                    self.append_header = not bool(len(f.readline()))
            self.csv_file = open(self.filename, "a", newline="")
        else:
            self.csv_file = open(self.filename, "w", newline="")

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        if self.csv_file is None:                           # driven without on_train_begin: open on first use
            self.on_train_begin()

        def cell(v):
            if isinstance(v, (list, tuple, np.ndarray)) and np.ndim(v) > 0:
                return '"[%s]"' % ", ".join(map(str, v))
            return v
        if self.keys is None:
            self.keys = sorted(logs.keys())
        if self.writer is None:
            self.writer = csv.DictWriter(self.csv_file, fieldnames=["epoch"] + self.keys, delimiter=self.sep)
            if self.append_header:
                self.writer.writeheader()
        row = {"epoch": epoch}
        row.update((k, cell(logs.get(k, "NA"))) for k in self.keys)
        self.writer.writerow(row)
        self.csv_file.flush()

    def on_train_end(self, logs=None):
        if self.csv_file is not None:
            self.csv_file.close()
        self.csv_file, self.writer = None, None
