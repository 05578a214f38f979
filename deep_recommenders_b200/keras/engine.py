"""Keras-style ``compile`` / ``fit`` / ``evaluate`` for the models in this package -- a THIN loop over the CUDA layers
(autograd + torch.optim), so that the reference's Keras example
(/root/reference/examples/train_deepfm_on_movielens_keras.py:38-54: ``model.compile(loss=..., optimizer=Adam(),
metrics=[AUC(), Precision(), Recall()])``, ``model.fit(input_fn, epochs=..., steps_per_epoch=..., validation_data=...,
validation_steps=..., callbacks=[EarlyStopping(patience=3)])``) has every name it uses.  The objects below stand in for
the ``tf.keras.losses / optimizers / metrics / callbacks`` members that script touches; nothing else of Keras is
restated.  The high-throughput path is ``DeepFMTrainStep`` (one CUDA-graph replay per step); this loop is the
drop-in convenience surface.
"""
from __future__ import annotations

import itertools
from typing import Callable, Iterable, List, Optional

import numpy as np
import torch


# ---- tf.keras.losses ----------------------------------------------------------------------------------------------
def binary_crossentropy(y_true: torch.Tensor, y_pred: torch.Tensor) -> torch.Tensor:
    """tf.keras.losses.binary_crossentropy on probabilities (Keras clips to [eps, 1 - eps], eps = 1e-7)."""
    p = y_pred.clamp(1e-7, 1.0 - 1e-7)
    return -(y_true * torch.log(p) + (1.0 - y_true) * torch.log(1.0 - p)).mean(dim=-1)


_LOSSES = {"binary_crossentropy": binary_crossentropy}


# ---- tf.keras.optimizers ------------------------------------------------------------------------------------------
class Adam:
    """tf.keras.optimizers.Adam defaults (learning_rate 0.001, beta 0.9 / 0.999, epsilon 1e-7)."""

    def __init__(self, learning_rate: float = 0.001, beta_1: float = 0.9, beta_2: float = 0.999, epsilon: float = 1e-7):
        self.kw = dict(lr=learning_rate, betas=(beta_1, beta_2), eps=epsilon)

    def build(self, params):
        return torch.optim.Adam(params, **self.kw)


class SGD:
    def __init__(self, learning_rate: float = 0.01, momentum: float = 0.0):
        self.kw = dict(lr=learning_rate, momentum=momentum)

    def build(self, params):
        return torch.optim.SGD(params, **self.kw)


_OPTIMIZERS = {"adam": Adam, "sgd": SGD}


# ---- tf.keras.metrics ---------------------------------------------------------------------------------------------
class _Confusion:
    """Thresholded confusion counts, accumulated on the device (no host sync per batch)."""

    def __init__(self, thresholds):
        self.thresholds = torch.as_tensor(thresholds, dtype=torch.float32)
        self.reset_state()

    def reset_state(self):
        self.tp = self.fp = self.fn = self.tn = None

    def update_state(self, y_true, y_pred):
        t = self.thresholds.to(y_pred.device)
        pred = y_pred.reshape(1, -1) > t.reshape(-1, 1)
        pos = (y_true.reshape(1, -1) > 0.5)
        add = lambda old, new: new if old is None else old + new
        self.tp = add(self.tp, (pred & pos).sum(1).double())
        self.fp = add(self.fp, (pred & ~pos).sum(1).double())
        self.fn = add(self.fn, (~pred & pos).sum(1).double())
        self.tn = add(self.tn, (~pred & ~pos).sum(1).double())


class AUC(_Confusion):
    """tf.keras.metrics.AUC defaults: ROC, 200 thresholds, trapezoidal ("interpolation") summation."""
    name = "auc"

    def __init__(self, num_thresholds: int = 200):
        eps = 1e-7
        th = [0.0 - eps] + [(i + 1) / (num_thresholds - 1) for i in range(num_thresholds - 2)] + [1.0 + eps]
        super().__init__(th)

    def result(self) -> float:
        if self.tp is None:
            return 0.0
        tpr = (self.tp / (self.tp + self.fn).clamp_min(1e-12)).cpu().numpy()
        fpr = (self.fp / (self.fp + self.tn).clamp_min(1e-12)).cpu().numpy()
        return float(np.sum((fpr[:-1] - fpr[1:]) * (tpr[:-1] + tpr[1:]) / 2.0))


class Precision(_Confusion):
    name = "precision"

    def __init__(self, thresholds: float = 0.5):
        super().__init__([thresholds])

    def result(self) -> float:
        return 0.0 if self.tp is None else float(self.tp[0] / (self.tp[0] + self.fp[0]).clamp_min(1e-12))


class Recall(_Confusion):
    name = "recall"

    def __init__(self, thresholds: float = 0.5):
        super().__init__([thresholds])

    def result(self) -> float:
        return 0.0 if self.tp is None else float(self.tp[0] / (self.tp[0] + self.fn[0]).clamp_min(1e-12))


# ---- tf.keras.callbacks -------------------------------------------------------------------------------------------
class EarlyStopping:
    """tf.keras.callbacks.EarlyStopping(monitor="val_loss", patience=0, min_delta=0): stop when the monitored value
    has not improved for `patience` epochs."""

    def __init__(self, monitor: str = "val_loss", patience: int = 0, min_delta: float = 0.0):
        self.monitor, self.patience, self.min_delta = monitor, int(patience), float(min_delta)
        self.best, self.wait, self.stopped_epoch = None, 0, None

    def on_epoch_end(self, epoch: int, logs: dict) -> bool:
        cur = logs.get(self.monitor)
        if cur is None:
            return False
        if self.best is None or cur < self.best - self.min_delta:
            self.best, self.wait = cur, 0
            return False
        self.wait += 1
        if self.wait >= self.patience:
            self.stopped_epoch = epoch
            return True
        return False


# ---- the loop -----------------------------------------------------------------------------------------------------
def _batches(data) -> Iterable:
    """`data`: an iterable / generator of (features, labels) batches, or a zero-argument callable returning one (the
    reference passes `movielens.training_input_fn`, a tf.data pipeline)."""
    return data() if callable(data) else data


def _labels(y, device) -> torch.Tensor:
    t = y if isinstance(y, torch.Tensor) else torch.as_tensor(np.asarray(y, dtype=np.float32))
    return t.to(device=device, dtype=torch.float32).reshape(t.shape[0], -1)


class TrainingLoopMixin:
    """compile / fit / evaluate for `Model` (keras/layers/base.py)."""

    def compile(self, loss="binary_crossentropy", optimizer="adam", metrics: Optional[List] = None, **kwargs):
        self._loss_fn: Callable = _LOSSES[loss] if isinstance(loss, str) else loss
        self._optimizer_spec = _OPTIMIZERS[optimizer]() if isinstance(optimizer, str) else optimizer
        self._optimizer = None
        self._metrics = list(metrics or [])
        return self

    def _step_loss(self, x, y):
        pred = self(x)
        yt = _labels(y, pred.device)
        loss = self._loss_fn(yt, pred).mean()
        reg = [l for l in getattr(self, "losses", [])] if hasattr(self, "losses") else []
        for r in reg:
            loss = loss + r
        return pred, yt, loss

    def evaluate(self, data, steps: Optional[int] = None) -> dict:
        if not hasattr(self, "_loss_fn"):
            raise RuntimeError("call compile() before evaluate()")
        for m in self._metrics:
            m.reset_state()
        tot, n = None, 0
        was = self.training
        self.eval()
        with torch.no_grad():
            for x, y in itertools.islice(_batches(data), steps):
                pred, yt, loss = self._step_loss(x, y)
                tot = loss.detach().double() if tot is None else tot + loss.detach().double()
                n += 1
                for m in self._metrics:
                    m.update_state(yt, pred)
        self.train(was)
        logs = {"loss": float(tot / max(n, 1)) if tot is not None else float("nan")}
        logs.update({m.name: m.result() for m in self._metrics})
        return logs

    def fit(self, x, y=None, epochs: int = 1, steps_per_epoch: Optional[int] = None, validation_data=None,
            validation_steps: Optional[int] = None, callbacks: Optional[List] = None, verbose: int = 0):
        """`x`: iterable (or callable returning one) of (features, labels) batches that is consumed ACROSS epochs, as
        the reference's repeated tf.data pipeline is; `steps_per_epoch` batches make one epoch.  Returns a dict of
        per-epoch logs (`history`)."""
        if not hasattr(self, "_loss_fn"):
            raise RuntimeError("call compile() before fit()")
        if y is not None:
            x = [(x, y)] * (steps_per_epoch or 1) * epochs
        stream = iter(_batches(x))
        history: dict = {}
        for epoch in range(epochs):
            for m in self._metrics:
                m.reset_state()
            tot, n = None, 0
            for xb, yb in itertools.islice(stream, steps_per_epoch):
                pred, yt, loss = self._step_loss(xb, yb)
                if self._optimizer is None:        # parameters of lazily built layers exist after the first forward
                    self._optimizer = self._optimizer_spec.build([p for p in self.parameters() if p.requires_grad])
                self._optimizer.zero_grad(set_to_none=True)
                loss.backward()
                self._optimizer.step()
                tot = loss.detach().double() if tot is None else tot + loss.detach().double()
                n += 1
                for m in self._metrics:
                    m.update_state(yt, pred.detach())
            if n == 0:
                break                                # input exhausted
            logs = {"loss": float(tot / n)}
            logs.update({m.name: m.result() for m in self._metrics})
            if validation_data is not None:
                logs.update({"val_" + k: v for k, v in self.evaluate(validation_data, validation_steps).items()})
            for k, v in logs.items():
                history.setdefault(k, []).append(v)
            if verbose:
                print(f"epoch {epoch + 1}/{epochs} " + " ".join(f"{k}={v:.4f}" for k, v in logs.items()), flush=True)
            if any(cb.on_epoch_end(epoch, logs) for cb in (callbacks or [])):
                break
        self.history = history
        return history
