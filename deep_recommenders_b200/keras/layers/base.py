"""The tf.keras Layer protocol (``__init__ / build / call / get_config``) hosted on
``torch.nn.Module``, plus ``Dense`` and the initializer / regularizer name registry.

The reference's "operator API" is this protocol: every layer is a
``tf.keras.layers.Layer`` decorated with ``register_keras_serializable`` and round-trips
through ``get_config`` (reference: keras/models/ranking/fm.py:7-37, dcn.py:8,90-108).
TensorFlow is not available in this environment, so the protocol is mirrored here; the
arithmetic goes to the C-ABI library through ``deep_recommenders_b200.ops``.
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import torch
from torch import nn

from ... import ops

_SERIALIZABLE: dict[str, type] = {}


def register_keras_serializable(package: str = "Custom", name: Optional[str] = None):
    """Mirror of tf.keras.utils.register_keras_serializable (decorator)."""

    def deco(cls):
        _SERIALIZABLE[f"{package}>{name or cls.__name__}"] = cls
        cls._keras_registered_name = f"{package}>{name or cls.__name__}"
        return cls

    return deco


def get_registered(name: str) -> type:
    return _SERIALIZABLE[name]


# ---- initializers ---------------------------------------------------------------------------
def _truncated_normal_(t: torch.Tensor, std: float, gen=None) -> torch.Tensor:
    # TF's truncated normal re-draws outside 2 sigma.
    with torch.no_grad():
        tmp = torch.empty(t.shape, dtype=torch.float32)
        nn.init.trunc_normal_(tmp, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=gen)
        t.copy_(tmp)
    return t


def _glorot_uniform_(t: torch.Tensor, gen=None) -> torch.Tensor:
    fan_in, fan_out = t.shape[0], t.shape[1]
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    with torch.no_grad():
        tmp = torch.empty(t.shape, dtype=torch.float32).uniform_(-limit, limit, generator=gen)
        t.copy_(tmp)
    return t


INITIALIZERS: dict[str, Callable] = {
    "zeros": lambda t, gen=None: t.detach().zero_(),
    "ones": lambda t, gen=None: t.detach().fill_(1.0),
    "glorot_uniform": _glorot_uniform_,
    # tf.keras.initializers.get("truncated_normal") == TruncatedNormal(mean=0, stddev=0.05)
    "truncated_normal": lambda t, gen=None: _truncated_normal_(t, 0.05, gen),
}


def get_initializer(identifier):
    """tf.keras.initializers.get for the names the reference uses; callables pass through."""
    if callable(identifier):
        return identifier
    if isinstance(identifier, dict):
        identifier = identifier.get("class_name", "")
    key = {"TruncatedNormal": "truncated_normal", "Zeros": "zeros", "Ones": "ones",
           "GlorotUniform": "glorot_uniform"}.get(identifier, identifier)
    if key not in INITIALIZERS:
        raise ValueError(f"Unknown initializer: {identifier!r}")
    return INITIALIZERS[key]


def serialize_initializer(identifier):
    """Shape of tf.keras.initializers.serialize: {'class_name': ..., 'config': {...}}."""
    if isinstance(identifier, dict):
        return identifier
    if callable(identifier):
        return {"class_name": getattr(identifier, "__name__", "callable"), "config": {}}
    cls = {"truncated_normal": "TruncatedNormal", "zeros": "Zeros", "ones": "Ones",
           "glorot_uniform": "GlorotUniform"}.get(identifier, identifier)
    cfg = {"mean": 0.0, "stddev": 0.05, "seed": None} if cls == "TruncatedNormal" else (
        {"seed": None} if cls == "GlorotUniform" else {})
    return {"class_name": cls, "config": cfg}


def get_regularizer(identifier):
    if identifier is None:
        return None
    if isinstance(identifier, dict):
        return identifier
    if identifier in ("l1", "l2", "l1_l2"):
        # tf.keras.regularizers.get("l2") == L2(l2=0.01)
        return {"class_name": identifier.upper() if identifier != "l1_l2" else "L1L2",
                "config": {k: 0.01 for k in (("l1",) if identifier == "l1" else ("l2",) if identifier == "l2" else ("l1", "l2"))}}
    raise ValueError(f"Unknown regularizer: {identifier!r}")


def regularization_penalty(reg: Optional[dict], w: torch.Tensor) -> torch.Tensor:
    """Tiny host-side glue (not on the hot path): l1*sum|w| + l2*sum w^2."""
    if reg is None:
        return w.new_zeros(())
    cfg = reg["config"]
    out = w.new_zeros(())
    if cfg.get("l1"):
        out = out + cfg["l1"] * w.abs().sum()
    if cfg.get("l2"):
        out = out + cfg["l2"] * (w * w).sum()
    return out


# ---- Layer / Model --------------------------------------------------------------------------
class Layer(nn.Module):
    """tf.keras.layers.Layer protocol: lazy ``build(input_shape)`` then ``call(...)``."""

    def __init__(self, name: Optional[str] = None, trainable: bool = True, dtype: str = "float32", **kwargs):
        if kwargs:
            raise TypeError(f"Keyword argument not understood: {sorted(kwargs)}")
        super().__init__()
        self._name = name or self.__class__.__name__.lower()
        self.trainable = trainable
        self._dtype = dtype
        self.built = False

    @property
    def name(self) -> str:
        return self._name

    def build(self, input_shape):
        self.built = True

    def call(self, *args, **kwargs):
        raise NotImplementedError

    def forward(self, *args, **kwargs):
        if not self.built:
            first = args[0] if args else next(iter(kwargs.values()))
            shape = tuple(first.shape) if hasattr(first, "shape") else None
            self.build(shape)
            self.built = True
        return self.call(*args, **kwargs)

    def get_config(self) -> dict:
        return {"name": self._name, "trainable": self.trainable, "dtype": self._dtype}

    @classmethod
    def from_config(cls, config: dict):
        return cls(**config)

    @property
    def losses(self) -> list:
        return []


from ..engine import TrainingLoopMixin  # noqa: E402


class Model(TrainingLoopMixin, Layer):
    """tf.keras.Model stand-in: a Layer with ``predict``, ``compile`` / ``fit`` / ``evaluate`` (keras/engine.py) and
    config-based save / load."""

    @torch.no_grad()
    def predict(self, inputs, **kwargs):
        was = self.training
        self.eval()
        out = self(inputs)
        self.train(was)
        return out

    def save(self, path: str) -> None:
        """SavedModel round trip of the reference tests (tests/keras/test_fm.py:44-65): config + weights."""
        torch.save({"class": getattr(self, "_keras_registered_name", self.__class__.__name__),
                    "config": self.get_config(), "state": self.state_dict()}, path)


class Dense(Layer):
    """tf.keras.layers.Dense: y = activation(x @ kernel + bias), kernel [in, units].

    reference call sites: keras/models/ranking/fm.py:16-20, deepfm.py:30-34, dcn.py:39-67.
    """

    def __init__(self, units: int, activation=None, use_bias: bool = True,
                 kernel_initializer="glorot_uniform", bias_initializer="zeros",
                 kernel_regularizer=None, bias_regularizer=None, seed: Optional[int] = None, **kwargs):
        super().__init__(**kwargs)
        self.units = int(units)
        self.activation = activation
        self._act = ops.act_code(activation)
        self.use_bias = use_bias
        self._kernel_initializer = kernel_initializer
        self._bias_initializer = bias_initializer
        self._kernel_regularizer = get_regularizer(kernel_regularizer)
        self._bias_regularizer = get_regularizer(bias_regularizer)
        self._seed = seed
        self.kernel = None
        self.bias = None
        self._device = None

    def build(self, input_shape, device=None):
        in_dim = int(input_shape[-1])
        gen = torch.Generator().manual_seed(self._seed) if self._seed is not None else None
        dev = device or self._device or ("cuda" if torch.cuda.is_available() else "cpu")
        k = torch.empty((in_dim, self.units), dtype=torch.float32, device=dev)
        get_initializer(self._kernel_initializer)(k, gen)
        self.kernel = nn.Parameter(k)
        if self.use_bias:
            b = torch.empty((self.units,), dtype=torch.float32, device=dev)
            get_initializer(self._bias_initializer)(b, gen)
            self.bias = nn.Parameter(b)
        self.built = True

    def forward(self, x, **kwargs):
        if not self.built:
            self._device = x.device
            self.build(tuple(x.shape))
        return self.call(x)

    def call(self, x, **kwargs):
        return ops.DenseFn.apply(x, self.kernel, self.bias, self._act)

    @property
    def losses(self):
        out = []
        if self._kernel_regularizer is not None:
            out.append(regularization_penalty(self._kernel_regularizer, self.kernel))
        if self._bias_regularizer is not None and self.bias is not None:
            out.append(regularization_penalty(self._bias_regularizer, self.bias))
        return out

    def get_config(self):
        cfg = super().get_config()
        cfg.update({
            "units": self.units,
            "activation": self.activation if not callable(self.activation) else self.activation.__name__,
            "use_bias": self.use_bias,
            "kernel_initializer": serialize_initializer(self._kernel_initializer),
            "bias_initializer": serialize_initializer(self._bias_initializer),
            "kernel_regularizer": self._kernel_regularizer,
            "bias_regularizer": self._bias_regularizer,
        })
        return cfg


class Sequential(Model):
    """tf.keras.Sequential of Layers (deepfm.py:30-34 builds the DNN tower with it)."""

    def __init__(self, layers=None, **kwargs):
        super().__init__(**kwargs)
        self.layers = nn.ModuleList(layers or [])
        self.built = True

    def call(self, x, **kwargs):
        for layer in self.layers:
            x = layer(x)
        return x

    def get_config(self):
        cfg = super().get_config()
        cfg["layers"] = [{"class_name": l.__class__.__name__, "config": l.get_config()} for l in self.layers]
        return cfg
