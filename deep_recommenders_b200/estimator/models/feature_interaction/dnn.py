"""``dnn(inputs, hidden_units, activation=relu, batch_normalization=False, dropout=None)``.

Mirrors reference estimator/models/feature_interaction/dnn.py:9-31: activation on
``hidden_units[:-1]``, last layer linear.  TF1's ``tf.layers.dense`` creates variables in the
graph on each call; here a ``DnnTower`` object owns them, and ``dnn(...)`` builds one on
first use per (name, shapes) so a model_fn-style call site keeps working.
The reference's ``batch_normalization=True`` branch calls tf.nn.batch_normalization with
missing arguments (dnn.py:23-24) and cannot run; it raises here too.
"""
from __future__ import annotations

import torch

from ....keras.layers.base import Dense, Sequential

_TOWERS = {}


def relu(x):  # name used for the default activation, like tf.nn.relu
    return torch.relu(x)


class DnnTower(Sequential):
    def __init__(self, hidden_units, activation="relu", seed=None, **kwargs):
        act = getattr(activation, "__name__", activation)
        layers = [Dense(u, activation=act, seed=None if seed is None else seed + i)
                  for i, u in enumerate(hidden_units[:-1])]
        layers.append(Dense(hidden_units[-1], seed=None if seed is None else seed + len(hidden_units)))
        super().__init__(layers, **kwargs)


def dnn(inputs,
        hidden_units,
        activation=relu,
        batch_normalization=False,
        dropout=None,
        name="dnn",
        **kwargs):
    if batch_normalization is True:
        raise TypeError("batch_normalization=True is broken in the reference "
                        "(tf.nn.batch_normalization called without mean/variance) and is not supported")
    if dropout is not None:
        raise NotImplementedError("dropout inside dnn() is not on the accelerated path")
    key = (name, int(inputs.shape[-1]), tuple(hidden_units), getattr(activation, "__name__", activation))
    tower = _TOWERS.get(key)
    if tower is None:
        tower = DnnTower(list(hidden_units), activation=activation, **kwargs)
        _TOWERS[key] = tower
    return tower(inputs)
