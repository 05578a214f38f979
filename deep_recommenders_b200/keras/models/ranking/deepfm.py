"""DeepFM at the reference's import path ``deep_recommenders.keras.models.ranking.DeepFM``.

Mirrors reference keras/models/ranking/deepfm.py:9-55: same constructor arguments
(indicator_columns, embedding_columns, dnn_units_size, dnn_activation="relu"), same
``call(inputs)`` -> sigmoid([B,1]) and same ``get_config`` keys.  One fused CUDA launch
produces both the stacked embeddings (the DNN input) and the FM logit; the DNN tower
(``DNN``: Dense(u, act)... + Dense(1), deepfm.py:30-34) runs on dr_dense_fwd/bwd.
"""
from __future__ import annotations

from typing import Sequence

import torch

from ...layers.base import Dense, Sequential
from .fm import _ColumnModel


class DNN(Sequential):
    """The deep tower the reference builds inline: [Dense(u, activation) for u in units] + [Dense(1)].

    (BASELINE.json's north_star calls this layer `DNN`; the reference has no class of that
    name -- keras deepfm.py:30-34 and estimator dnn.py:9-31 are the two spellings of it.)
    """

    def __init__(self, units: Sequence[int], activation="relu", out_units: int = 1, seed=None, **kwargs):
        layers = [Dense(u, activation=activation, seed=None if seed is None else seed + i)
                  for i, u in enumerate(units)]
        if out_units:
            layers.append(Dense(out_units, seed=None if seed is None else seed + len(units)))
        super().__init__(layers, **kwargs)
        self._units = list(units)
        self._activation = activation


class DeepFM(_ColumnModel):

    def __init__(self,
                 indicator_columns,
                 embedding_columns,
                 dnn_units_size,
                 dnn_activation="relu",
                 **kwargs):
        seed = kwargs.get("seed")
        super().__init__(indicator_columns, embedding_columns, **kwargs)
        self._dnn_units_size = list(dnn_units_size)
        self._dnn_activation = dnn_activation
        self._dnn = DNN(self._dnn_units_size, activation=self._dnn_activation, seed=seed)

    def logits(self, inputs) -> torch.Tensor:
        stack, fm_logit = self._embed(inputs)                           # [B,S,D], [B]
        concat_embeddings = stack.view(stack.shape[0], -1)              # tf.concat(embeddings, axis=1)
        return fm_logit.unsqueeze(1) + self._dnn(concat_embeddings)

    def call(self, inputs, **kwargs):
        return torch.sigmoid(self.logits(inputs))

    def get_config(self):
        config = {
            "dnn_units_size": self._dnn_units_size,
            "dnn_activation": self._dnn_activation,
        }
        base_config = super().get_config()
        return {**base_config, **config}
