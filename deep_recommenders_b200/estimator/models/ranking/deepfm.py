"""Estimator-side DeepFM twin (reference estimator/models/ranking/deepfm.py:9-43):
``DeepFM(indicator_columns, embedding_columns, dnn_units, dnn_activation=relu, ...)(features)``
-> sigmoid(fm + dnn(concat(fm.embeddings), dnn_units + [1]))."""
from __future__ import annotations

import torch

from ..feature_interaction import FM
from ..feature_interaction.dnn import DnnTower, relu


class DeepFM(object):

    def __init__(self,
                 indicator_columns,
                 embedding_columns,
                 dnn_units,
                 dnn_activation=relu,
                 dnn_batch_normalization=False,
                 dnn_dropout=None,
                 sparse_lr=None, seed=None, device=None,
                 **dnn_kwargs):
        if dnn_batch_normalization or dnn_dropout is not None:
            raise NotImplementedError("batch normalisation / dropout are not on the accelerated path")
        self._dnn_hidden_units = list(dnn_units)
        self._fm = FM(indicator_columns, embedding_columns, sparse_lr=sparse_lr, seed=seed, device=device)
        self._dnn = DnnTower(self._dnn_hidden_units + [1], activation=dnn_activation, seed=seed)

    def parameters(self):
        yield from self._fm.parameters()
        yield from self._dnn.parameters()

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)

    def call(self, features):
        fm_outputs = self._fm(features)
        stack = self._fm._stack
        concat_embeddings = stack.view(stack.shape[0], -1)
        return torch.sigmoid(fm_outputs + self._dnn(concat_embeddings))
