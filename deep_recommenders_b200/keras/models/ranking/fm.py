"""Factorization Machine layer and model at the reference's import path
``deep_recommenders.keras.models.ranking.{FM, FactorizationMachine}``.

Mirrors (same names, call signatures, get_config keys, error behaviour):
  reference keras/models/ranking/fm.py:8-37   class FM(Layer)
  reference keras/models/ranking/fm.py:40-72  class FactorizationMachine(Model)
but the arithmetic runs in the CUDA library: `FM.call` on dense tensors uses dr_dense_fwd +
dr_fm_fwd; the model uses the fused gather+linear+FM kernel (dr_embed_fm_fwd), one launch for
all slots instead of one DenseFeatures layer per column.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from .... import ops
from ....embedding import EmbeddingCollection
from ....feature_column import EmbeddingColumn, IndicatorColumn
from ....hashing import ids_and_bags
from ...layers.base import Dense, Layer, Model, register_keras_serializable


@register_keras_serializable()
class FM(Layer):
    """ Factorization Machine """

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self._linear = None

    def build(self, input_shape):
        self._linear = Dense(units=1, kernel_initializer="zeros", name="linear")
        self.built = True

    def call(self, sparse_inputs, embedding_inputs=None, **kwargs):
        if embedding_inputs is None:
            return self._linear(sparse_inputs)
        interaction = ops.FMInteraction.apply(embedding_inputs)          # [B, 1]
        return self._linear(sparse_inputs) + interaction


class _ColumnModel(Model):
    """Shared plumbing of FactorizationMachine / DeepFM: columns -> one EmbeddingCollection."""

    def __init__(self, indicator_columns: Sequence[IndicatorColumn], embedding_columns: Sequence[EmbeddingColumn],
                 sparse_lr: Optional[float] = None, seed: Optional[int] = None, device=None, **kwargs):
        super().__init__(**kwargs)
        self._indicator_columns = list(indicator_columns)
        self._embedding_columns = list(embedding_columns)
        dims = {c.dimension for c in self._embedding_columns}
        if len(dims) != 1:
            raise ValueError(f"all embedding columns must share one dimension (FM stacks them), got {sorted(dims)}")
        ekeys = [c.categorical_column.key for c in self._embedding_columns]
        ikeys = [c.categorical_column.key for c in self._indicator_columns]
        if sorted(ekeys) != sorted(ikeys):
            raise ValueError("indicator_columns and embedding_columns must be built over the same categorical "
                             f"columns (got {sorted(ikeys)} vs {sorted(ekeys)})")
        self._keys: List[str] = ekeys
        self._cat = {c.categorical_column.key: c.categorical_column for c in self._embedding_columns}
        combs = {c.combiner for c in self._embedding_columns}
        if combs != {"mean"}:
            raise NotImplementedError(f"only the default 'mean' combiner is implemented, got {sorted(combs)}")
        std = self._embedding_columns[0].initializer_stddev
        self.embeddings = EmbeddingCollection([self._cat[k].num_buckets for k in ekeys], dims.pop(), with_linear=True,
                                              device=device, init_stddev=std, seed=seed, sparse_lr=sparse_lr)
        self.built = True

    def _ids_and_bags(self, inputs):
        """dict[str -> raw feature] (reference: deepfm.py:39-43 iterates inputs.items()) or a ready [B,S] id matrix
        -> (ids [B,S] int64, {slot: (flat ids, row_splits)} for multi-valued features)."""
        if isinstance(inputs, torch.Tensor):
            return inputs, None
        return ids_and_bags(self._keys, self._cat, inputs, self.embeddings.weight.device)

    def _ids_matrix(self, inputs) -> torch.Tensor:
        ids, bags = self._ids_and_bags(inputs)
        if bags:
            raise ValueError("multi-valued features present: use _ids_and_bags")
        return ids

    def _embed(self, inputs):
        ids, bags = self._ids_and_bags(inputs)
        return self.embeddings(ids, want_logit=True, bags=bags)


class FactorizationMachine(_ColumnModel):

    def __init__(self, indicator_columns, embedding_columns, **kwargs):
        super().__init__(indicator_columns, embedding_columns, **kwargs)

    def call(self, inputs, training=None, mask=None):
        _, logit = self._embed(inputs)
        return torch.sigmoid(logit).unsqueeze(1)

    def logits(self, inputs) -> torch.Tensor:
        _, logit = self._embed(inputs)
        return logit

    def get_config(self):
        config = {
            "indicator_columns": self._indicator_columns,
            "embedding_columns": self._embedding_columns,
        }
        base_config = super().get_config()
        return {**base_config, **config}
