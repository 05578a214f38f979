"""Cross layer (DCN-v2 matrix form) at ``deep_recommenders.keras.models.ranking.dcn.Cross``
and a ``DCN`` model.

``Cross`` mirrors reference keras/models/ranking/dcn.py:9-109 (constructor arguments :12-19,
the diag-scale assertion :32-33, the projection_dim range check :48-53, ``call(x0, x=None)``
:70-88, ``get_config`` keys :90-108).  One GEMM with the whole ``x0 * (xW + b + a*x) + x``
epilogue fused (dr_cross_fwd), hand-written backward (dr_cross_bwd).

The reference has no DCN *model* (SURVEY.md section 0 fact 3): ``DCN`` below is this repo's
definition for BASELINE config C3 -- stacked Cross layers in parallel with a DNN tower,
concat -> Dense(1) -> sigmoid, the structure of the reference's own test
(tests/keras/test_dcn.py:27-32) plus the deep tower.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
from torch import nn

from .... import ops
from ....embedding import EmbeddingCollection
from ...layers.base import (Dense, Layer, Model, get_initializer, get_regularizer, register_keras_serializable,
                            regularization_penalty, serialize_initializer)
from .deepfm import DNN


@register_keras_serializable()
class Cross(Layer):
    """ Cross net in Deep & Cross Network (DCN) """

    def __init__(self,
                 projection_dim: Optional[int] = None,
                 diag_scale: Optional[float] = 0.0,
                 use_bias: bool = True,
                 kernel_init="truncated_normal",
                 kernel_regu=None,
                 bias_init="zeros",
                 bias_regu=None,
                 seed: Optional[int] = None,
                 **kwargs):
        super().__init__(**kwargs)
        self._projection_dim = projection_dim
        self._diag_scale = diag_scale
        self._use_bias = use_bias
        self._kernel_init_id = kernel_init
        self._bias_init_id = bias_init
        self._kernel_init = get_initializer(kernel_init)
        self._kernel_regu = get_regularizer(kernel_regu)
        self._bias_init = get_initializer(bias_init)
        self._bias_regu = get_regularizer(bias_regu)
        self._seed = seed
        self.kernel = None      # full rank  [d, d]
        self.kernel_u = None    # low rank   [d, r]  (no bias)
        self.kernel_v = None    # low rank   [r, d]
        self.bias = None

        assert self._diag_scale >= 0, \
            ValueError("diag scale must be non-negative, got {}".format(self._diag_scale))

    def build(self, input_shape, device=None):
        last_dim = int(input_shape[-1])
        dev = device or ("cuda" if torch.cuda.is_available() else "cpu")
        gen = torch.Generator().manual_seed(self._seed) if self._seed is not None else None

        def mk(shape, init):
            t = torch.empty(shape, dtype=torch.float32, device=dev)
            init(t, gen)
            return nn.Parameter(t)

        if self._projection_dim is None:
            self.kernel = mk((last_dim, last_dim), self._kernel_init)
        else:
            if self._projection_dim < 0 or self._projection_dim > last_dim / 2:
                raise ValueError(
                    "`projection_dim` should be smaller than last_dim / 2 to improve "
                    "the model efficiency, and should be positive. Got "
                    "`projection_dim` {}, and last dimension of input {}".format(
                        self._projection_dim, last_dim))
            self.kernel_u = mk((last_dim, self._projection_dim), self._kernel_init)
            self.kernel_v = mk((self._projection_dim, last_dim), self._kernel_init)
        if self._use_bias:
            self.bias = mk((last_dim,), self._bias_init)
        self.built = True

    def forward(self, x0, x=None, **kwargs):
        if not self.built:
            self.build(tuple(x0.shape), device=x0.device)
        return self.call(x0, x, **kwargs)

    def call(self, x0, x=None, **kwargs):
        same = x is None or x is x0
        if x is None:
            x = x0
        if x0.shape[-1] != x.shape[-1]:
            raise ValueError("`x0` and `x` dim mismatch. "
                             "Got `x0` dim = {} and `x` dim = {}".format(
                                 x0.shape[-1], x.shape[-1]))
        return ops.CrossFn.apply(x0, x, self.kernel, self.kernel_u, self.kernel_v, self.bias,
                                 float(self._diag_scale or 0.0), same)

    @property
    def losses(self):
        out = []
        for k in (self.kernel, self.kernel_u, self.kernel_v):
            if k is not None and self._kernel_regu is not None:
                out.append(regularization_penalty(self._kernel_regu, k))
        if self.bias is not None and self._bias_regu is not None:
            out.append(regularization_penalty(self._bias_regu, self.bias))
        return out

    def get_config(self):
        config = {
            "projection_dim":
                self._projection_dim,
            "diag_scale":
                self._diag_scale,
            "use_bias":
                self._use_bias,
            "kernel_init":
                serialize_initializer(self._kernel_init_id),
            "kernel_regu":
                self._kernel_regu,
            "bias_init":
                serialize_initializer(self._bias_init_id),
            "bias_regu":
                self._bias_regu,
        }
        base_config = super().get_config()
        return {**base_config, **config}


class DCN(Model):
    """Embedding lookup -> [cross stack || DNN tower] -> concat -> Dense(1) -> sigmoid (config C3)."""

    def __init__(self, rows: Sequence[int], dim: int, num_cross: int = 3, dnn_units: Sequence[int] = (512, 256, 128),
                 projection_dim: Optional[int] = None, diag_scale: float = 0.0, dnn_activation="relu",
                 sparse_lr: Optional[float] = None, seed: Optional[int] = None, device=None, **kwargs):
        super().__init__(**kwargs)
        self.embeddings = EmbeddingCollection(rows, dim, with_linear=False, device=device, seed=seed,
                                              sparse_lr=sparse_lr)
        self.cross = nn.ModuleList([Cross(projection_dim=projection_dim, diag_scale=diag_scale,
                                          seed=None if seed is None else seed + 100 + i)
                                    for i in range(num_cross)])
        self.dnn = DNN(list(dnn_units), activation=dnn_activation, out_units=0,
                       seed=None if seed is None else seed + 200)
        self.head = Dense(1, seed=None if seed is None else seed + 300)
        self._cfg = dict(rows=list(rows), dim=dim, num_cross=num_cross, dnn_units=list(dnn_units),
                         projection_dim=projection_dim, diag_scale=diag_scale, dnn_activation=dnn_activation)
        self.built = True

    def logits(self, ids: torch.Tensor) -> torch.Tensor:
        stack, _ = self.embeddings(ids, want_logit=False)
        x0 = stack.view(stack.shape[0], -1)
        x = x0
        for i, layer in enumerate(self.cross):
            x = layer(x0, None if i == 0 else x)
        deep = self.dnn(x0)
        return self.head(torch.cat([x, deep], dim=1))

    def call(self, ids, **kwargs):
        return torch.sigmoid(self.logits(ids))

    def get_config(self):
        return {**super().get_config(), **self._cfg}
