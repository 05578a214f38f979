"""EmbeddingCollection: the S embedding tables + first-order weights of a ranking model, laid
out for HBM as ONE arena per kind so a batch is served by ONE fused kernel launch.

Replaces the per-column ``tf.keras.layers.DenseFeatures(embedding_column)`` /
``DenseFeatures(indicator_columns) -> Dense(1)`` objects the reference builds
(keras/models/ranking/fm.py:47-51, deepfm.py:24-28; estimator/.../fm.py:43-52).

Layout in HBM
  weight : [sum_s rows_s, D] fp32   table s = rows [off_s, off_s + rows_s); every row is
           D*4 bytes (multiple of 16) so every row base is 16-B aligned for LDG.128.
  linear : [sum_s rows_s]   fp32   the Dense(1, kernel_initializer="zeros") kernel of the
           reference's multi-hot first-order term, stored as one scalar per (slot, id).
  bias   : [1]                     that Dense's bias.
The reference densifies the indicator columns into a [B, sum_s N_s] multi-hot and multiplies
by a [sum N, 1] kernel (fm.py:16-20,26); the sparse gather-sum here is the same number.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch
from torch import nn

from . import ops


class EmbeddingCollection(nn.Module):
    def __init__(self, rows: Sequence[int], dim: int, with_linear: bool = True, device=None,
                 init_stddev: Optional[float] = None, seed: Optional[int] = None,
                 sparse_lr: Optional[float] = None, init: str = "truncated_normal"):
        super().__init__()
        if dim % 4 != 0 or not (4 <= dim <= 128):
            raise ValueError(f"embedding dimension must be a multiple of 4 in [4,128], got {dim}")
        self.rows_list = [int(r) for r in rows]
        self.dim = int(dim)
        self.num_slots = len(self.rows_list)
        self.sparse_lr = sparse_lr           # None: dense .grad via autograd; float: fused sparse SGD
        dev = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
        offs = [0]
        for r in self.rows_list:
            offs.append(offs[-1] + r)
        self.total_rows = offs[-1]
        std = init_stddev if init_stddev is not None else 1.0 / math.sqrt(self.dim)
        w = torch.empty((self.total_rows, self.dim), dtype=torch.float32, device=dev)
        if init == "truncated_normal":
            # TF embedding_column default: truncated_normal_initializer(mean=0, stddev=1/sqrt(D))
            gen = None
            if seed is not None:
                gen = torch.Generator(device=dev).manual_seed(seed)
            nn.init.trunc_normal_(w, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=gen)
        elif init == "zeros":
            w.zero_()
        elif init != "empty":
            raise ValueError(f"unknown init {init!r}")
        self.weight = nn.Parameter(w)
        self.linear = nn.Parameter(torch.zeros((self.total_rows,), dtype=torch.float32, device=dev)) if with_linear else None
        self.bias = nn.Parameter(torch.zeros((1,), dtype=torch.float32, device=dev)) if with_linear else None
        self.register_buffer("_offsets", torch.tensor(offs[:-1], dtype=torch.int64, device=dev), persistent=False)
        self.register_buffer("_rows", torch.tensor(self.rows_list, dtype=torch.int64, device=dev), persistent=False)
        self._ptr_cache = {}

    # device pointer arrays for the kernel (computed on device: no host sync)
    def pointers(self, weight: torch.Tensor, linear: Optional[torch.Tensor], cache: bool = True):
        key = (weight.data_ptr(), None if linear is None else linear.data_ptr())
        if cache and key in self._ptr_cache:
            return self._ptr_cache[key]
        offs = self._offsets.to(weight.device)
        tp = offs * (self.dim * 4) + weight.data_ptr()
        lp = (offs * 4 + linear.data_ptr()) if linear is not None else tp
        out = (tp, lp, self._rows.to(weight.device))
        if cache:
            self._ptr_cache = {key: out}
        return out

    def table(self, s: int) -> torch.Tensor:
        o = int(self._offsets[s])
        return self.weight[o:o + self.rows_list[s]]

    def linear_of(self, s: int) -> torch.Tensor:
        o = int(self._offsets[s])
        return self.linear[o:o + self.rows_list[s]]

    def forward(self, ids: torch.Tensor, want_logit: bool = True):
        """ids [B, S] int64/int32 -> (stack [B,S,D], logit [B] = bias + linear + FM 2nd order)."""
        if ids.dim() != 2 or ids.shape[1] != self.num_slots:
            raise ValueError(f"ids must be [B, {self.num_slots}], got {tuple(ids.shape)}")
        want = want_logit
        stack, logit = ops.EmbedFM.apply(ids, self.weight, self.linear if want else None,
                                         self.bias if want else None, self, want)
        return stack, (logit if want else None)
