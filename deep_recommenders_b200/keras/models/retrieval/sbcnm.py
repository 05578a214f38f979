"""Sampling-bias-corrected in-batch softmax (two-tower retrieval task) at the reference's
import path ``deep_recommenders.keras.models.retrieval.sbcnm``.

Mirrors reference keras/models/retrieval/sbcnm.py:
  :9-10    MAX_FLOAT / MIN_FLOAT
  :33-49   HardNegativeMining(num_hard_negatives)(logits, labels) -> (logits, labels)
  :52-75   RemoveAccidentalNegative()(logits, labels, identifiers)
  :78-86   SamplingProbabilityCorrection()(logits, candidate_sampling_probability)
  :89-163  Retrieval(loss, metrics, temperature, num_hard_negatives).call(...)
In the reference the three optional branches of Retrieval.call (:136-146) name a module that
does not exist and raise NameError if taken; here they do what the helper layers define.

The default path (no hard negatives) is ONE fused kernel per direction: Q @ C^T with the
corrections, online log-sum-exp and the diagonal pick, never materialising the [B,B] scores
(dr_inbatch_softmax_fwd / _bwd).  `TwoTower` (not in the reference) adds the user / item
embedding towers that feed it (BASELINE config C4).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch
from torch import nn

from .... import ops
from ...layers.base import Layer, Model
from ..ranking.deepfm import DNN

MAX_FLOAT = np.finfo(np.float32).max / 100.0
MIN_FLOAT = np.finfo(np.float32).min / 100.0


def _gather_elements_along_row(data: torch.Tensor, column_indices: torch.Tensor) -> torch.Tensor:
    """data[i, column_indices[i, j]] (reference :15-30)."""
    if data.shape[0] != column_indices.shape[0]:
        raise ValueError("The number of rows of `data` and `column_indices` must be equal.")
    return torch.gather(data, 1, column_indices.to(torch.int64))


class HardNegativeMining(Layer):
    """Hard Negative: keep the positive and the num_hard_negatives highest-scoring negatives per row."""

    def __init__(self, num_hard_negatives: int, **kwargs):
        super().__init__(**kwargs)
        self._num_hard_negatives = num_hard_negatives
        self.built = True

    def call(self, logits: torch.Tensor, labels: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        num_sampled = min(self._num_hard_negatives + 1, logits.shape[1])
        eye = torch.eye(logits.shape[0], logits.shape[1], device=logits.device, dtype=logits.dtype)
        if labels.shape == eye.shape and torch.equal(labels, eye) and not logits.requires_grad:
            # in-batch labels: one kernel does the masked top-k and both gathers
            out_logits, out_labels, _ = ops.hard_negative_topk(logits, num_sampled)
            return out_logits, out_labels
        # general one-hot labels (the reference test permutes the positives) / autograd path:
        # selection is index work; the values are gathered differentiably.
        _, indices = torch.topk(logits.detach() + labels * MAX_FLOAT, k=num_sampled, dim=1, sorted=False)
        return _gather_elements_along_row(logits, indices), _gather_elements_along_row(labels, indices)


class RemoveAccidentalNegative(Layer):

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.built = True

    def call(self, logits: torch.Tensor, labels: torch.Tensor, identifiers: torch.Tensor) -> torch.Tensor:
        """Zeros logits of accidental negatives
        Args:
            logits: [batch_size, num_candidates] 2D tensor
            labels: [batch_size, num_candidates] one-hot 2D tensor
            identifiers: [num_candidates] candidates identifiers tensor
        Returns:
            logits: Modified logits.
        """
        identifiers = torch.as_tensor(identifiers, device=logits.device).reshape(-1, 1)
        positive_indices = torch.argmax(labels, dim=1)
        positive_identifier = identifiers[positive_indices]                  # [B, 1]
        duplicate = (positive_identifier == identifiers.t()).to(labels.dtype)  # [B, C]
        duplicate = duplicate - labels
        return logits + duplicate * MIN_FLOAT


class SamplingProbabilityCorrection(Layer):
    """Sampling probability correction."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.built = True

    def call(self, logits: torch.Tensor, candidate_sampling_probability: torch.Tensor) -> torch.Tensor:
        """Corrects the input logits to account for candidate sampling probability."""
        return logits - torch.log(candidate_sampling_probability)


class Retrieval(Layer):
    """Retrieval task: in-batch softmax loss over query x candidate scores."""

    def __init__(self,
                 loss=None,
                 metrics=None,
                 temperature: Optional[float] = None,
                 num_hard_negatives: Optional[int] = None,
                 **kwargs):
        super().__init__(**kwargs)
        # default = CategoricalCrossentropy(from_logits=True, reduction=SUM) (reference :100-102)
        self._loss = loss
        self._factorized_metrics = metrics
        self._temperature = temperature
        self._num_hard_negatives = num_hard_negatives
        self.built = True

    @property
    def factorized_metrics(self):
        """The metrics object used to compute retrieval metrics."""
        return self._factorized_metrics

    @factorized_metrics.setter
    def factorized_metrics(self, value) -> None:
        """Sets factorized metrics."""
        self._factorized_metrics = value

    def call(self,
             query_embeddings: torch.Tensor,
             candidate_embeddings: torch.Tensor,
             sample_weight: Optional[torch.Tensor] = None,
             candidate_sampling_probability: Optional[torch.Tensor] = None,
             candidate_ids: Optional[torch.Tensor] = None,
             compute_metrics: bool = True) -> torch.Tensor:
        """Compute loss and metrics"""
        inv_tau = 1.0 if self._temperature is None else 1.0 / float(self._temperature)

        if self._num_hard_negatives is None and self._loss is None:
            loss = ops.InBatchSoftmax.apply(query_embeddings, candidate_embeddings, sample_weight,
                                            candidate_sampling_probability, candidate_ids, inv_tau)
        else:
            loss = self._materialised_loss(query_embeddings, candidate_embeddings, sample_weight,
                                           candidate_sampling_probability, candidate_ids)

        if compute_metrics is False or not self._factorized_metrics:
            return loss
        self._factorized_metrics.update_state(query_embeddings, candidate_embeddings)
        return loss

    def _materialised_loss(self, q, c, sample_weight, p, candidate_ids):
        """Hard-negative / custom-loss path: needs the score matrix (sbcnm.py:129-151 order)."""
        scores = _ScoresFn.apply(q, c, p, candidate_ids)                    # QC^T - log p + dup*MIN_FLOAT
        nq, nc = scores.shape
        labels = torch.eye(nq, nc, device=scores.device, dtype=scores.dtype)
        if self._num_hard_negatives is not None:
            num_sampled = min(self._num_hard_negatives + 1, nc)
            _, _, idx = ops.hard_negative_topk(scores.detach(), num_sampled)
            scores = _gather_elements_along_row(scores, idx)
            labels = _gather_elements_along_row(labels, idx)
        if self._temperature is not None:
            scores = scores / self._temperature
        if self._loss is not None:
            return self._loss(labels, scores, sample_weight)
        per_row = -(labels * torch.log_softmax(scores, dim=1)).sum(dim=1)
        if sample_weight is not None:
            per_row = per_row * sample_weight.reshape(-1)
        return per_row.sum()


class _ScoresFn(torch.autograd.Function):
    """Differentiable materialised scores (kernel forward; backward = two GEMMs through dr_dense_*)."""

    @staticmethod
    def forward(ctx, q, c, p, ids):
        ctx.save_for_backward(q, c)
        return ops.scores(q, c, p, ids)

    @staticmethod
    def backward(ctx, g):
        q, c = ctx.saved_tensors
        g = g.contiguous()
        # gQ = g @ C ; gC = g^T @ Q   (Dense kernels: x @ W with W = C / Q)
        gq = ops.DenseFn.apply(g, c, None, 0)
        gc = ops.DenseFn.apply(g.t().contiguous(), q, None, 0)
        return gq, gc, None, None


class TwoTower(Model):
    """User tower / item tower -> Retrieval task (BASELINE config C4; not in the reference).

    tower = single-table embedding gather (dr_gather_fwd) [+ optional DNN projection].
    """

    def __init__(self, num_users: int, num_items: int, dim: int = 64, tower_units: Sequence[int] = (),
                 temperature: Optional[float] = None, sparse_lr: Optional[float] = None,
                 seed: Optional[int] = None, device=None, **kwargs):
        super().__init__(**kwargs)
        dev = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
        std = 1.0 / (dim ** 0.5)
        gen = torch.Generator(device=dev).manual_seed(seed) if seed is not None else None

        def table(n):
            t = torch.empty((n, dim), dtype=torch.float32, device=dev)
            nn.init.trunc_normal_(t, 0.0, std, -2 * std, 2 * std, generator=gen)
            return nn.Parameter(t)

        self.user_table = table(num_users)
        self.item_table = table(num_items)
        self.user_dnn = DNN(list(tower_units[:-1]), out_units=tower_units[-1]) if tower_units else None
        self.item_dnn = DNN(list(tower_units[:-1]), out_units=tower_units[-1]) if tower_units else None
        self.task = Retrieval(temperature=temperature)
        self.sparse_lr = sparse_lr
        self.built = True

    def user_tower(self, user_ids):
        e = ops.Gather.apply(self.user_table, user_ids, self.sparse_lr)
        return self.user_dnn(e) if self.user_dnn is not None else e

    def item_tower(self, item_ids):
        e = ops.Gather.apply(self.item_table, item_ids, self.sparse_lr)
        return self.item_dnn(e) if self.item_dnn is not None else e

    def call(self, user_ids, item_ids, sample_weight=None, candidate_sampling_probability=None,
             remove_accidental_hits: bool = False):
        q = self.user_tower(user_ids)
        c = self.item_tower(item_ids)
        return self.task(q, c, sample_weight=sample_weight,
                         candidate_sampling_probability=candidate_sampling_probability,
                         candidate_ids=item_ids if remove_accidental_hits else None)
