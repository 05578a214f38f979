"""ORACLE (timing twin) -- test infrastructure only.  A torch-CPU restatement of the reference's
DeepFM train step, used ONLY by bench.py's `cpu_baseline` / `--impl reference` legs and by
tests.  It is "the reference's own CPU implementation of the path" in the only form that can
run here: TensorFlow is not installed in this image, so this is NOT TensorFlow -- it follows the
reference op for op on torch's CPU kernels (MKL/oneDNN), with all host threads:

  per-column embedding lookup (DenseFeatures per key, keras deepfm.py:39-43)  -> F.embedding, sparse grads
  tf.stack / tf.concat (deepfm.py:44-45)                                       -> torch.stack / reshape
  FM.call (fm.py:23-37)                                                        -> sum / pow / sub ops
  Sequential(Dense(relu)..., Dense(1)) (deepfm.py:30-34)                       -> torch.nn.Linear
  sigmoid + binary_crossentropy (examples/train_deepfm_on_movielens_keras.py:43) -> BCE
  backward: autograd (the reference defines no custom gradient); IndexedSlices -> sparse grads
  optimizer: SGD, row-sparse on the tables (same update rule as the GPU step it is compared with)

One stated deviation: the reference's first-order term densifies the
indicator columns to [B, sum N] (6.8 TB at C2) and cannot run at this size; the mathematically
identical gather-sum is used.
"""
from __future__ import annotations

import time

import torch
import torch.nn.functional as F


class DeepFMCPU(torch.nn.Module):
    def __init__(self, rows, dim, dnn_units, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        std = 1.0 / dim ** 0.5
        self.tables = torch.nn.ParameterList(
            [torch.nn.Parameter(torch.nn.init.trunc_normal_(torch.empty(r, dim), 0, std, -2 * std, 2 * std, generator=g))
             for r in rows])
        self.linear = torch.nn.ParameterList([torch.nn.Parameter(torch.zeros(r, 1)) for r in rows])
        self.bias = torch.nn.Parameter(torch.zeros(1))
        dims = [len(rows) * dim] + list(dnn_units) + [1]
        self.dnn = torch.nn.ModuleList([torch.nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1)])

    def forward(self, ids):
        embs, lin = [], self.bias
        for s, t in enumerate(self.tables):
            col = ids[:, s]
            embs.append(F.embedding(col, t, sparse=True))
            lin = lin + F.embedding(col, self.linear[s], sparse=True)[:, 0]
        stack = torch.stack(embs, dim=1)
        x_sum = stack.sum(dim=1)
        x_square_sum = stack.pow(2).sum(dim=1)
        interaction = 0.5 * (x_sum.pow(2) - x_square_sum).sum(dim=1)
        h = stack.reshape(stack.shape[0], -1)
        for i, l in enumerate(self.dnn):
            h = l(h)
            if i < len(self.dnn) - 1:
                h = torch.relu(h)
        return lin + interaction + h[:, 0]

    def train_step(self, ids, labels, lr):
        for p in self.parameters():
            p.grad = None
        logit = self.forward(ids)
        loss = F.binary_cross_entropy_with_logits(logit, labels)
        loss.backward()
        with torch.no_grad():
            for p in self.parameters():
                if p.grad is not None:
                    p.add_(p.grad, alpha=-lr)
        return float(loss)


def time_deepfm_cpu(rows, dim, dnn_units, batch, steps, warmup=1, lr=0.01, seed=0, max_seconds=60.0):
    """Returns dict(examples_per_sec, seconds_per_step, steps, cores).  Bounded by max_seconds."""
    model = DeepFMCPU(rows, dim, dnn_units, seed)
    g = torch.Generator().manual_seed(seed + 1)
    pool = [(torch.stack([torch.randint(0, r, (batch,), generator=g) for r in rows], dim=1),
             torch.randint(0, 2, (batch,), generator=g).float()) for _ in range(2)]
    for i in range(warmup):
        model.train_step(*pool[i % 2], lr)
    t0 = time.perf_counter()
    done = 0
    for i in range(steps):
        model.train_step(*pool[i % 2], lr)
        done += 1
        if time.perf_counter() - t0 > max_seconds:
            break
    dt = time.perf_counter() - t0
    return dict(examples_per_sec=batch * done / dt, seconds_per_step=dt / done, steps=done,
                cores=torch.get_num_threads())
