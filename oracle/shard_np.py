"""ORACLE -- test infrastructure only.  numpy twin of the padded equal-split exchange plan that
dr_shard_bucket_ids computes on the GPU (deep_recommenders_b200/csrc/shard.cu), used to test the
multi-rank PROTOCOL on CPU with the gloo backend (the reference itself is single-process)."""
from __future__ import annotations

import numpy as np


def bucket_ids(ids: np.ndarray, slot_offsets, rows, world: int, cap: int):
    """ids [B,S] -> (send_ids [world*cap] int64 (-1 padded), inv [B*S] int32, counts [world], overflow)."""
    B, S = ids.shape
    flat = ids.reshape(-1).astype(np.int64)
    slot = np.arange(B * S) % S
    rws = np.asarray(rows, np.int64)[slot]
    ok = (flat >= 0) & (flat < rws)
    grow = np.where(ok, np.asarray(slot_offsets, np.int64)[slot] + flat, -1)
    owner = np.where(grow < 0, 0, grow % world)
    send = np.full((world * cap,), -1, np.int64)
    inv = np.full((B * S,), -1, np.int32)
    counts = np.zeros((world,), np.int64)
    overflow = False
    for i in range(B * S):
        if grow[i] < 0:          # out-of-vocabulary: zero row, nothing to exchange
            continue
        g = owner[i]
        pos = counts[g]
        counts[g] += 1
        if pos < cap:
            send[g * cap + pos] = grow[i] // world
            inv[i] = g * cap + pos
        else:
            overflow = True
    return send, inv, counts, overflow
