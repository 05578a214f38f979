"""world_size-2 gloo test (CPU) of the N>1 exchange protocol of deep_recommenders_b200/sharded.py:
ids out (padded equal-split all-to-all) -> owners gather -> vectors back -> gradients out ->
owners scatter-add; tower gradients all-reduced.  Partition arithmetic comes from the product's
shard_plan.py; the per-rank compute is the numpy oracle.  The result must equal the unsharded oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deep_recommenders_b200 import shard_plan
from oracle import reference_np as R
from oracle import shard_np


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    rng = np.random.default_rng(0)
    rows, D, Bg = [37, 5, 64, 11], 8, 24            # global batch 24 = 2 ranks x 12
    tables = [rng.standard_normal((r, D)) for r in rows]
    lins = [rng.standard_normal(r) for r in rows]
    ids = np.stack([rng.integers(-1, r + 1, Bg) for r in rows], axis=1)
    g_logit = rng.standard_normal(Bg)
    g_stack = rng.standard_normal((Bg, len(rows), D))
    return rows, D, tables, lins, ids, g_logit, g_stack


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows, D, tables, lins, ids_g, gl_g, gs_g = _problem()
        S, V = len(rows), D + 4
        offs = shard_plan.slot_offsets(rows)
        total = sum(rows)
        # this rank's shard of the fused global arena [emb | w | pad]
        arena = np.zeros((total, V))
        for s in range(S):
            arena[offs[s]:offs[s] + rows[s], :D] = tables[s]
            arena[offs[s]:offs[s] + rows[s], D] = lins[s]
        mine = np.arange(rank, total, world)
        assert len(mine) == shard_plan.local_rows(total, rank, world)
        local = arena[mine].copy()
        Bl = ids_g.shape[0] // world
        ids = ids_g[rank * Bl:(rank + 1) * Bl]
        cap = shard_plan.capacity(Bl * S, world, slack=1.0, floor=4)
        send, inv, counts, overflow = shard_np.bucket_ids(ids, offs, rows, world, cap)
        assert not overflow
        recv = torch.empty(world * cap, dtype=torch.int64)
        dist.all_to_all_single(recv, torch.from_numpy(send))
        rid = recv.numpy()
        vec = np.zeros((world * cap, V))
        ok = (rid >= 0) & (rid < len(mine))
        vec[ok] = local[rid[ok]]
        back = torch.empty((world * cap, V), dtype=torch.float64)
        dist.all_to_all_single(back, torch.from_numpy(vec))
        rowsv = np.zeros((Bl * S, V))
        okv = inv >= 0
        rowsv[okv] = back.numpy()[inv[okv]]
        stack = rowsv[:, :D].reshape(Bl, S, D)
        lin = rowsv[:, D].reshape(Bl, S).sum(1)
        ref_logit, ref_stack = R.fm_logit(tables, lins, 0.0, ids, np.float64)
        assert np.array_equal(stack, ref_stack)
        assert np.allclose(lin.reshape(-1, 1) + R.fm_second_order(stack, np.float64), ref_logit)
        # backward: per-lookup gradient rows -> owners
        gl, gs = gl_g[rank * Bl:(rank + 1) * Bl], gs_g[rank * Bl:(rank + 1) * Bl]
        dE = gl[:, None, None] * (stack.sum(1, keepdims=True) - stack) + gs
        grow = np.zeros((Bl * S, V))
        grow[:, :D] = dE.reshape(-1, D)
        grow[:, D] = np.repeat(gl, S)
        gsend = np.zeros((world * cap, V))
        gsend[inv[okv]] = grow[okv]
        grecv = torch.empty((world * cap, V), dtype=torch.float64)
        dist.all_to_all_single(grecv, torch.from_numpy(gsend))
        glocal = np.zeros_like(local)
        np.add.at(glocal, rid[ok], grecv.numpy()[ok])
        # expected: the unsharded oracle gradient of the GLOBAL batch, restricted to my rows
        gts, gls, gb = R.embed_fm_grad(rows, ids_g, R.stack_embeddings(tables, ids_g), gl_g, gs_g, np.float64)
        garena = np.zeros((total, V))
        for s in range(S):
            garena[offs[s]:offs[s] + rows[s], :D] = gts[s]
            garena[offs[s]:offs[s] + rows[s], D] = gls[s]
        assert np.allclose(glocal, garena[mine])
        # tower gradient all-reduce (bias grad as the stand-in)
        t = torch.tensor([gl.sum()])
        dist.all_reduce(t)
        assert np.allclose(float(t), gb)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_sharded_exchange_protocol_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_plan_arithmetic():
    for total in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            assert sum(shard_plan.local_rows(total, r, world) for r in range(world)) == total
    assert shard_plan.slot_offsets([3, 4, 5]) == [0, 3, 7]
    assert shard_plan.owner(13, 8) == 5 and shard_plan.local_row(13, 8) == 1
    assert shard_plan.capacity(1000, 1) == 1000
    assert shard_plan.capacity(1_703_936, 8) >= 1_703_936 // 8 + 1024
