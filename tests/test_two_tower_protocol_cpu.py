"""world_size-2 gloo test (CPU) of the exchange plan of deep_recommenders_b200/sharded_two_tower.py (BASELINE config
C4 at N > 1): sharded gather of user / item rows -> all-gather of the candidate embeddings -> own block first ->
per-rank in-batch softmax -> gradient blocks back to rank order -> reduce-scatter -> owner-side row updates.
Block ordering and row ownership come from the product's shard_plan.py; the per-rank arithmetic is the numpy oracle.
The composition must equal the unsharded oracle on the GLOBAL batch (loss SUM over ranks, every gradient row)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deep_recommenders_b200 import shard_plan
from oracle import reference_np as R


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem(world):
    rng = np.random.default_rng(5)
    U, I, D, b = 13, 17, 8, 6
    users = rng.standard_normal((U, D))
    items = rng.standard_normal((I, D))
    uid = rng.integers(0, U, world * b)
    iid = rng.integers(0, I, world * b)
    iid[3] = iid[world * b - 2]                       # an accidental hit across ranks
    uid[1] = uid[4]                                   # duplicate rows inside one rank's batch
    return U, I, D, b, users, items, uid, iid


def _worker(rank, world, port, q, temperature, accidental):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        U, I, D, b, users, items, uid_g, iid_g = _problem(world)
        lr = 0.1
        arena = np.concatenate([users, items], 0)                        # one global row space: users, then items
        total = U + I
        mine = np.arange(rank, total, world)
        assert len(mine) == shard_plan.local_rows(total, rank, world)
        shard = arena[mine].copy()
        # "peer memory": every rank can read every shard (here: all-gather of the small shards, padded)
        max_rows = shard_plan.local_rows(total, 0, world)
        pad = np.zeros((max_rows, D))
        pad[:len(mine)] = shard
        peers = [torch.empty((max_rows, D), dtype=torch.float64) for _ in range(world)]
        dist.all_gather(peers, torch.from_numpy(pad))

        def read(global_rows):
            return np.stack([peers[shard_plan.owner(int(r), world)][shard_plan.local_row(int(r), world)].numpy()
                             for r in global_rows])

        uid, iid = uid_g[rank * b:(rank + 1) * b], iid_g[rank * b:(rank + 1) * b]
        qv, c_local = read(uid), read(U + iid)
        assert np.array_equal(qv, users[uid]) and np.array_equal(c_local, items[iid])
        blocks = [torch.empty((b, D), dtype=torch.float64) for _ in range(world)]
        dist.all_gather(blocks, torch.from_numpy(c_local))
        idblocks = [torch.empty((b,), dtype=torch.int64) for _ in range(world)]
        dist.all_gather(idblocks, torch.from_numpy(iid))
        order = shard_plan.own_first_order(rank, world)
        inv = shard_plan.inverse_order(order)
        assert order[0] == rank and sorted(order) == list(range(world)) and [order[p] for p in inv] == list(range(world))
        c_rot = np.concatenate([blocks[r].numpy() for r in order], 0)
        ids_rot = np.concatenate([idblocks[r].numpy() for r in order], 0) if accidental else None
        loss, _, _ = R.retrieval_loss(qv, c_rot, candidate_ids=ids_rot, temperature=temperature, dtype=np.float64)
        gq, gc_rot = R.retrieval_grad(qv, c_rot, candidate_ids=ids_rot, temperature=temperature, dtype=np.float64)
        gc_blocks = gc_rot.reshape(world, b, D)
        gc_all = np.stack([gc_blocks[inv[r]] for r in range(world)], 0)  # back to rank order
        t = torch.from_numpy(gc_all.copy())
        dist.all_reduce(t)                                               # reduce-scatter = all-reduce + own block
        gc_local = t.numpy()[rank]
        tl = torch.tensor([float(loss)], dtype=torch.float64)
        dist.all_reduce(tl)

        # ---- expected: the unsharded oracle on the GLOBAL batch -----------------------------------------------
        Q, C = users[uid_g], items[iid_g]
        ref_loss, _, _ = R.retrieval_loss(Q, C, candidate_ids=iid_g if accidental else None, temperature=temperature,
                                          dtype=np.float64)
        ref_gq, ref_gc = R.retrieval_grad(Q, C, candidate_ids=iid_g if accidental else None, temperature=temperature,
                                          dtype=np.float64)
        assert np.allclose(float(tl), float(ref_loss), rtol=1e-12)
        assert np.allclose(gq, ref_gq[rank * b:(rank + 1) * b], rtol=1e-10, atol=1e-12)
        assert np.allclose(gc_local, ref_gc[rank * b:(rank + 1) * b], rtol=1e-10, atol=1e-12)

        # ---- owner-side updates: every rank sends its row gradients to the owners (emulated: all-gather, filter) ---
        contrib_rows = np.concatenate([uid, U + iid])
        contrib_vals = np.concatenate([gq, gc_local], 0)
        rows_l = [torch.empty((2 * b,), dtype=torch.int64) for _ in range(world)]
        vals_l = [torch.empty((2 * b, D), dtype=torch.float64) for _ in range(world)]
        dist.all_gather(rows_l, torch.from_numpy(contrib_rows))
        dist.all_gather(vals_l, torch.from_numpy(contrib_vals))
        for rr, vv in zip(rows_l, vals_l):
            rr, vv = rr.numpy(), vv.numpy()
            own = rr % world == rank
            np.add.at(shard, rr[own] // world, -lr * vv[own])
        full = arena.copy()
        np.add.at(full, uid_g, -lr * ref_gq)
        np.add.at(full, U + iid_g, -lr * ref_gc)
        assert np.allclose(shard, full[mine], rtol=1e-10, atol=1e-12)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, repr(e) + traceback.format_exc()[-600:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("temperature,accidental", [(None, False), (0.5, True)])
def test_two_tower_exchange_protocol_gloo_world2(temperature, accidental):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, temperature, accidental)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_block_order_helpers():
    for world in (1, 2, 3, 8):
        for rank in range(world):
            o = shard_plan.own_first_order(rank, world)
            inv = shard_plan.inverse_order(o)
            assert o[0] == rank and sorted(o) == list(range(world))
            assert all(o[inv[r]] == r for r in range(world))
