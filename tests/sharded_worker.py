"""Launched by tests/test_gpu_sharded.py under torchrun (one rank per GPU)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from deep_recommenders_b200 import _lib
    for exchange in ("p2p", "nccl"):
        run(exchange, rank, world, dev)
    # wide rows (BASELINE C3-C5: D = 32 ... 128): rows of D floats, first-order weights trail the shard
    for D in (32, 128):
        run("p2p", rank, world, dev, D=D)
    # LINX lane mapping of the peer-memory forward (knob embed_fwd_linx_shard): same numbers
    _lib.tune("embed_fwd_linx_shard", 1)
    try:
        run("p2p", rank, world, dev)
    finally:
        _lib.tune("embed_fwd_linx_shard", 0)
    for temperature, accidental in ((None, False), (0.5, True)):
        run_two_tower(rank, world, dev, temperature, accidental)
    dist.barrier()
    if rank == 0:
        print("SHARDED_OK")
    dist.destroy_process_group()


def run(exchange, rank, world, dev, D=16):
    """Row-sharded DeepFM step on `world` GPUs == the unsharded step on the global batch."""
    from deep_recommenders_b200 import feature_column as fc
    from deep_recommenders_b200.keras.models.ranking import DeepFM
    from deep_recommenders_b200.sharded import ShardedDeepFMTrainStep
    from deep_recommenders_b200.training import DeepFMTrainStep
    rows, B = [3000, 7, 500, 1000, 21, 64], 1024
    cols = [fc.categorical_column_with_identity(f"c{i}", r) for i, r in enumerate(rows)]
    sh = ShardedDeepFMTrainStep(cols, D, [32, 8], batch_size=B, lr=0.05, seed=3, device=dev, exchange=exchange)
    total = sum(rows)
    # global arena, identical on all ranks (seeded), scattered into the shards
    g = torch.Generator(device=dev).manual_seed(11)
    arena_e = torch.zeros((total, D), device=dev).normal_(0, 0.25, generator=g)
    arena_l = torch.zeros((total,), device=dev).normal_(0, 0.1, generator=g)
    sh.emb.weight[:, :D].copy_(arena_e[rank::world])
    sh.emb.lin_view().copy_(arena_l[rank::world])
    torch.cuda.synchronize()
    dist.barrier()
    # unsharded twin on the GLOBAL batch (every rank computes it redundantly)
    model = DeepFM([fc.indicator_column(c) for c in cols], [fc.embedding_column(c, D) for c in cols],
                   dnn_units_size=[32, 8], seed=3, device=dev, sparse_lr=0.05)
    ref = DeepFMTrainStep(model, batch_size=B * world, lr=0.05, use_graph=False)
    with torch.no_grad():
        model.embeddings.emb_view().copy_(arena_e)
        model.embeddings.lin_view().copy_(arena_l)
        for i in range(len(ref.layers)):
            ref.w[i].copy_(sh.w[i])
            ref.b[i].copy_(sh.b[i])
    g2 = torch.Generator(device=dev).manual_seed(5)
    ids_g = torch.stack([torch.randint(-1, r + 1, (B * world,), device=dev, generator=g2) for r in rows], dim=1)
    lab_g = torch.randint(0, 2, (B * world,), device=dev, generator=g2).float()
    for it in range(3):
        l = sh.step(ids_g[rank * B:(rank + 1) * B], lab_g[rank * B:(rank + 1) * B]).clone()
        sh.check_overflow()
        dist.all_reduce(l)
        l_sh = float(l) / world
        l_ref = float(ref.step(ids_g, lab_g).item())
        assert abs(l_sh - l_ref) <= 2e-5 * abs(l_ref) + 1e-6, (exchange, D, l_sh, l_ref)
        mine = ref.stack[rank * B:(rank + 1) * B]
        if it == 0:    # same tables on both sides: the rows that travelled over NVLink are bit-exact copies
            assert torch.equal(sh.stack.view(B, -1), mine.view(B, -1)), (exchange, D, "gathered rows differ")
        assert torch.allclose(sh.stack.view(B, -1), mine.view(B, -1), rtol=1e-5, atol=1e-6)
    want_e = model.embeddings.emb_view()[rank::world]
    want_l = model.embeddings.lin_view()[rank::world]
    assert torch.allclose(sh.emb.weight[:, :D], want_e, rtol=1e-4, atol=1e-6), \
        (exchange, D, float((sh.emb.weight[:, :D] - want_e).abs().max()))
    assert torch.allclose(sh.emb.lin_view(), want_l, rtol=1e-4, atol=1e-6), (exchange, D, "first-order weights")
    for i in range(len(ref.layers)):
        assert torch.allclose(sh.w[i], ref.w[i], rtol=1e-4, atol=1e-6)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print(f"world {world} sharded DeepFM ok: exchange={exchange} D={D}", flush=True)


def run_two_tower(rank, world, dev, temperature, accidental):
    """Row-sharded two-tower step (sbcnm.py:120-151 on the GLOBAL batch) against the float64 oracle."""
    import numpy as np
    from oracle import reference_np as R
    from deep_recommenders_b200.sharded_two_tower import ShardedTwoTowerTrainStep
    U, I, D, b, lr = 300, 500, 64, 256, 0.05
    st = ShardedTwoTowerTrainStep(U, I, D, b, lr=lr, temperature=temperature, remove_accidental_hits=accidental,
                                  seed=2, device=dev)
    total = U + I

    def global_arena():
        mx = (total + world - 1) // world
        mine = torch.zeros((mx, D), device=dev)
        mine[:st.weight.shape[0]].copy_(st.weight)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        out = torch.zeros((mx * world, D), device=dev)
        for r in range(world):
            out[r::world] = parts[r]                        # global row = local * world + owner
        return out[:total].cpu().numpy().astype(np.float64)

    torch.cuda.synchronize()
    dist.barrier()
    arena = global_arena()
    rng = np.random.default_rng(0)                           # the same global batch on every rank
    for it in range(3):
        uid = rng.integers(0, U, b * world)
        iid = rng.integers(0, I, b * world)
        iid[: b // 8] = iid[0]                               # duplicates inside rank 0's block
        iid[b: b + b // 8] = iid[0]                          # ... and across ranks: cross-rank accidental hits
        sl = slice(rank * b, (rank + 1) * b)
        loss = st.step(torch.from_numpy(uid[sl]).to(dev), torch.from_numpy(iid[sl]).to(dev)).clone()
        dist.all_reduce(loss)
        Q, C = arena[uid], arena[U + iid]
        ids = iid if accidental else None
        ref_loss, _, _ = R.retrieval_loss(Q, C, candidate_ids=ids, temperature=temperature, dtype=np.float64)
        gq, gc = R.retrieval_grad(Q, C, candidate_ids=ids, temperature=temperature, dtype=np.float64)
        assert abs(float(loss) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss)) + 1e-4, (float(loss), float(ref_loss))
        if it == 0:
            assert np.array_equal(st.q.cpu().numpy(), Q[sl].astype(np.float32))      # gathered rows bit-exact
        np.add.at(arena, uid, -lr * gq)
        np.add.at(arena, U + iid, -lr * gc)
        torch.cuda.synchronize()
        dist.barrier()
        got = global_arena()
        tol = 1e-5 * (np.abs(arena).max() + lr * (np.abs(gq).max() + np.abs(gc).max()) * b / 8)
        assert np.abs(got - arena).max() <= tol, (np.abs(got - arena).max(), tol)
    if rank == 0:
        print(f"world {world} sharded two-tower ok: temperature={temperature} accidental={accidental}", flush=True)


if __name__ == "__main__":
    main()
