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
    from deep_recommenders_b200 import feature_column as fc
    from deep_recommenders_b200.keras.models.ranking import DeepFM
    from deep_recommenders_b200.sharded import ShardedDeepFMTrainStep
    from deep_recommenders_b200.training import DeepFMTrainStep
    for exchange in ("p2p", "nccl"):
        run(exchange, rank, world, dev)
    dist.barrier()
    if rank == 0:
        print("SHARDED_OK")
    dist.destroy_process_group()


def run(exchange, rank, world, dev):
    from deep_recommenders_b200 import feature_column as fc
    from deep_recommenders_b200.keras.models.ranking import DeepFM
    from deep_recommenders_b200.sharded import ShardedDeepFMTrainStep
    from deep_recommenders_b200.training import DeepFMTrainStep
    S, D, rows, B = 6, 16, [3000, 7, 500, 1000, 21, 64], 1024
    cols = [fc.categorical_column_with_identity(f"c{i}", r) for i, r in enumerate(rows)]
    sh = ShardedDeepFMTrainStep(cols, D, [32, 8], batch_size=B, lr=0.05, seed=3, device=dev, exchange=exchange)
    total = sum(rows)
    # global arena, identical on all ranks (seeded), scattered into the shards
    g = torch.Generator(device=dev).manual_seed(11)
    arena = torch.zeros((total, D + 4), device=dev)
    arena[:, :D].normal_(0, 0.25, generator=g)
    arena[:, D].normal_(0, 0.1, generator=g)
    sh.emb.weight[:, :D + 4].copy_(arena[rank::world])
    dist.barrier()
    # unsharded twin on the GLOBAL batch (every rank computes it redundantly)
    model = DeepFM([fc.indicator_column(c) for c in cols], [fc.embedding_column(c, D) for c in cols],
                   dnn_units_size=[32, 8], seed=3, device=dev, sparse_lr=0.05)
    ref = DeepFMTrainStep(model, batch_size=B * world, lr=0.05, use_graph=False)
    with torch.no_grad():
        model.embeddings.emb_view().copy_(arena[:, :D])
        model.embeddings.lin_view().copy_(arena[:, D])
        for i in range(len(ref.layers)):
            ref.w[i].copy_(sh.w[i])
            ref.b[i].copy_(sh.b[i])
    g2 = torch.Generator(device=dev).manual_seed(5)
    ids_g = torch.stack([torch.randint(-1, r + 1, (B * world,), device=dev, generator=g2) for r in rows], dim=1)
    lab_g = torch.randint(0, 2, (B * world,), device=dev, generator=g2).float()
    for _ in range(3):
        l = sh.step(ids_g[rank * B:(rank + 1) * B], lab_g[rank * B:(rank + 1) * B]).clone()
        sh.check_overflow()
        dist.all_reduce(l)
        l_sh = float(l) / world
        l_ref = float(ref.step(ids_g, lab_g).item())
        assert abs(l_sh - l_ref) <= 2e-5 * abs(l_ref) + 1e-6, (l_sh, l_ref)
        assert torch.allclose(sh.stack, ref.stack[rank * B:(rank + 1) * B], rtol=1e-5, atol=1e-6)
    want = model.embeddings.weight[rank::world, :D + 1]
    assert torch.allclose(sh.emb.weight[:, :D + 1], want, rtol=1e-4, atol=1e-6), float((sh.emb.weight[:, :D + 1] - want).abs().max())
    for i in range(len(ref.layers)):
        assert torch.allclose(sh.w[i], ref.w[i], rtol=1e-4, atol=1e-6)
    torch.cuda.synchronize()
    dist.barrier()


if __name__ == "__main__":
    main()
