"""Step time of the other BASELINE configs through the public layer API (autograd path):
C3 = DCN 26 slots, D=32 (d=832), 3 full-matrix Cross + DNN[512,256,128], B=131072;
C4 = two-tower, 1M users / 10M items, D=64, in-batch softmax, B=16384.
fwd + bwd + SGD (row-sparse fused update for the tables, dr_sgd_step for dense parameters)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_recommenders_b200 import ops  # noqa: E402
from deep_recommenders_b200.keras.layers import DCN, TwoTower  # noqa: E402


def timed(step, n=5, warm=2):
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def dense_sgd(params, lr):
    with torch.no_grad():
        for p in params:
            if p.grad is not None:
                ops.sgd_step_(p, p.grad, lr)
                p.grad = None


def main():
    out = []
    gen = torch.Generator(device="cuda").manual_seed(0)
    # ---- C3 ----
    B, S, D = 131072, 26, 32
    model = DCN([1_000_000] * S, D, num_cross=3, dnn_units=[512, 256, 128], sparse_lr=0.01, seed=0, device="cuda")
    ids = [torch.randint(0, 1_000_000, (B, S), device="cuda", generator=gen) for _ in range(3)]
    y = torch.randint(0, 2, (B, 1), device="cuda", generator=gen).float()
    model.logits(ids[0])
    dense = [p for n, p in model.named_parameters() if not n.startswith("embeddings.")]
    it = [0]

    def step_c3():
        z = model.logits(ids[it[0] % 3])
        it[0] += 1
        loss, gz, _ = ops.bce_with_logits(z, y)
        z.backward(gz.view_as(z))
        dense_sgd(dense, 0.01)

    ms = timed(step_c3)
    flops = 3 * (3 * 2 * B * 832 * 832 + 2 * B * (832 * 512 + 512 * 256 + 256 * 128))
    r = dict(config="C3 DCN 26x1M D=32 B=131072, 3 Cross(832) + DNN[512,256,128]", ms_per_step=ms,
             examples_per_s=B / ms * 1e3, gemm_tflops_useful=flops / ms / 1e9)
    print(json.dumps(r), flush=True)
    out.append(r)
    del model, ids, dense
    torch.cuda.empty_cache()
    # ---- C4 ----
    B, D = 16384, 64
    tt = TwoTower(1_000_000, 10_000_000, dim=D, sparse_lr=0.01, seed=0, device="cuda")
    users = [torch.randint(0, 1_000_000, (B,), device="cuda", generator=gen) for _ in range(3)]
    items = [torch.randint(0, 10_000_000, (B,), device="cuda", generator=gen) for _ in range(3)]

    def step_c4():
        k = it[0] % 3
        it[0] += 1
        loss = tt(users[k], items[k])
        loss.backward()

    ms = timed(step_c4)
    r = dict(config="C4 two-tower 1M users / 10M items D=64 B=16384, in-batch softmax", ms_per_step=ms,
             examples_per_s=B / ms * 1e3, softmax_tflops_useful=(2 * B * B * D) * 5 / ms / 1e9)
    print(json.dumps(r), flush=True)
    out.append(r)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/bench_configs.json", "w"), indent=1)


if __name__ == "__main__":
    main()
