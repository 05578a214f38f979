"""Controlled experiments on the fused gather+FM forward/backward at config C2 (developer tool).

Each line answers one question about where the time goes: first-order gathers, stacked write,
table layout (fused row = [emb | w | pad] vs split arrays), id width, L2-resident tables.
Writes gpurun_out/probe_embed.json.  `--ncu` runs only the default fwd+bwd a few times (for an
`ncu --set full` capture).
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_recommenders_b200 import _lib  # noqa: E402
from deep_recommenders_b200.embedding import EmbeddingCollection  # noqa: E402

HBM = 6480.5
try:
    HBM = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, nrep=30, warm=6):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(nrep):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / nrep * 1e-3


def main():
    ncu_mode = "--ncu" in sys.argv
    lib = _lib.load()
    B, S = 65536, 26
    st = torch.cuda.current_stream().cuda_stream
    out = []
    NP = 6
    gran = 0
    for a in sys.argv:
        if a.startswith("--gran="):
            gran = int(a.split("=")[1])
    if gran:
        _lib.tune("l2_fetch_granularity", gran)
    for D, rows, layout, idb in ((16, 1_000_000, "fused", 8), (16, 1_000_000, "split", 8),
                                 (16, 20_000, "split", 8), (32, 1_000_000, "fused", 8), (32, 1_000_000, "split", 8),
                                 (128, 200_000, "split", 8)):
        if ncu_mode and not (D == 16 and rows == 1_000_000 and layout == "fused" and idb == 8):
            continue
        coll = EmbeddingCollection([rows] * S, D, device="cuda", seed=1, layout=layout)
        gen = torch.Generator(device="cuda").manual_seed(0)
        ids_pool = [torch.randint(0, rows, (B, S), device="cuda", generator=gen,
                                  dtype=torch.int64 if idb == 8 else torch.int32) for _ in range(NP)]
        stacks = [torch.empty((B, S, D), device="cuda") for _ in range(NP)]
        sums = [torch.empty((B, D), device="cuda") for _ in range(NP)]
        logits = [torch.empty((B,), device="cuda") for _ in range(NP)]
        gl = torch.randn(B, device="cuda") * 1e-3
        gs = [torch.randn((B, S, D), device="cuda") * 1e-3 for _ in range(NP)]
        tp, lp, rws = coll.pointers(coll.weight, coll.linear)

        def fwd(i, lin=True, stack=True, logit=True, sume=True):
            k = i % NP
            _lib.check(lib.dr_embed_fm_fwd(tp.data_ptr(), lp.data_ptr() if lin else None, rws.data_ptr(),
                                           ids_pool[k].data_ptr(), idb, coll.bias.data_ptr(), B, S, D,
                                           coll.row_stride, coll.lin_stride, coll.flags if lin else 0,
                                           stacks[k].data_ptr() if stack else None,
                                           sums[k].data_ptr() if sume else None,
                                           logits[k].data_ptr() if logit else None, st), "fwd")

        def bwd(i, fm=True, gstack=True, lin=True):
            k = i % NP
            _lib.check(lib.dr_embed_fm_bwd(ids_pool[k].data_ptr(), idb, rws.data_ptr(), stacks[k].data_ptr(),
                                           sums[k].data_ptr(), gl.data_ptr() if fm else None,
                                           gs[k].data_ptr() if gstack else None, B, S, D, coll.row_stride,
                                           coll.lin_stride, coll.flags if lin else 0, tp.data_ptr(), lp.data_ptr() if lin else None,
                                           coll.bias.data_ptr(), -1e-6, st), "bwd")

        if ncu_mode:
            for i in range(4):
                fwd(i)
            for i in range(4):
                bwd(i)
            torch.cuda.synchronize()
            return
        alg_f = B * (S * (idb + 4 * D + 4) + 4 * S * D + 4 * D + 4)
        alg_b = B * (S * idb + 4 * S * D * 2 + 4 * D + 4 + 4 * S * D + 4 * S)
        exps = {
            "fwd_full": (lambda i: fwd(i), alg_f),
            "fwd_no_lin": (lambda i: fwd(i, lin=False), alg_f - B * S * 4),
            "fwd_no_stack": (lambda i: fwd(i, stack=False), alg_f - B * 4 * S * D),
            "fwd_gather_only": (lambda i: fwd(i, lin=False, logit=False, sume=False), B * (S * (idb + 8 * D))),
            "bwd_full": (lambda i: bwd(i), alg_b),
            "bwd_no_lin": (lambda i: bwd(i, lin=False), alg_b - B * S * 4),
            "bwd_gstack_only": (lambda i: bwd(i, fm=False), B * (S * idb + 8 * S * D)),
        }
        for i in range(NP):
            fwd(i)
        for mode, unroll in ((0, 1), (0, 2), (0, 4), (1, 4)):
            _lib.tune("embed_bwd_mode", mode)
            _lib.tune("embed_bwd_unroll", unroll)
            t = timeit(lambda i: bwd(i))
            r = dict(exp=f"bwd_full_mode{mode}_u{unroll}", gran=gran, D=D, rows=rows, layout=layout, id_bytes=idb,
                     us=round(t * 1e6, 2), alg_gbs=round(alg_b / t / 1e9, 1), frac=round(alg_b / t / 1e9 / HBM, 4))
            print(json.dumps(r), flush=True)
            out.append(r)
        _lib.tune("embed_bwd_mode", 0)
        _lib.tune("embed_bwd_unroll", 0)
        for name, (fn, nbytes) in exps.items():
            t = timeit(fn)
            r = dict(exp=name, gran=gran, D=D, rows=rows, layout=layout, id_bytes=idb, us=round(t * 1e6, 2),
                     alg_gbs=round(nbytes / t / 1e9, 1), frac=round(nbytes / t / 1e9 / HBM, 4))
            print(json.dumps(r), flush=True)
            out.append(r)
        del coll, ids_pool, stacks, sums, gs
        torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/probe_embed_g{gran}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
