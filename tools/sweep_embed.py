"""Developer sweep of the fused gather+FM kernels' tuning knobs on one B200 (not a bench).

Times dr_embed_fm_fwd / dr_embed_fm_bwd at BASELINE config C2 (B=65536, S=26, D=16, 1M-row
tables) and C3-like D=32 for every knob combination, with CUDA events around NREP launches that
cycle through a pool of id batches and output buffers (working set >> 126 MB L2).  Prints one
JSON line per variant and writes gpurun_out/sweep_embed.json.
"""
import itertools
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


def time_launches(fn, nrep=20, warm=5):
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
    lib = _lib.load()
    B, S, rows = 65536, 26, 1_000_000
    results = []
    for D, idb in ((16, 8), (16, 4), (32, 8)):
        coll = EmbeddingCollection([rows] * S, D, device="cuda", seed=1)
        with torch.no_grad():
            coll.lin_view().normal_(0, 0.1)
        gen = torch.Generator(device="cuda").manual_seed(0)
        NP = 6
        ids_pool = [torch.randint(0, rows, (B, S), device="cuda", generator=gen,
                                  dtype=torch.int64 if idb == 8 else torch.int32) for _ in range(NP)]
        stacks = [torch.empty((B, S, D), device="cuda") for _ in range(NP)]
        sums = [torch.empty((B, D), device="cuda") for _ in range(NP)]
        logits = [torch.empty((B,), device="cuda") for _ in range(NP)]
        tp, lp, rws = coll.pointers(coll.weight, coll.linear)
        st = torch.cuda.current_stream().cuda_stream
        fwd_bytes = B * (S * (idb + 4 * D + 4) + 4 * S * D + 4 * D + 4)
        bwd_bytes = B * (S * idb + 4 * S * D * 2 + 4 * D + 4 + 4 * S * D + 4 * S)

        def fwd(i):
            k = i % NP
            _lib.check(lib.dr_embed_fm_fwd(tp.data_ptr(), lp.data_ptr(), rws.data_ptr(), ids_pool[k].data_ptr(), idb,
                                           coll.bias.data_ptr(), B, S, D, coll.row_stride, coll.lin_stride, coll.flags, stacks[k].data_ptr(), sums[k].data_ptr(),
                                           logits[k].data_ptr(), st), "fwd")

        gl = torch.randn(B, device="cuda") * 1e-3
        gs = [torch.randn((B, S, D), device="cuda") * 1e-3 for _ in range(NP)]

        def bwd(i):
            k = i % NP
            _lib.check(lib.dr_embed_fm_bwd(ids_pool[k].data_ptr(), idb, rws.data_ptr(), stacks[k].data_ptr(),
                                           sums[k].data_ptr(), gl.data_ptr(), gs[k].data_ptr(), B, S, D, coll.row_stride, coll.lin_stride, coll.flags,
                                           tp.data_ptr(), lp.data_ptr(), coll.bias.data_ptr(), -1e-6, st), "bwd")

        for block, unroll, cps in itertools.product((128, 256, 512), (2, 4, 8, 13, 26), (0,)):
            _lib.tune("embed_block", block)
            _lib.tune("embed_fwd_unroll", unroll)
            _lib.tune("embed_ctas_per_sm", cps)
            t = time_launches(fwd)
            r = dict(kernel="fwd", D=D, id_bytes=idb, block=block, unroll=unroll, ctas_per_sm=cps, us=t * 1e6,
                     gbs=fwd_bytes / t / 1e9, frac=fwd_bytes / t / 1e9 / HBM)
            print(json.dumps(r), flush=True)
            results.append(r)
        _lib.tune("embed_fwd_unroll", 0)
        for i in range(NP):
            fwd(i)
        for block, unroll, agg in itertools.product((128, 256, 512), (1, 2, 4, 8), (0, 1)):
            _lib.tune("embed_block", block)
            _lib.tune("embed_bwd_unroll", unroll)
            _lib.tune("embed_bwd_agg", agg)
            t = time_launches(bwd)
            r = dict(kernel="bwd", D=D, id_bytes=idb, block=block, unroll=unroll, agg=agg, us=t * 1e6,
                     gbs=bwd_bytes / t / 1e9, frac=bwd_bytes / t / 1e9 / HBM)
            print(json.dumps(r), flush=True)
            results.append(r)
        _lib.tune("embed_block", 256)
        _lib.tune("embed_bwd_unroll", 0)
        _lib.tune("embed_bwd_agg", 1)
        del coll, ids_pool, stacks, sums, gs
        torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(results, open("gpurun_out/sweep_embed.json", "w"), indent=1)


if __name__ == "__main__":
    main()
