"""A/B of the fused gather+FM forward variants at config C2 (developer tool, one short GPU run).

Variants = knobs of libdeeprec_b200.so that change scheduling only (results must stay bit-identical):
  embed_fwd_minblocks 0/3/4   register cap -> 2 / 3 / 4 CTAs of 256 threads per SM (ncu r01: 108 regs, 24 % warps active)
  embed_fwd_linx 0/1          lane group sized for the embedding chunks, lane 0 also fetches the in-row weight
  embed_fwd_unroll 8/13       independent row loads in flight per lane
  embed_ctas_per_sm           grid = one resident wave instead of 8 CTAs / SM
Each variant: outputs compared bit-for-bit with the default, then CUDA-event timing over rotating id / output
buffers (tables 3.3 GB >> L2).  Writes gpurun_out/ab_embed_fwd.json.
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


def main():
    lib = _lib.load()
    B, S, D, rows, NP = 65536, 26, 16, 1_000_000, 6
    st = torch.cuda.current_stream().cuda_stream
    coll = EmbeddingCollection([rows] * S, D, device="cuda", seed=1, layout="fused")
    with torch.no_grad():
        coll.lin_view().normal_(0, 0.1)
    gen = torch.Generator(device="cuda").manual_seed(0)
    ids = [torch.randint(0, rows, (B, S), device="cuda", generator=gen) for _ in range(NP)]
    stacks = [torch.empty((B, S, D), device="cuda") for _ in range(NP)]
    sums = [torch.empty((B, D), device="cuda") for _ in range(NP)]
    logits = [torch.empty((B,), device="cuda") for _ in range(NP)]
    tp, lp, rws = coll.pointers(coll.weight, coll.linear)

    def fwd(i):
        k = i % NP
        _lib.check(lib.dr_embed_fm_fwd(tp.data_ptr(), lp.data_ptr(), rws.data_ptr(), ids[k].data_ptr(), 8,
                                       coll.bias.data_ptr(), B, S, D, coll.row_stride, coll.lin_stride, coll.flags,
                                       stacks[k].data_ptr(), sums[k].data_ptr(), logits[k].data_ptr(), st), "fwd")

    def knobs(minb=0, linx=0, unroll=0, cps=0):
        _lib.tune("embed_fwd_minblocks", minb)
        _lib.tune("embed_fwd_linx", linx)
        _lib.tune("embed_fwd_unroll", unroll)
        _lib.tune("embed_ctas_per_sm", cps)

    def timeit(nrep=40, warm=6):
        for i in range(warm):
            fwd(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(nrep):
            fwd(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / nrep * 1e-3

    knobs(linx=0)
    fwd(0)
    torch.cuda.synchronize()
    ref = (stacks[0].clone(), sums[0].clone(), logits[0].clone())
    alg = B * (S * (8 + 4 * D + 4) + 4 * S * D + 4 * D + 4)
    out = []
    variants = [dict(minb=m, linx=l, unroll=u, cps=c)
                for l, m, u, c in itertools.product((0, 1), (0, 3, 4), (8, 13), (0,))]
    variants += [dict(minb=0, linx=0, unroll=8, cps=2), dict(minb=0, linx=1, unroll=8, cps=2)]
    for v in variants:
        knobs(**v)
        stacks[0].zero_(); sums[0].zero_(); logits[0].zero_()
        fwd(0)
        torch.cuda.synchronize()
        same = bool(torch.equal(stacks[0], ref[0]) and torch.equal(sums[0], ref[1]) and torch.equal(logits[0], ref[2]))
        t = min(timeit(), timeit())
        r = dict(v, us=round(t * 1e6, 2), alg_gbs=round(alg / t / 1e9, 1), frac=round(alg / t / 1e9 / HBM, 4),
                 bit_identical=same)
        print(json.dumps(r), flush=True)
        out.append(r)
    knobs(linx=1)      # library default
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/ab_embed_fwd.json", "w"), indent=1)


if __name__ == "__main__":
    main()
