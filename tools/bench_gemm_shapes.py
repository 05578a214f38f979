#!/usr/bin/env python
"""Times the three GEMMs of one Dense layer (forward, dX, dW) through the C-ABI at the tower shapes of C2 / C5, per rank
batch M, under the tile-width knob.  Buffers rotate (3 sets) so the operands come from HBM as in a train step.
usage: bench_gemm_shapes.py [KEY=VALUE ...]   -> one JSON line per shape"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_recommenders_b200 import _lib
from deep_recommenders_b200._lib import check

knobs = [a for a in sys.argv[1:] if "=" in a]
for kv in knobs:
    k, v = kv.split("=")
    _lib.tune(k, int(v))
lib = _lib.load()
dev = torch.device("cuda", 0)
SHAPES = [(65536, 416, 256), (8192, 3328, 512), (16384, 3328, 512), (65536, 3328, 512), (8192, 512, 256)]
R = 3


def timed(fn, iters=6):
    for i in range(2):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for M, K, N in SHAPES:
    st = torch.cuda.current_stream().cuda_stream
    xs = [torch.randn(M, K, device=dev) for _ in range(R)]
    gys = [torch.randn(M, N, device=dev) for _ in range(R)]
    ys = [torch.empty(M, N, device=dev) for _ in range(R)]
    gxs = [torch.empty(M, K, device=dev) for _ in range(R)]
    w = torch.randn(K, N, device=dev) / K ** 0.5
    b = torch.zeros(N, device=dev)
    gw = torch.empty(K, N, device=dev)
    fwd = lambda i: check(lib.dr_dense_fwd(xs[i % R].data_ptr(), w.data_ptr(), b.data_ptr(), M, K, N, 1, ys[i % R].data_ptr(), st), "fwd")
    dx = lambda i: check(lib.dr_dense_bwd(xs[i % R].data_ptr(), w.data_ptr(), None, gys[i % R].data_ptr(), M, K, N, 0, None,
                                          gxs[i % R].data_ptr(), None, None, st), "dx")
    dw = lambda i: check(lib.dr_dense_bwd(xs[i % R].data_ptr(), w.data_ptr(), None, gys[i % R].data_ptr(), M, K, N, 0, None,
                                          None, gw.data_ptr(), None, st), "dw")
    t = {"fwd_us": timed(fwd), "dx_us": timed(dx), "dw_us": timed(dw)}
    fl = 2.0 * M * K * N
    print(json.dumps({"M": M, "K": K, "N": N, "knobs": knobs, **{k: round(v, 1) for k, v in t.items()},
                      **{k.replace("_us", "_tflops"): round(fl / v / 1e6, 1) for k, v in t.items()}}), flush=True)
    del xs, gys, ys, gxs
    torch.cuda.empty_cache()
