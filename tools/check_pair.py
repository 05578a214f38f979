#!/usr/bin/env python
"""CTA-pair tcgen05 kernel (knob tc_pair=1) against the single-CTA kernel and float64, through dr_dense_fwd / dr_dense_bwd
(forward: A K-major, B MN-major; dX: both K-major; dW: both MN-major + split-K reduce-adds), ragged and full-size shapes.
Exit code 1 on mismatch.  Run under `timeout`: a protocol bug in the pair kernel shows as a hang."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_recommenders_b200 import _lib
from deep_recommenders_b200._lib import check

lib = _lib.load()
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
shapes = [(256, 64, 128), (128, 32, 256), (300, 96, 200), (1000, 416, 256), (4096, 256, 416), (777, 832, 832),
          (65536, 416, 256), (8192, 3328, 512)]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
ok = True
for M, K, N in shapes:
    g = torch.Generator(device=dev).manual_seed(M + K + N)
    x = torch.randn(M, K, device=dev, generator=g)
    w = torch.randn(K, N, device=dev, generator=g) / K ** 0.5
    b = torch.randn(N, device=dev, generator=g) * 0.1
    gy = torch.randn(M, N, device=dev, generator=g)
    res = {}
    for pair in (0, 1):
        _lib.tune("tc_pair", pair)
        y, gx, gw = torch.empty(M, N, device=dev), torch.empty(M, K, device=dev), torch.empty(K, N, device=dev)
        check(lib.dr_dense_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), M, K, N, 0, y.data_ptr(), st), "fwd")
        check(lib.dr_dense_bwd(x.data_ptr(), w.data_ptr(), None, gy.data_ptr(), M, K, N, 0, None, gx.data_ptr(), gw.data_ptr(),
                               None, st), "bwd")
        torch.cuda.synchronize()
        res[pair] = (y, gx, gw)
        print(f"  M={M} K={K} N={N} tc_pair={pair} ran", flush=True)
    _lib.tune("tc_pair", 0)
    xd, wd, gd = x.double(), w.double(), gy.double()
    ref = (xd @ wd + b.double(), gd @ wd.T, xd.T @ gd)
    sc = (xd.abs() @ wd.abs() + b.abs().double(), gd.abs() @ wd.abs().T, xd.abs().T @ gd.abs())
    line = dict(M=M, K=K, N=N)
    for i, name in enumerate(("y", "gx", "gw")):
        e = ((res[1][i].double() - ref[i]).abs() / (sc[i] + 1e-30)).max().item()
        e0 = ((res[0][i].double() - ref[i]).abs() / (sc[i] + 1e-30)).max().item()
        line[name + "_pair_relerr"], line[name + "_single_relerr"] = e, e0
        line[name + "_pair_eq_single"] = bool(torch.equal(res[0][i], res[1][i]))
        if not (e <= 1e-5):
            ok = False
    print(json.dumps(line), flush=True)
print("PAIR_OK" if ok else "PAIR_MISMATCH", flush=True)
sys.exit(0 if ok else 1)
