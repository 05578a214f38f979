"""Runs the Dense forward/backward once at the C2 layer-1 shape (for an ncu --set full capture)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_recommenders_b200 import _lib, ops  # noqa: E402

_lib.enable_tensor_core_gemm(1 << 30)
M, K, N = 65536, 416, 256
x = torch.randn(M, K, device="cuda", requires_grad=True)
w = (torch.randn(K, N, device="cuda") / K ** 0.5).requires_grad_(True)
b = torch.randn(N, device="cuda", requires_grad=True)
for _ in range(2):
    y = ops.DenseFn.apply(x, w, b, 1)
    y.backward(torch.randn_like(y))
torch.cuda.synchronize()
