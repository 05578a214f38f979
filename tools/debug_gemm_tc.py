"""Structured-input probes of the tcgen05 GEMM: which (m,k)/(k,n) element lands where."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_recommenders_b200 import _lib  # noqa: E402

torch.set_printoptions(linewidth=220, precision=1, sci_mode=False)


def gemm(A, B, M, N, K, ta, tb, variant):
    _lib.tune("gemm_variant", variant)
    C = torch.full((M, N), -7.0, device="cuda")
    _lib.check(_lib.load().dr_debug_gemm(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, ta, tb,
                                         torch.cuda.current_stream().cuda_stream), "gemm")
    torch.cuda.synchronize()
    return C


def main():
    _lib.enable_tensor_core_gemm(1 << 30)
    _lib.tune("tc_mn", int(os.environ.get("DR_TC_MN", "0")))
    M, N, K = [int(x) for x in os.environ.get("DR_MNK", "128,128,32").split(",")]
    for ta in (0, 1):
        for tb in (0, 1):
            print(f"==== ta={ta} tb={tb}")
            # logical A[m,k] = 1 if k == k0 ; logical B[k,n] = n + 1000*k  => C[m,n] = n + 1000*k0
            k0 = 5
            Al = torch.zeros(M, K, device="cuda")
            Al[:, k0] = 1.0
            Bl = (torch.arange(N, device="cuda")[None, :] + 1000.0 * torch.arange(K, device="cuda")[:, None]).float()
            A = Al.t().contiguous() if ta else Al
            B = Bl.t().contiguous() if tb else Bl
            C = gemm(A, B, M, N, K, ta, tb, 1)
            ref = Al @ Bl
            print("probe1 (pick B row k0=5): max abs err", float((C - ref).abs().max()))
            print(" C[0,:40]   ", C[0, :40].tolist())
            print(" C[0,64:80] ", C[0, 64:80].tolist())
            print(" C[:12,0]   ", C[:12, 0].tolist())
            print(" C[64:76,3] ", C[64:76, 3].tolist())
            # logical A[m,k] = m + 1000*k ; B[k,n] = 1 if (k==k0 and n==n0) => C[m,n0] = m + 1000*k0
            n0 = 9
            Al = (torch.arange(M, device="cuda")[:, None] + 1000.0 * torch.arange(K, device="cuda")[None, :]).float()
            Bl = torch.zeros(K, N, device="cuda")
            Bl[k0, n0] = 1.0
            A = Al.t().contiguous() if ta else Al
            B = Bl.t().contiguous() if tb else Bl
            C = gemm(A, B, M, N, K, ta, tb, 1)
            ref = Al @ Bl
            print("probe2 (pick A col k0=5 into n0=9): max abs err", float((C - ref).abs().max()))
            nz = (C.abs() > 0.5).nonzero()
            print(" nonzero count", nz.shape[0], "first", nz[:12].tolist())
            print(" C[:16,9]   ", C[:16, 9].tolist())
            print(" C[100:116,9]", C[100:116, 9].tolist())
            # random small check
            g = torch.Generator(device="cuda").manual_seed(1)
            Al = torch.randn(M, K, device="cuda", generator=g)
            Bl = torch.randn(K, N, device="cuda", generator=g)
            A = Al.t().contiguous() if ta else Al
            B = Bl.t().contiguous() if tb else Bl
            C = gemm(A, B, M, N, K, ta, tb, 1)
            ref = (Al.double() @ Bl.double())
            print("random: max abs err", float((C.double() - ref).abs().max()), "max ref", float(ref.abs().max()))
            # per-k-step contribution: A nonzero only in k-step j (8 columns)
            for j in range(4):
                A2 = torch.zeros_like(Al)
                A2[:, 8 * j:8 * j + 8] = Al[:, 8 * j:8 * j + 8]
                Aa = A2.t().contiguous() if ta else A2
                C = gemm(Aa, B, M, N, K, ta, tb, 1)
                ref = A2.double() @ Bl.double()
                print(f"  kstep {j}: max abs err", float((C.double() - ref).abs().max()))


if __name__ == "__main__":
    main()
