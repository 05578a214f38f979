"""Where does the tcgen05 GEMM core (variant 2) wait?  Runs the C2 layer shapes with the instrumented instantiation
(knob gemm_prof = 1) and prints, per shape, the share of the kernel span each warp role spent waiting on its
mbarrier (dr_gemm_prof_read).  One short GPU run; writes gpurun_out/gemm_prof.json."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_recommenders_b200 import _lib  # noqa: E402

NAMES = ["producer_wait_free_stage", "splitter_wait_tma", "splitter_work", "mma_wait_operands", "mma_wait_accumulator",
         "epilogue_wait_accumulator", "epilogue_work", "kernel_span", "ctas", "kblocks"]


def main():
    lib = _lib.load()
    _lib.enable_tensor_core_gemm(variant=2)
    _lib.tune("gemm_bn", 128)      # the instrumented instantiation exists for the 128-wide tile only
    st = torch.cuda.current_stream().cuda_stream
    out = []
    buf = (C.c_uint64 * 16)()
    for (M, K, N, ta, tb, label) in [(65536, 416, 256, 0, 0, "layer-0 forward  X[M,K] @ W[K,N]"),
                                     (65536, 256, 416, 0, 1, "layer-0 dX       gZ[M,256] @ W^T"),
                                     (65536, 256, 256, 0, 0, "256 -> 256 forward")]:
        A = torch.randn((K, M) if ta else (M, K), device="cuda")
        B = torch.randn((N, K) if tb else (K, N), device="cuda")
        Cm = torch.zeros((M, N), device="cuda")
        for prof in (0, 1):
            _lib.tune("gemm_prof", prof)
            for _ in range(3):
                _lib.check(lib.dr_debug_gemm(A.data_ptr(), B.data_ptr(), Cm.data_ptr(), M, N, K, ta, tb, st), "gemm")
            torch.cuda.synchronize()
            if prof:
                _lib.check(lib.dr_gemm_prof_read(buf, 1), "prof_read")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                _lib.check(lib.dr_debug_gemm(A.data_ptr(), B.data_ptr(), Cm.data_ptr(), M, N, K, ta, tb, st), "gemm")
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            if prof:
                _lib.check(lib.dr_gemm_prof_read(buf, 1), "prof_read")
                v = [int(x) for x in buf]
                span = max(v[7], 1)
                r = dict(shape=label, M=M, K=K, N=N, ms_instrumented=ms,
                         **{n: (v[i] if i >= 8 else round(v[i] / span, 3)) for i, n in enumerate(NAMES)})
                r["cycles_per_kblock_per_cta"] = round(v[7] / max(v[9], 1), 1)
            else:
                r0 = ms
        r["ms_plain"] = r0
        print(json.dumps(r), flush=True)
        out.append(r)
    # the weight-gradient GEMM of layer 0 (both operands MN-major as stored, split-K, TMA reduce-add) through dr_dense_bwd
    for (M, K, N, label) in [(65536, 416, 256, "layer-0 dW  X^T[416,M] @ gZ[M,256] (C2)"),
                             (65536, 3328, 512, "layer-0 dW  X^T[3328,M] @ gZ[M,512] (C5)")]:
        x = torch.randn((M, K), device="cuda")
        g = torch.randn((M, N), device="cuda")
        w = torch.randn((K, N), device="cuda")
        gw = torch.zeros((K, N), device="cuda")
        call = lambda: _lib.check(lib.dr_dense_bwd(x.data_ptr(), w.data_ptr(), None, g.data_ptr(), M, K, N, 0, None, None,
                                                   gw.data_ptr(), None, st), "dw")
        for prof in (0, 1):
            _lib.tune("gemm_prof", prof)
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            if prof:
                _lib.check(lib.dr_gemm_prof_read(buf, 1), "prof_read")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                call()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            if prof:
                _lib.check(lib.dr_gemm_prof_read(buf, 1), "prof_read")
                v = [int(x) for x in buf]
                span = max(v[7], 1)
                r = dict(shape=label, M=M, K=K, N=N, ms_instrumented=ms,
                         **{n: (v[i] if i >= 8 else round(v[i] / span, 3)) for i, n in enumerate(NAMES)})
                r["cycles_per_kblock_per_cta"] = round(v[7] / max(v[9], 1), 1)
            else:
                r0 = ms
        r["ms_plain"] = r0
        print(json.dumps(r), flush=True)
        out.append(r)
        del x, g, w, gw
    _lib.tune("gemm_prof", 0)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/gemm_prof.json", "w"), indent=1)


if __name__ == "__main__":
    main()
