"""Where does the C4 in-batch softmax (B = 16384, D = 64) spend its time?  Times the pieces of the tensor-core form
(score GEMM, row passes, the two gradient GEMMs) and both complete forms.  One GPU, ~10 s."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_recommenders_b200 import _lib  # noqa: E402


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    lib = _lib.load()
    B, D = int(os.environ.get("B", 16384)), 64
    st = torch.cuda.current_stream().cuda_stream
    q = torch.randn(B, D, device="cuda") * 0.3
    c = torch.randn(B, D, device="cuda") * 0.3
    ws = torch.empty((B, B), device="cuda")
    lse, loss, gl = torch.empty(B, device="cuda"), torch.zeros(1, device="cuda"), torch.ones(1, device="cuda")
    gq, gc = torch.empty_like(q), torch.empty_like(c)
    ck = _lib.check
    r = {"B": B, "D": D}
    r["scores_gemm_ms"] = timed(lambda: ck(lib.dr_scores_fwd(q.data_ptr(), c.data_ptr(), None, None, B, B, D, ws.data_ptr(), st), "s"))
    r["fwd_ws_ms"] = timed(lambda: ck(lib.dr_inbatch_softmax_fwd_ws(q.data_ptr(), c.data_ptr(), None, None, None, 1.0, B, B, D,
                                                                    ws.data_ptr(), B, lse.data_ptr(), loss.data_ptr(), st), "f"))
    r["bwd_ws_reuse_ms"] = timed(lambda: ck(lib.dr_inbatch_softmax_bwd_ws(q.data_ptr(), c.data_ptr(), None, None, None, 1.0, B, B, D,
                                                                          lse.data_ptr(), gl.data_ptr(), ws.data_ptr(), B, 1,
                                                                          gq.data_ptr(), gc.data_ptr(), st), "b"))
    r["bwd_ws_recompute_ms"] = timed(lambda: ck(lib.dr_inbatch_softmax_bwd_ws(q.data_ptr(), c.data_ptr(), None, None, None, 1.0, B, B, D,
                                                                              lse.data_ptr(), gl.data_ptr(), ws.data_ptr(), B, 0,
                                                                              gq.data_ptr(), gc.data_ptr(), st), "b"))
    # the two gradient contractions alone (plain GEMMs of the same shapes)
    r["gq_gemm_ms"] = timed(lambda: ck(lib.dr_debug_gemm(ws.data_ptr(), c.data_ptr(), gq.data_ptr(), B, D, B, 0, 0, st), "g"))
    r["gc_gemm_ms"] = timed(lambda: ck(lib.dr_debug_gemm(ws.data_ptr(), q.data_ptr(), gc.data_ptr(), B, D, B, 1, 0, st), "g"))
    r["ffma_fwd_ms"] = timed(lambda: ck(lib.dr_inbatch_softmax_fwd(q.data_ptr(), c.data_ptr(), None, None, None, 1.0, B, B, D,
                                                                   lse.data_ptr(), loss.data_ptr(), st), "f"))
    r["ffma_bwd_ms"] = timed(lambda: ck(lib.dr_inbatch_softmax_bwd(q.data_ptr(), c.data_ptr(), None, None, None, 1.0, B, B, D,
                                                                   lse.data_ptr(), gl.data_ptr(), gq.data_ptr(), gc.data_ptr(), st), "b"))
    flops = 2.0 * B * B * D
    r["tc_total_ms"] = r["fwd_ws_ms"] + r["bwd_ws_reuse_ms"]
    r["ffma_total_ms"] = r["ffma_fwd_ms"] + r["ffma_bwd_ms"]
    r["tc_useful_tflops"] = 3 * flops / r["tc_total_ms"] / 1e9
    print(json.dumps(r), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(r, open("gpurun_out/bench_softmax.json", "w"), indent=1)


if __name__ == "__main__":
    main()
