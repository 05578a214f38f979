#!/usr/bin/env python
"""Do the layer-0 weight-gradient GEMM and the embedding update really run at the same time?  (C2 shapes, one GPU)

Times both kernels alone and together (either launch order) with CUDA events on their own streams, all relative to one
start event: if the two end times of a concurrent run are (a, a + b) the kernels ran one after the other; if both are
about max(a, b) they overlapped.  Knobs from the command line: KEY=VALUE ... (dr_tune_set)."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_recommenders_b200 import _lib, feature_column as fc, ops
from deep_recommenders_b200._lib import check
from deep_recommenders_b200.keras.models.ranking import DeepFM
from deep_recommenders_b200.training import DeepFMTrainStep

for kv in sys.argv[1:]:
    k, v = kv.split("=")
    _lib.tune(k, int(v))
dev = torch.device("cuda", 0)
S, D, B, rows = 26, 16, 65536, 1_000_000
cols = [fc.categorical_column_with_identity(f"C{i}", rows) for i in range(S)]
model = DeepFM([fc.indicator_column(c) for c in cols], [fc.embedding_column(c, D) for c in cols], dnn_units_size=[256, 32],
               seed=1, device=dev, sparse_lr=0.0)
tr = DeepFMTrainStep(model, batch_size=B, lr=0.0, use_graph=False)
gen = torch.Generator(device=dev).manual_seed(1)
pool = [torch.randint(0, rows, (B, S), device=dev, generator=gen) for _ in range(4)]
lab = torch.randint(0, 2, (B,), device=dev, generator=gen).float()
for i in range(3):
    tr.step(pool[i], lab)
torch.cuda.synchronize()
lib, c = tr.lib, tr.coll
main, side = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
l = tr.layers[0]
gz0, gz = tr.g_acts[0], tr.g_acts[-1]


def dw(st):
    check(lib.dr_dense_bwd(tr.stack.data_ptr(), tr.w[0].data_ptr(), None, gz0.data_ptr(), B, S * D, l.units, 0, None, None,
                           tr.gw[0].data_ptr(), None, st.cuda_stream), "dw")


def upd(st, nb=None):
    check(lib.dr_embed_fm_bwd(tr.ids.data_ptr(), 8, tr.rows.data_ptr(), tr.stack.data_ptr(), tr.sum_e.data_ptr(),
                              gz.data_ptr(), tr.g_stack.data_ptr(), nb or B, S, D, c.row_stride, c.lin_stride, c.flags,
                              tr.tp.data_ptr(), tr.lp.data_ptr(), c.bias.data_ptr(), 0.0, st.cuda_stream), "upd")


def flush():
    torch.empty(256 << 20, dtype=torch.uint8, device=dev).zero_()


def run(mode, it):
    tr.ids.copy_(pool[it % 4])
    flush()
    torch.cuda.synchronize()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    s0, gd, ud = ev(), ev(), ev()
    s0.record(main)
    side.wait_event(s0)
    if mode == "dw":
        dw(main); gd.record(main)
    elif mode == "upd":
        upd(side); ud.record(side)
    elif mode == "dw_first":
        dw(main); gd.record(main)
        upd(side); ud.record(side)
    elif mode == "small_upd":                 # 1/16 of the batch alone
        upd(side, B // 16); ud.record(side)
    elif mode == "dw_first_small_upd":        # ... and behind the GEMM: ends early only if its CTAs become resident beside it
        dw(main); gd.record(main)
        upd(side, B // 16); ud.record(side)
    else:
        upd(side); ud.record(side)
        dw(main); gd.record(main)
    torch.cuda.synchronize()
    return (s0.elapsed_time(gd) * 1e3 if mode not in ("upd", "small_upd") else None,
            s0.elapsed_time(ud) * 1e3 if mode != "dw" else None)


out = {}
for mode in ("dw", "upd", "dw_first", "upd_first", "small_upd", "dw_first_small_upd"):
    r = [run(mode, i) for i in range(8)][2:]
    mean = lambda xs: None if xs[0] is None else round(sum(xs) / len(xs), 1)
    out[mode] = {"dw_end_us": mean([a for a, _ in r]), "upd_end_us": mean([b for _, b in r])}
print(json.dumps({"knobs": sys.argv[1:], **out}))
