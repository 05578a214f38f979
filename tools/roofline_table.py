"""Roofline table of the C2 train step (developer tool, no GPU needed): for every kernel group of `bench.py`'s
`kernel_ms`, the algorithmic bytes / flops, the time those would take at the MEASURED peaks of this pool's B200s
(MEASURED_PEAKS.json: HBM copy bandwidth, cuBLAS bf16 throughput; TF32 dense taken as half the bf16 rate, and a
3xTF32 product costs three TF32 MMAs), and the measured time.  Usage:
    python tools/roofline_table.py profiles/bench_n1_r01_linx.json > profiles/roofline_r01.md
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    bench = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm = peaks["hbm_gbs"] * 1e9
    tf32 = peaks["bf16_tflops"] * 1e12 / 2          # dense TF32 = half the bf16 rate
    B, S, D = 65536, 26, 16
    dims = [S * D, 256, 32, 1]
    km = bench["kernel_ms"]

    def gemm(M, K, N, reads, writes, tc):
        flops = 2.0 * M * K * N
        t_hbm = 4.0 * (reads + writes) / hbm
        t_tc = 3.0 * flops / tf32 if tc else None        # FFMA layers: no tensor-core floor quoted
        return flops, 4.0 * (reads + writes), t_hbm, t_tc

    rows = []
    alg_f = B * (S * (8 + 4 * D + 4) + 4 * S * D + 4)
    rows.append(("embed_fm_fwd (fused gather + first-order + FM)", "hbm", alg_f, None, alg_f / hbm, None, km["embed_fm_fwd"]))
    head = "head_bce" in km
    nfwd = 2 if head else 3
    for i in range(nfwd):
        M, K, N = B, dims[i], dims[i + 1]
        fl, by, th, tt = gemm(M, K, N, M * K + K * N, M * N, N >= 32)
        rows.append((f"dense_fwd_{i}  [{M} x {K}] @ [{K} x {N}]", "tensor" if tt and tt > th else "hbm", by, fl, th, tt,
                     km[f"dense_fwd_{i}"]))
    if head:
        by = 4.0 * B * (2 * dims[2] + 4)
        rows.append(("head_bce  Dense(1) + BCE + their backward (one kernel)", "hbm", by, None, by / hbm, None, km["head_bce"]))
    else:
        rows.append(("bce (loss + dL/dlogit)", "hbm", 4.0 * B * 4, None, 4.0 * B * 4 / hbm, None, km["bce"]))
    for i in ((1,) if head else (2, 1)):
        M, K, N = B, dims[i], dims[i + 1]
        fl = 2 * 2.0 * M * K * N                       # dX (chained: x act'(y) in the epilogue) and dW
        by = 4.0 * (2 * M * N + 3 * M * K + 2 * K * N)
        tt = 3.0 * fl / tf32 if N >= 32 else None
        rows.append((f"dense_bwd_{i}  dX + dW of layer {i}", "hbm", by, fl, by / hbm, tt, km[f"dense_bwd_{i}"]))
    M, K, N = B, dims[0], dims[1]
    fl = 2.0 * M * K * N
    by = 4.0 * (2 * M * N + M * K + K * N)
    rows.append(("dense_bwd_0_dx  column sums of gZ; dX = gZ @ W^T", "tensor", by, fl, by / hbm, 3.0 * fl / tf32,
                 km["dense_bwd_0_dx"]))
    alg_b = B * (S * 8 + 8 * S * D + 4 * D + 4 + 4 * S * D + 4 * S)
    by = 4.0 * (M * K + M * N + K * N) + alg_b
    rows.append(("dense_bwd_0_dw (tensor) then embed_fm_bwd + fused sparse SGD (hbm), two streams", "hbm", by, fl, by / hbm,
                 3.0 * fl / tf32, km["dense_bwd_0_dw+embed_fm_bwd"]))
    nparam = sum(dims[i] * dims[i + 1] + dims[i + 1] for i in range(3))
    rows.append(("sgd (tower)", "hbm", 12.0 * nparam, None, 12.0 * nparam / hbm, None, km["sgd"]))

    print("# Roofline of the C2 train step\n")
    print(f"Source: `{os.path.relpath(sys.argv[1], ROOT)}` (`kernel_ms`: CUDA events between kernel groups, eager pass, GEMM core "
          f"`{bench.get('gemm_core', 'tcgen05 3xTF32, pre-split planes (tc)')}`), peaks from `MEASURED_PEAKS.json`: HBM "
          f"{peaks['hbm_gbs']:.1f} GB/s, bf16 {peaks['bf16_tflops']:.1f} TFLOP/s (TF32 dense = half; a 3xTF32 product = 3 MMAs).\n")
    print("| kernel group | bound | algorithmic MB | useful GFLOP | floor at HBM peak (µs) | floor at tensor peak, 3xTF32 (µs) | "
          "measured (µs) | measured / max(floor) |")
    print("|---|---|---:|---:|---:|---:|---:|---:|")
    tot_floor = tot_meas = 0.0
    for name, bound, by, fl, th, tt, ms in rows:
        floor = max(th, tt or 0.0)
        tot_floor += floor
        tot_meas += ms * 1e-3
        print(f"| {name} | {bound} | {by / 1e6:.1f} | {'' if fl is None else f'{fl / 1e9:.1f}'} | {th * 1e6:.1f} | "
              f"{'' if tt is None else f'{tt * 1e6:.1f}'} | {ms * 1e3:.1f} | {ms * 1e-3 / floor:.1f}x |")
    print(f"| **step** | | | | | | **{tot_meas * 1e6:.0f}** (graph replay: {bench['ms_per_step'] * 1e3:.0f}) | "
          f"{tot_meas / tot_floor:.1f}x of {tot_floor * 1e6:.0f} µs |")
    print("\nReading it (end of round 2): the two big tensor-core GEMMs run on the CTA-pair kernel (tcgen05 cta_group::2) at "
          "1.4-1.7x of the 3xTF32 tensor floor (ncu: tensor pipe 76-82 % active, profiles/gemm_r02z_ncu.csv); before it the "
          "single-CTA main loop was bound by shared-memory bandwidth (2 250 cycles per 128 x 256 x 32 k-block against 1 536 of MMA "
          "time, profiles/r02u_gemm_prof.json).  The gather + FM forward sits at the DRAM's random-row rate, not at its byte "
          "rate (DESIGN.md section 4: same time at half the DRAM bytes, same time on half the SMs).  The layer-0 weight "
          "gradient and the embedding update share a label because they are issued on two streams; they run one after the "
          "other (about 70 + 130 us): making them co-resident is possible (profiles/r02r_probe_coresidency.jsonl: equal "
          "shared-memory carveout + a register-capped GEMM build) but the update then runs at 40 % of its solo rate and nothing is "
          "gained (profiles/r02q_probe_overlap.jsonl).  The skinny layers (256 -> 32 -> 1 and their backward) and the two "
          "launch-latency-bound tails are 2.8x and more above their HBM floors: 160 us of the step.")

if __name__ == "__main__":
    main()
