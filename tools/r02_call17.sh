#!/bin/bash
# N=2 experiments: two-stage dW GEMM (co-residency with the p2p update), deeper unroll of the wide-row p2p gather
N=2
mkdir -p gpurun_out
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
B="bench.py --gpus $N --steps 20 --warmup 5"
run() { tag=$1; shift; $T --master-port $((29600 + RANDOM % 200)) $B "$@" > gpurun_out/r02k_n2_$tag.log 2>&1
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r02k_n2_{tag}.log") if l.startswith("{")][-1])
    print(tag, round(d["value"] / 1e6, 2), "M ex/s", round(d["ms_per_step"], 4), "ms",
          {k: round(v * 1e3, 1) for k, v in d.get("kernel_ms", {}).items()})
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/r02k_n2_{tag}.log").read()[-1200:])
PY
}
run c2_base
run c2_dw2 --tune tc_dw_stages=2
run c5_base --workload c5
run c5_dw2 --workload c5 --tune tc_dw_stages=2
run c5_u8 --workload c5 --tune embed_fwd_unroll=8
$T --master-port 29555 tools/bench_two_tower.py > gpurun_out/r02c_n2_c4.log 2>&1; grep '^{' gpurun_out/r02c_n2_c4.log | cut -c1-400
