#!/bin/bash
# N-GPU benches on the final tree (no parity worker): C5 strong, C2 weak, C4 strong
N=${1:-8}
TAG=r02z
mkdir -p gpurun_out
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
B="bench.py --gpus $N --steps 20 --warmup 5"
$T --master-port 29513 $B --workload c5        > gpurun_out/${TAG}_n${N}_c5.log 2>&1
$T --master-port 29511 $B                      > gpurun_out/${TAG}_n${N}_default.log 2>&1
$T --master-port 29514 tools/bench_two_tower.py > gpurun_out/${TAG}_n${N}_c4.log 2>&1
for f in c5 default c4; do
  python - "$f" "$N" "$TAG" <<'PY'
import json, sys
f, n, tag = sys.argv[1:4]
try:
    d = json.loads([l for l in open(f"gpurun_out/{tag}_n{n}_{f}.log") if l.startswith("{")][-1])
    print(f, "N=" + n, round(d["value"] / 1e6, 2), "M ex/s", round(d["ms_per_step"], 4), "ms",
          {k: round(v * 1e3, 1) for k, v in d.get("kernel_ms", {}).items()})
except Exception as e:
    print(f, "FAILED", e); print(open(f"gpurun_out/{tag}_n{n}_{f}.log").read()[-1200:])
PY
done
