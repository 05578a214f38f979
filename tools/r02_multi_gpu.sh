#!/bin/bash
# Round-2 multi-GPU call (gpurun --gpus N): world-N parity worker, then C2 (weak), C5 (strong) and C4 (strong) benches.
# Usage: gpurun --gpus 2 -- 'bash tools/r02_multi_gpu.sh 2 [tag]'
N=${1:-2}
TAG=${2:-r02b}
WHAT=${3:-all}          # all | c45 (only the C5 and C4 benches)
mkdir -p gpurun_out
T="timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$WHAT" = "c45" ]; then
  :
elif [ "$N" = "2" ]; then
  timeout 900 python -u -m pytest tests/test_gpu_sharded.py -m gpu -v --timeout=900 -rfP --tb=short -p no:cacheprovider \
      > gpurun_out/${TAG}_world${N}_tests.log 2>&1; tail -2 gpurun_out/${TAG}_world${N}_tests.log
  grep -a " ok: \|SHARDED_OK\|Error\|assert" gpurun_out/${TAG}_world${N}_tests.log | tail -14
else
  $T --master-port 29510 tests/sharded_worker.py > gpurun_out/${TAG}_world${N}_worker.log 2>&1
  grep -a " ok: \|SHARDED_OK\|Error\|assert" gpurun_out/${TAG}_world${N}_worker.log | tail -14
fi
B="bench.py --gpus $N --steps 20 --warmup 5"
[ "$WHAT" = "c45" ] || $T --master-port 29511 $B            > gpurun_out/${TAG}_n${N}_default.log 2>&1
$T --master-port 29513 $B --workload c5                     > gpurun_out/${TAG}_n${N}_c5.log 2>&1
$T --master-port 29514 tools/bench_two_tower.py             > gpurun_out/${TAG}_n${N}_c4.log 2>&1
for f in $([ "$WHAT" = "c45" ] || echo default) c5 c4; do
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
