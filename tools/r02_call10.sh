#!/bin/bash
mkdir -p gpurun_out
python tools/bench_softmax.py 2>&1 | tail -2 | cut -c1-900
python -m pytest tests/test_gpu_softmax.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
B="python -u bench.py --steps 20 --warmup 5 --no-cpu-baseline"
timeout 200 $B --tune gemm_bn=256 > gpurun_out/r02h_bench_bn256.log 2>&1
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02h_bench_bn256.log") if l.startswith("{")][-1])
    print("bn256", round(d["value"] / 1e6, 2), "M ex/s", round(d["ms_per_step"], 4), "ms loss", round(d["final_loss"], 5),
          {k: round(v * 1e3, 1) for k, v in d["kernel_ms"].items()})
except Exception as e:
    print("bn256 FAILED", e); print(open("gpurun_out/r02h_bench_bn256.log").read()[-1500:])
PY
