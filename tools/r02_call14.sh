#!/bin/bash
mkdir -p gpurun_out
python -u -m pytest tests/test_gpu_dense_cross.py tests/test_gpu_softmax.py tests/test_gpu_zz_next_rows.py -m gpu -q --timeout=900 -rf --tb=short -n 4 -p no:cacheprovider > gpurun_out/r02j_tests.log 2>&1
tail -3 gpurun_out/r02j_tests.log
timeout 300 python -u bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_n1_c5.log 2>&1
timeout 200 python -u tools/bench_two_tower.py > gpurun_out/r02b_n1_c4.log 2>&1
timeout 200 python -u bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02b_n1_default.log 2>&1
for f in default c5 c4; do
  python - "$f" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r02b_n1_{f}.log") if l.startswith("{")][-1])
    print(f, "N=1", round(d["value"] / 1e6, 2), "M ex/s", round(d["ms_per_step"], 4), "ms",
          {k: round(v * 1e3, 1) for k, v in d.get("kernel_ms", {}).items()})
except Exception as e:
    print(f, "FAILED", e); print(open(f"gpurun_out/r02b_n1_{f}.log").read()[-1200:])
PY
done
