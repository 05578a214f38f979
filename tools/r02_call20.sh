#!/bin/bash
mkdir -p gpurun_out
python -u -m pytest tests/test_gpu_zz_next_rows.py -m gpu -q --timeout=900 -rf --tb=short -n 4 -p no:cacheprovider -k "adam" > gpurun_out/r02l_tests.log 2>&1
tail -12 gpurun_out/r02l_tests.log | cut -c1-300
B="python -u bench.py --steps 20 --warmup 5 --no-cpu-baseline"
for o in adam_rows_tf adam_rows; do
  timeout 300 $B --optimizer $o > gpurun_out/r02l_bench_$o.log 2>&1
  python - "$o" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r02l_bench_{tag}.log") if l.startswith("{")][-1])
    print(tag, round(d["value"] / 1e6, 2), "M ex/s", round(d["ms_per_step"], 4), "ms loss", round(d["final_loss"], 5),
          {k: round(v * 1e3, 1) for k, v in d["kernel_ms"].items()})
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/r02l_bench_{tag}.log").read()[-1500:])
PY
done
