#!/bin/bash
mkdir -p gpurun_out
python -u -m pytest tests -m gpu -q --timeout=900 -rf --tb=short -n 4 -p no:cacheprovider > gpurun_out/r02e_tests_all.log 2>&1
tail -6 gpurun_out/r02e_tests_all.log
B="python -u bench.py --steps 20 --warmup 5 --no-cpu-baseline"
for h in 0 1 2 3 4 5 7 12 15; do
  timeout 200 $B --tune embed_l2_hints=$h > gpurun_out/r02e_bench_h$h.log 2>&1
  python - "$h" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r02e_bench_h{tag}.log") if l.startswith("{")][-1])
    print("hints", tag, round(d["value"] / 1e6, 2), "M ex/s", round(d["ms_per_step"], 4), "ms b2b_fwd_us", round(d["roofline"]["us_per_launch"], 1),
          {k: round(v * 1e3, 1) for k, v in d["kernel_ms"].items()})
except Exception as e:
    print(tag, "FAILED", e)
PY
done
