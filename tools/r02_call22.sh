#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -u -m pytest tests/test_gpu_embed.py -m gpu -q --timeout=200 -rf --tb=short -n 4 -p no:cacheprovider -k "tma" > gpurun_out/r02m_tests.log 2>&1
tail -6 gpurun_out/r02m_tests.log | cut -c1-300
B="timeout 200 python -u bench.py --steps 20 --warmup 5 --no-cpu-baseline"
for t in "tma --embed-fwd tma" "tma_zipf --embed-fwd tma --ids zipf"; do
  set -- $t; tag=$1; shift
  $B "$@" > gpurun_out/r02m_bench_$tag.log 2>&1
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r02m_bench_{tag}.log") if l.startswith("{")][-1])
    print(tag, round(d["value"] / 1e6, 2), "M ex/s", round(d["ms_per_step"], 4), "ms loss", round(d["final_loss"], 5), "b2b", round(d["roofline"]["back_to_back"]["us_per_launch"], 1),
          {k: round(v * 1e3, 1) for k, v in d["kernel_ms"].items()})
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/r02m_bench_{tag}.log").read()[-1500:])
PY
done
