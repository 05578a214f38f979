#!/bin/bash
# pair kernel as the default: full GPU suite, the C2 bench line (with CPU arm), C5, shape timings
mkdir -p gpurun_out
python -u -m pytest tests -m gpu -q --timeout=900 -rf --tb=short -n 4 -p no:cacheprovider > gpurun_out/r02x_tests.log 2>&1
tail -4 gpurun_out/r02x_tests.log | cut -c1-300
timeout 200 python -u tools/bench_gemm_shapes.py 2>&1 | grep '^{' | tee -a gpurun_out/r02x_gemm_shapes.jsonl
B="timeout 240 python -u bench.py --steps 20 --warmup 5 --no-cpu-baseline"
run() { tag=$1; shift; $B "$@" > gpurun_out/r02x_bench_$tag.log 2>&1
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r02x_bench_{tag}.log") if l.startswith("{")][-1])
    print(tag, round(d["value"] / 1e6, 2), "M ex/s", round(d["ms_per_step"], 4), "ms e2e", round(d["e2e"]["value"] / 1e6, 2),
          "blk", round(d["e2e"]["blocking_per_step"]["value"] / 1e6, 2), "frac", round(d["roofline"]["frac"], 4),
          "loss", round(d["final_loss"], 5), {k: round(v * 1e3, 1) for k, v in d["kernel_ms"].items()})
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/r02x_bench_{tag}.log").read()[-1500:])
PY
}
run default
run single --tune tc_pair=0
run c5 --workload c5
run zipf --ids zipf
run adam_rows --optimizer adam_rows
run adam_rows_tf --optimizer adam_rows_tf
timeout 200 python -u tools/bench_configs.py > gpurun_out/r02x_bench_configs.log 2>&1; grep '^{' gpurun_out/r02x_bench_configs.log | cut -c1-300
