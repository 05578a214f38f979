#!/bin/bash
# CTA-pair GEMM (tc_pair=1): parity suites, GEMM shape timings, the C2 step, C5 at N=1
mkdir -p gpurun_out
timeout 200 python -u tools/bench_gemm_shapes.py tc_pair=1 2>&1 | grep '^{' | tee -a gpurun_out/r02w_gemm_shapes_pair.jsonl
timeout 200 python -u tools/bench_gemm_shapes.py tc_pair=0 2>&1 | grep '^{' | tee -a gpurun_out/r02w_gemm_shapes_pair.jsonl
python -u -m pytest tests/test_gpu_dense_cross.py tests/test_gpu_fullsize_gemm.py tests/test_gpu_softmax.py tests/test_gpu_models.py \
    -m gpu -q --timeout=900 -rf --tb=short -n 4 -p no:cacheprovider > gpurun_out/r02w_tests.log 2>&1
tail -4 gpurun_out/r02w_tests.log | cut -c1-300
B="timeout 240 python -u bench.py --steps 20 --warmup 5 --no-cpu-baseline"
run() { tag=$1; shift; $B "$@" > gpurun_out/r02w_bench_$tag.log 2>&1
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r02w_bench_{tag}.log") if l.startswith("{")][-1])
    print(tag, round(d["value"] / 1e6, 2), "M ex/s", round(d["ms_per_step"], 4), "ms e2e", round(d["e2e"]["value"] / 1e6, 2),
          "frac", round(d["roofline"]["frac"], 4),
          "loss", round(d["final_loss"], 5), {k: round(v * 1e3, 1) for k, v in d["kernel_ms"].items()})
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/r02w_bench_{tag}.log").read()[-1500:])
PY
}
run default
run pair --tune tc_pair=1
run c5 --workload c5
run c5_pair --workload c5 --tune tc_pair=1
