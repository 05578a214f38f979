#!/bin/bash
mkdir -p gpurun_out
python -u -m pytest tests/test_gpu_dense_cross.py tests/test_gpu_fullsize_gemm.py tests/test_gpu_softmax.py -m gpu -q --timeout=900 -rf --tb=short -n 4 -p no:cacheprovider > gpurun_out/r02n_tests.log 2>&1
tail -4 gpurun_out/r02n_tests.log | cut -c1-300
B="timeout 200 python -u bench.py --steps 20 --warmup 5 --no-cpu-baseline"
for t in "rawhi" "storehi --tune tc_store_hi=1"; do
  set -- $t; tag=$1; shift
  $B "$@" > gpurun_out/r02n_bench_$tag.log 2>&1
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r02n_bench_{tag}.log") if l.startswith("{")][-1])
    print(tag, round(d["value"] / 1e6, 2), "M ex/s", round(d["ms_per_step"], 4), "ms loss", round(d["final_loss"], 5),
          {k: round(v * 1e3, 1) for k, v in d["kernel_ms"].items()})
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/r02n_bench_{tag}.log").read()[-1500:])
PY
done
timeout 300 python -u tools/bench_configs.py > gpurun_out/r02n_bench_configs.log 2>&1; grep '^{' gpurun_out/r02n_bench_configs.log | cut -c1-260
