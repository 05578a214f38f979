#!/bin/bash
mkdir -p gpurun_out
python tools/bench_softmax.py 2>&1 | tail -1 | cut -c1-900
python -u -m pytest tests/test_gpu_dense_cross.py tests/test_gpu_fullsize_gemm.py tests/test_gpu_softmax.py tests/test_gpu_zz_next_rows.py -m gpu -q --timeout=900 -rf --tb=short -n 4 -p no:cacheprovider > gpurun_out/r02i_tests.log 2>&1
tail -4 gpurun_out/r02i_tests.log
B="python -u bench.py --steps 20 --warmup 5 --no-cpu-baseline"
run() { tag=$1; shift; timeout 300 $B "$@" > gpurun_out/r02i_bench_$tag.log 2>&1
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r02i_bench_{tag}.log") if l.startswith("{")][-1])
    print(tag, round(d["value"] / 1e6, 2), "M ex/s", round(d["ms_per_step"], 4), "ms e2e", round(d["e2e"]["value"] / 1e6, 2), "loss", round(d["final_loss"], 5),
          {k: round(v * 1e3, 1) for k, v in d["kernel_ms"].items()})
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/r02i_bench_{tag}.log").read()[-1500:])
PY
}
run default
run promo128 --tune tc_l2_promo=128
timeout 300 python -u tools/bench_configs.py > gpurun_out/r02i_bench_configs.log 2>&1; grep '^{' gpurun_out/r02i_bench_configs.log | cut -c1-300
