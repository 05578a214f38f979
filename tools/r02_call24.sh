#!/bin/bash
# Round-2 (second session) call 1: verify HEAD's GEMM splitter default (raw tile = hi plane) and A/B the new step options:
# dW GEMM enqueued first + co-residency build (tc_dw_share), sliced forward (fwd_chunks), pipelined e2e loop (fit_host).
mkdir -p gpurun_out
python -u -m pytest tests/test_gpu_models.py tests/test_gpu_dense_cross.py tests/test_gpu_fullsize_gemm.py tests/test_gpu_softmax.py tests/test_gpu_sharded.py \
    -m gpu -q --timeout=900 -rf --tb=short -n 4 -p no:cacheprovider > gpurun_out/r02o_tests.log 2>&1
tail -4 gpurun_out/r02o_tests.log | cut -c1-300
B="timeout 240 python -u bench.py --steps 20 --warmup 5 --no-cpu-baseline"
run() { tag=$1; shift; $B "$@" > gpurun_out/r02o_bench_$tag.log 2>&1
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r02o_bench_{tag}.log") if l.startswith("{")][-1])
    print(tag, round(d["value"] / 1e6, 2), "M ex/s", round(d["ms_per_step"], 4), "ms e2e", round(d["e2e"]["value"] / 1e6, 2),
          "blk", round(d["e2e"]["blocking_per_step"]["value"] / 1e6, 2), "frac", round(d["roofline"]["frac"], 4),
          "loss", round(d["final_loss"], 5), {k: round(v * 1e3, 1) for k, v in d["kernel_ms"].items()})
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/r02o_bench_{tag}.log").read()[-1500:])
PY
}
run default
run storehi --tune tc_store_hi=1
run dwfirst_share --dw-first 1 --tune tc_dw_share=1
run dwfirst --dw-first 1
run share --tune tc_dw_share=1
run chunks2 --fwd-chunks 2
run chunks4 --fwd-chunks 4
run chunks4_hints --fwd-chunks 4 --tune embed_l2_hints=3
run chunks2_hints --fwd-chunks 2 --tune embed_l2_hints=3
run all --dw-first 1 --tune tc_dw_share=1 --fwd-chunks 2
