#!/bin/bash
# call 2 of the session: does the embedding update become resident beside the weight-gradient GEMM once both kernels ask for
# the same shared-memory carveout?  (call 24: dW-first + the co-residency build alone changed nothing: 225 us either way)
mkdir -p gpurun_out
python -u -m pytest tests/test_gpu_models.py -m gpu -q --timeout=900 -rf --tb=short -n 4 -p no:cacheprovider > gpurun_out/r02p_tests.log 2>&1
tail -3 gpurun_out/r02p_tests.log | cut -c1-300
B="timeout 240 python -u bench.py --steps 20 --warmup 5 --no-cpu-baseline"
run() { tag=$1; shift; $B "$@" > gpurun_out/r02p_bench_$tag.log 2>&1
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r02p_bench_{tag}.log") if l.startswith("{")][-1])
    print(tag, round(d["value"] / 1e6, 2), "M ex/s", round(d["ms_per_step"], 4), "ms e2e", round(d["e2e"]["value"] / 1e6, 2),
          "blk", round(d["e2e"]["blocking_per_step"]["value"] / 1e6, 2), "frac", round(d["roofline"]["frac"], 4),
          "loss", round(d["final_loss"], 5), {k: round(v * 1e3, 1) for k, v in d["kernel_ms"].items()})
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/r02p_bench_{tag}.log").read()[-1500:])
PY
}
run default
run cta2 --tune embed_ctas_per_sm=2
run carve_dwfirst_share --dw-first 1 --tune tc_dw_share=1 --tune embed_bwd_carveout=100
run carve_dwfirst --dw-first 1 --tune embed_bwd_carveout=100
run carve_share --tune tc_dw_share=1 --tune embed_bwd_carveout=100
run carve_dwfirst_share_cta2 --dw-first 1 --tune tc_dw_share=1 --tune embed_bwd_carveout=100 --tune embed_ctas_per_sm=2
run carve_dwfirst_stages2 --dw-first 1 --tune tc_dw_stages=2 --tune embed_bwd_carveout=100
