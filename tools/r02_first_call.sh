#!/bin/bash
# Round-2 opener (one B200, ~3 min of box time): everything round 1 wrote after its GPU budget was spent gets
# its first run here, then the A/B lines that decide the next defaults.  Usage: gpurun -- 'bash tools/r02_first_call.sh'
mkdir -p gpurun_out
# 1. tests that have never run on a GPU (sharded two-tower at world 1, LINX backward) + the golden retrieval test
DR_UNVERIFIED=1 python -u -m pytest tests/test_gpu_zz_next_rows.py -m gpu -q -n 4 --timeout=150 -rf --tb=short \
    -p no:cacheprovider > gpurun_out/r02_unverified.log 2>&1
tail -3 gpurun_out/r02_unverified.log
# 2. the real step under each candidate default (each line records its knobs in "tune" / "gemm_core")
B="python -u bench.py --steps 20 --warmup 5 --no-cpu-baseline"
$B                               > gpurun_out/r02_bench_default.log 2>&1
$B --tune tc_min_n=32            > gpurun_out/r02_bench_tcmin32.log 2>&1      # N = 32 layers on the tcgen05 core
$B --tune embed_bwd_linx=1       > gpurun_out/r02_bench_bwdlinx.log 2>&1
DR_GEMM=tc $B                    > gpurun_out/r02_bench_tc1.log 2>&1          # pre-split planes, for the record
timeout 120 $B --ids zipf        > gpurun_out/r02_bench_zipf.log 2>&1         # SURVEY 8d second run (skewed ids)
for f in default tcmin32 bwdlinx tc1 zipf; do
  python - "$f" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r02_bench_{tag}.log") if l.startswith("{")][-1])
    print(tag, round(d["value"] / 1e6, 2), "M ex/s", round(d["ms_per_step"], 4), "ms",
          {k: round(v * 1e3, 1) for k, v in d["kernel_ms"].items()})
except Exception as e:
    print(tag, "FAILED", e)
PY
done
# 3. where the tcgen05 GEMM core waits (per-role mbarrier wait shares of the instrumented instantiation)
timeout 120 python -u tools/gemm_prof.py > gpurun_out/r02_gemm_prof.log 2>&1; cat gpurun_out/r02_gemm_prof.log | cut -c1-600
# 4. launch list of the default step (shares, not absolutes)
bash tools/ncu_launches.sh r02a > /dev/null 2>&1; python -c "
import json; [print(o) for o in sorted(json.load(open('gpurun_out/launches_r02a.json')), key=lambda o: -o['share'])[:12]]"
