#!/bin/bash
# Round-2 call 4 (one B200): hybrid gather variants; GEMM with the TMA-store epilogue: parity + step time.
mkdir -p gpurun_out
timeout 300 tools/bin/mb_gather --stride 32 --only hyb > gpurun_out/mb_hyb_s32.jsonl 2> gpurun_out/mb_hyb_s32.err
timeout 120 tools/bin/mb_gather --stride 32 --only reg_nc_na > gpurun_out/mb_hyb_ref.jsonl 2>/dev/null
cat gpurun_out/mb_hyb_ref.jsonl gpurun_out/mb_hyb_s32.jsonl | cut -c1-200
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_requests_srcunit_tex_op_read.sum --clock-control none --csv \
   --log-file gpurun_out/mb_hyb_s32_ncu.csv tools/bin/mb_gather --stride 32 --iters 0 --only hyb > /dev/null 2>&1
python -u -m pytest tests/test_gpu_dense_cross.py tests/test_gpu_fullsize_gemm.py tests/test_gpu_assembled_models.py tests/test_gpu_softmax.py -m gpu -q -x --timeout=900 -rf --tb=short \
    -n 4 -p no:cacheprovider > gpurun_out/r02_gemmtests.log 2>&1
tail -8 gpurun_out/r02_gemmtests.log
B="python -u bench.py --steps 20 --warmup 5 --no-cpu-baseline"
$B                               > gpurun_out/r02b_bench_default.log 2>&1
$B --tune tc_tma_out=0           > gpurun_out/r02b_bench_regstore.log 2>&1
$B --tune tc_min_n=32            > gpurun_out/r02b_bench_tcmin32.log 2>&1
for f in default regstore tcmin32; do
  python - "$f" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r02b_bench_{tag}.log") if l.startswith("{")][-1])
    print(tag, round(d["value"] / 1e6, 2), "M ex/s", round(d["ms_per_step"], 4), "ms",
          {k: round(v * 1e3, 1) for k, v in d["kernel_ms"].items()})
except Exception as e:
    print(tag, "FAILED", e)
PY
done
timeout 120 python -u tools/gemm_prof.py > gpurun_out/r02b_gemm_prof.log 2>&1; cut -c1-420 gpurun_out/r02b_gemm_prof.log
