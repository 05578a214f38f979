#!/bin/bash
mkdir -p gpurun_out
timeout 300 tools/bin/mb_gather --stride 32 --only pf > gpurun_out/mb_pf_s32.jsonl 2> gpurun_out/mb_pf_s32.err
cut -c1-200 gpurun_out/mb_pf_s32.jsonl; grep CTAs gpurun_out/mb_pf_s32.err
echo "--- half the SMs (per-SM cap vs DRAM): reg_nc_na at 74 SMs"
timeout 100 tools/bin/mb_gather --stride 32 --only reg_nc_na --sms 74 2>/dev/null | cut -c1-200
timeout 100 tools/bin/mb_gather --stride 32 --only pf1_w8 --sms 74 2>/dev/null | cut -c1-200
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_requests_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct --clock-control none --csv \
   --log-file gpurun_out/mb_pf_s32_ncu.csv tools/bin/mb_gather --stride 32 --iters 0 --only pf > /dev/null 2>&1
python -u -m pytest tests/test_gpu_zz_next_rows.py tests/test_gpu_sharded.py -m gpu -q -x --timeout=900 -rf --tb=short -n 4 -p no:cacheprovider > gpurun_out/r02d_tests.log 2>&1
tail -4 gpurun_out/r02d_tests.log
B="python -u bench.py --steps 15 --warmup 4 --no-cpu-baseline"
timeout 300 $B --optimizer lazy_adam > gpurun_out/r02d_bench_lazy_adam.log 2>&1
timeout 300 $B --optimizer adam      > gpurun_out/r02d_bench_adam.log 2>&1
for f in lazy_adam adam; do
  python - "$f" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r02d_bench_{tag}.log") if l.startswith("{")][-1])
    print(tag, round(d["value"] / 1e6, 2), "M ex/s", round(d["ms_per_step"], 4), "ms",
          {k: round(v * 1e3, 1) for k, v in d["kernel_ms"].items()})
except Exception as e:
    print(tag, "FAILED", e)
PY
done
