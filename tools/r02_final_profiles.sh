#!/bin/bash
# Round-2 evidence run (one B200): full GPU test suite, the bench line, ncu captures of the shipped kernels, launch list.
mkdir -p gpurun_out
python -u -m pytest tests -m gpu -q --timeout=900 -rf --tb=short -n 4 -p no:cacheprovider > gpurun_out/r02_final_tests.log 2>&1
tail -3 gpurun_out/r02_final_tests.log
timeout 400 python -u bench.py --steps 30 --warmup 5 > gpurun_out/r02_final_bench.log 2>&1; grep '^{' gpurun_out/r02_final_bench.log | cut -c1-500
timeout 200 python -u bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_final_bench_ref.log 2>&1; grep '^{' gpurun_out/r02_final_bench_ref.log | cut -c1-400
B="python -u bench.py --steps 20 --warmup 5 --no-cpu-baseline"
run() { tag=$1; shift; timeout 300 $B "$@" > gpurun_out/r02_final_bench_$tag.log 2>&1
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r02_final_bench_{tag}.log") if l.startswith("{")][-1])
    print(tag, round(d["value"] / 1e6, 2), "M ex/s", round(d["ms_per_step"], 4), "ms e2e", round(d["e2e"]["value"] / 1e6, 2),
          {k: round(v * 1e3, 1) for k, v in d["kernel_ms"].items()})
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/r02_final_bench_{tag}.log").read()[-1200:])
PY
}
run zipf --ids zipf
run zipf_agg --ids zipf --tune embed_bwd_mode=1
run adam_rows --optimizer adam_rows
timeout 120 python -u tools/bench_two_tower.py > gpurun_out/r02c_n1_c4.log 2>&1; grep '^{' gpurun_out/r02c_n1_c4.log | cut -c1-300
timeout 120 python -u tools/gemm_prof.py > gpurun_out/r02_final_gemm_prof.log 2>&1; cut -c1-330 gpurun_out/r02_final_gemm_prof.log; cp gpurun_out/gemm_prof.json gpurun_out/r02_final_gemm_prof.json
bash tools/ncu_launches.sh r02_final > /dev/null 2>&1; python -c "
import json; [print(o) for o in sorted(json.load(open('gpurun_out/launches_r02_final.json')), key=lambda o: -o['share'])[:10]]"
# .ncu-rep files stay on the box (/tmp): only the CSV summaries travel back (gpurun_out is capped at 64 MiB)
timeout 500 ncu --set full --clock-control none -k regex:embed_fm -c 4 -f -o /tmp/embed_r02 \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_embed_r02.log 2>&1
timeout 500 ncu --set full --clock-control none -k regex:gemm_tc_kernel --launch-skip 6 -c 6 -f -o /tmp/gemm_r02 \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_gemm_r02.log 2>&1
python tools/ncu_summary.py /tmp/embed_r02.ncu-rep gpurun_out/embed_r02_ncu.csv
python tools/ncu_summary.py /tmp/gemm_r02.ncu-rep gpurun_out/gemm_r02_ncu.csv
ncu -i /tmp/embed_r02.ncu-rep --page details --csv > gpurun_out/embed_r02_ncu_details.csv 2>/dev/null
ncu -i /tmp/gemm_r02.ncu-rep --page details --csv > gpurun_out/gemm_r02_ncu_details.csv 2>/dev/null
du -sh gpurun_out
