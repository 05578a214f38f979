#!/bin/bash
# Final evidence of round 2, second session (one B200): the bench line with the CPU arm, the reference arm, the launch
# list, ncu --set full of the shipped embedding and GEMM kernels (CSV summaries travel, reports stay on the box).
mkdir -p gpurun_out
timeout 400 python -u bench.py --steps 30 --warmup 5 > gpurun_out/r02z_bench.log 2>&1; grep '^{' gpurun_out/r02z_bench.log | cut -c1-600
timeout 200 python -u bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02z_bench_ref.log 2>&1; grep '^{' gpurun_out/r02z_bench_ref.log | cut -c1-300
bash tools/ncu_launches.sh r02z > /dev/null 2>&1; python -c "
import json; [print(o) for o in sorted(json.load(open('gpurun_out/launches_r02z.json')), key=lambda o: -o['share'])[:12]]"
timeout 400 ncu --set full --clock-control none -k regex:embed_fm -c 4 -f -o /tmp/embed_r02z \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_embed_r02z.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:gemm_tc --launch-skip 6 -c 6 -f -o /tmp/gemm_r02z \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_gemm_r02z.log 2>&1
python tools/ncu_summary.py /tmp/embed_r02z.ncu-rep gpurun_out/embed_r02z_ncu.csv
python tools/ncu_summary.py /tmp/gemm_r02z.ncu-rep gpurun_out/gemm_r02z_ncu.csv
ncu -i /tmp/gemm_r02z.ncu-rep --page details --csv > gpurun_out/gemm_r02z_ncu_details.csv 2>/dev/null
ncu -i /tmp/embed_r02z.ncu-rep --page details --csv > gpurun_out/embed_r02z_ncu_details.csv 2>/dev/null
python -m pytest tests/test_gpu_models.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
du -sh gpurun_out
