#!/bin/bash
# Round-2 call 3 (one B200): gather microbenchmark (timing, then DRAM bytes under ncu), new full-size parity tests.
mkdir -p gpurun_out
timeout 300 tools/bin/mb_gather --stride 32 > gpurun_out/mb_gather_s32.jsonl 2> gpurun_out/mb_gather_s32.err
timeout 300 tools/bin/mb_gather --stride 16 > gpurun_out/mb_gather_s16.jsonl 2> gpurun_out/mb_gather_s16.err
cat gpurun_out/mb_gather_s32.jsonl gpurun_out/mb_gather_s16.jsonl | cut -c1-200
for st in 32 16; do
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum --clock-control none --csv \
   --log-file gpurun_out/mb_gather_s${st}_ncu.csv tools/bin/mb_gather --stride $st --iters 0 > gpurun_out/mb_gather_s${st}_ncu.out 2>&1
done
python -u -m pytest tests/test_gpu_fullsize_gemm.py tests/test_gpu_assembled_models.py tests/test_gpu_zz_next_rows.py -m gpu -q -x --timeout=900 -rf --tb=short \
    -p no:cacheprovider > gpurun_out/r02_newtests.log 2>&1
tail -15 gpurun_out/r02_newtests.log
timeout 300 python -u bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_n1_c5.log 2>&1; grep '^{' gpurun_out/r02_n1_c5.log | cut -c1-300
