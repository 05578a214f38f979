#!/bin/bash
mkdir -p gpurun_out
timeout 300 tools/bin/mb_gather --stride 32 --only r2_ > gpurun_out/mb_r2_s32.jsonl 2> gpurun_out/mb_r2_s32.err
cut -c1-200 gpurun_out/mb_r2_s32.jsonl; grep CTAs gpurun_out/mb_r2_s32.err
python -u -m pytest tests/test_gpu_sharded.py tests/test_gpu_models.py tests/test_gpu_zz_next_rows.py -m gpu -q -x --timeout=900 -rf --tb=short -n 4 \
    -p no:cacheprovider > gpurun_out/r02c_tests.log 2>&1
tail -6 gpurun_out/r02c_tests.log
