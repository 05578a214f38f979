#!/bin/bash
mkdir -p gpurun_out
for k in "" "gemm_bn=128" "gemm_bn=256" "tc_l2_promo=128"; do
  timeout 200 python -u tools/bench_gemm_shapes.py $k 2>&1 | grep '^{' | tee -a gpurun_out/r02t_gemm_shapes.jsonl
done
