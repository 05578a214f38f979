#!/bin/bash
mkdir -p gpurun_out
for k in "tc_dw_share=1" "tc_dw_share=1 embed_bwd_carveout=100" "tc_dw_share=1 embed_bwd_carveout=100 embed_bwd_unroll=1" "tc_dw_share=1 embed_bwd_carveout=50"; do
  timeout 200 python -u tools/probe_overlap.py $k 2>&1 | tail -1 | tee -a gpurun_out/r02q_probe_overlap.jsonl
done
