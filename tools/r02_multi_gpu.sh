#!/bin/bash
# Round-2 multi-GPU call (gpurun --gpus N, N = 2 or 8): what round 1 could only check at world 1.
# Usage: gpurun --gpus 2 -- 'bash tools/r02_multi_gpu.sh 2'
N=${1:-2}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
python -u -m pytest tests/test_gpu_sharded.py -m gpu -v --timeout=1000 -rfP --tb=short -p no:cacheprovider \
    > gpurun_out/r02_world${N}_tests.log 2>&1; tail -2 gpurun_out/r02_world${N}_tests.log
grep -a " ok: \|SHARDED_OK\|Error\|assert" gpurun_out/r02_world${N}_tests.log | tail -14
B="bench.py --gpus $N --steps 20 --warmup 5"
$T --master-port 29511 $B                                   > gpurun_out/r02_n${N}_default.log 2>&1
$T --master-port 29512 $B --tune embed_fwd_linx_shard=1     > gpurun_out/r02_n${N}_linxshard.log 2>&1
$T --master-port 29513 $B --workload c5                     > gpurun_out/r02_n${N}_c5.log 2>&1
$T --master-port 29514 tools/bench_two_tower.py             > gpurun_out/r02_n${N}_c4.log 2>&1
grep '^{' gpurun_out/r02_n${N}_c4.log | tail -1 | cut -c1-300
for f in default linxshard c5; do grep '^{' gpurun_out/r02_n${N}_$f.log | tail -1 | cut -c1-400; done
