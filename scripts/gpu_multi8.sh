#!/bin/bash
# 8-GPU pass (gpurun --gpus 8, charged 8x: keep it short): cross-rank parity + the bench line with phases_ms.
set -u
N=${1:-8}
TAG=${2:-r2h}
OUT=gpurun_out
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29511 scripts/check_multi_gpu.py 128 0.05 > $OUT/${TAG}_check_$N.log 2>&1
echo "check exit $?"; grep -E "rep|MULTI_GPU_CHECK|Error|error" $OUT/${TAG}_check_$N.log | head -20
timeout 600 $RUN --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${TAG}_bench_$N.json 2> $OUT/${TAG}_bench_$N.err
echo "bench exit $?"; cat $OUT/${TAG}_bench_$N.json | cut -c1-3500; grep -v "OMP_NUM_THREADS\|\*\*\*\*" $OUT/${TAG}_bench_$N.err | tail -15
