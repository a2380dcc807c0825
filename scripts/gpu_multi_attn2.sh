#!/bin/bash
set -u
N=${1:-2}
TAG=${2:-r3h}
OUT=gpurun_out
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29511 scripts/check_multi_gpu.py 128 0.05 > $OUT/${TAG}_check_$N.log 2>&1
echo "check exit $?"; grep -E "rep|MULTI_GPU_CHECK|Error|error" $OUT/${TAG}_check_$N.log | head -20
timeout 600 $RUN --master-port 29521 scripts/check_multi_gpu_attn.py 1.0 8 16 10 > $OUT/${TAG}_attn_$N.log 2> $OUT/${TAG}_attn_$N.err
echo "exit $?"; grep -E "^\{|MULTI_GPU_ATTN" $OUT/${TAG}_attn_$N.log | cut -c1-2500; grep -v "OMP_NUM\|\*\*\*" $OUT/${TAG}_attn_$N.err | tail -12
