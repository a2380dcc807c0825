#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): cross-rank parity of the three exchange schemes, then the bench line at N ranks
# (CUDA-graph replay and eager), all outputs under gpurun_out/.
set -u
N=${1:-2}
TAG=${2:-r2g}
OUT=gpurun_out
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29511 scripts/check_multi_gpu.py 128 0.05 > $OUT/${TAG}_check_$N.log 2>&1
echo "check exit $?"; grep -E "rep|MULTI_GPU_CHECK|Error|error" $OUT/${TAG}_check_$N.log | head -20
timeout 900 $RUN --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${TAG}_bench_$N.json 2> $OUT/${TAG}_bench_$N.err
echo "bench exit $?"; cat $OUT/${TAG}_bench_$N.json | cut -c1-3000; grep -v "OMP_NUM_THREADS\|\*\*\*\*" $OUT/${TAG}_bench_$N.err | tail -15
timeout 900 $RUN --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --no-graph > $OUT/${TAG}_bench_${N}_eager.json 2> $OUT/${TAG}_bench_${N}_eager.err
echo "bench eager exit $?"; python - <<PY
import json
try:
    d=json.loads(open('$OUT/${TAG}_bench_${N}_eager.json').read().strip().splitlines()[-1])
    print('eager: ms_per_step', d['ms_per_step'], 'value', d['value'], 'phases', d.get('phases_ms'))
except Exception as e: print('no eager line', e)
PY
