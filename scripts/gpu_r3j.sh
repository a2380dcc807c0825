#!/bin/bash
set -u
N=${1:-2}
TAG=${2:-r3j}
OUT=gpurun_out
mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${TAG}_bench_$N.json 2> $OUT/${TAG}_bench_$N.err
echo "bench exit $?"; python - <<PY
import json
d=json.loads([l for l in open('$OUT/${TAG}_bench_$N.json') if l.startswith('{')][0])
print(d['value'], d['ms_per_step'], d['parity_check'])
print(json.dumps(d['others'], indent=1)[:2500])
PY
grep -v "OMP_NUM_THREADS\|\*\*\*\*" $OUT/${TAG}_bench_$N.err | tail -8
