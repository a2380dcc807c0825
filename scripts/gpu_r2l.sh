#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q --timeout 600 --maxfail=30 -p no:cacheprovider -k "grouped_linear or mag_regnn" > $OUT/r2l_pytest.log 2>&1
echo "pytest exit $?"; tail -4 $OUT/r2l_pytest.log | cut -c1-300
for gi in 1 0; do
  echo "== REGNN_GROUPED_INPUT=$gi"
  REGNN_GROUPED_INPUT=$gi timeout 600 python bench.py --workload mag_ns --steps 30 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'steps_per_s', d['steps_per_s'], 'launches', d['gpu_launches'])"
  REGNN_GROUPED_INPUT=$gi timeout 600 python bench.py --workload mag_saint --steps 30 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('saint ms_per_step', d['ms_per_step'], 'steps_per_s', d['steps_per_s'])"
done | tee $OUT/r2l_ns.log
