#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q --timeout 600 --maxfail=30 -p no:cacheprovider -k "grouped_linear or model_golden or mag_regnn or saint" > $OUT/r2j_pytest.log 2>&1
echo "pytest exit $?"; tail -12 $OUT/r2j_pytest.log | cut -c1-300
SECONDS=0
timeout 1200 python bench.py > $OUT/r2j_bench.json 2> $OUT/r2j_bench.err
echo "bench exit $? in $SECONDS s"; cat $OUT/r2j_bench.json | cut -c1-12000; grep -E "Error|error|Traceback" $OUT/r2j_bench.err | tail -5
SECONDS=0
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/r2j_bench_reference.json 2>> $OUT/r2j_bench.err
echo "reference exit $? in $SECONDS s"; cat $OUT/r2j_bench_reference.json | cut -c1-1500
