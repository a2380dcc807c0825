#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q --timeout 600 --maxfail=30 -p no:cacheprovider -k "grouped_linear" > $OUT/r2k_pytest.log 2>&1
echo "pytest exit $?"; tail -6 $OUT/r2k_pytest.log | cut -c1-300; grep "parity.*grouped_linear out" $OUT/r2k_pytest.log | head
SECONDS=0
timeout 1200 python bench.py > $OUT/r2k_bench.json 2> $OUT/r2k_bench.err
echo "bench exit $? in $SECONDS s"; cat $OUT/r2k_bench.json | cut -c1-14000; grep -E "Error|error|Traceback" $OUT/r2k_bench.err | tail -5
