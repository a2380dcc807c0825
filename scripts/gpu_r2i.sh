#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --maxfail=30 -p no:cacheprovider > $OUT/r2i_pytest.log 2>&1
echo "pytest exit $?"; tail -6 $OUT/r2i_pytest.log | cut -c1-300
/usr/bin/time -v timeout 1200 python bench.py > $OUT/r2i_bench.json 2> $OUT/r2i_bench.err
echo "bench exit $?"; cat $OUT/r2i_bench.json | cut -c1-9000; grep -E "Elapsed|Error|error" $OUT/r2i_bench.err | tail -5
timeout 300 python scripts/attn_probe.py mag 8 16 2>&1 | grep -E "gatv2|gat_fwd " | tee $OUT/r2i_attn.log
timeout 300 python scripts/attn_probe.py acm 8 64 2>&1 | grep -E "gatv2|gat_fwd " | tee -a $OUT/r2i_attn.log
