#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 --maxfail=30 -p no:cacheprovider -rP -k "mag_full_graph_attention or model_golden" > $OUT/r3b_pytest.log 2>&1
echo "pytest exit $?"; tail -3 $OUT/r3b_pytest.log | cut -c1-300; grep "^\[parity\] MAG full" $OUT/r3b_pytest.log | cut -c1-200
timeout 600 python scripts/ns_profile.py > $OUT/r3b_ns_profile.log 2>&1; grep -E "sample only|full step" $OUT/r3b_ns_profile.log; sed -n '/Self CPU %/,$p' $OUT/r3b_ns_profile.log | tail -24 | cut -c1-200
