#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 --maxfail=30 -p no:cacheprovider -rP -k "mag_full_graph_attention" > $OUT/r3c_pytest.log 2>&1
echo "pytest exit $?"; tail -3 $OUT/r3c_pytest.log | cut -c1-300; grep "^\[parity\] MAG full" $OUT/r3c_pytest.log | cut -c1-200
