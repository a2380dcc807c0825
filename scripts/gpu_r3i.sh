#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q --timeout 600 --maxfail=30 -p no:cacheprovider -k "world1" > $OUT/r3i_pytest.log 2>&1
echo "pytest exit $?"; tail -5 $OUT/r3i_pytest.log | cut -c1-300
