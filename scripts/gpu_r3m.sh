#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider > $OUT/r3m_pytest.log 2>&1
echo "pytest exit $?"; tail -3 $OUT/r3m_pytest.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r3m_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/r3m_smoke.log
timeout 300 python bench.py > $OUT/r3m_bench.json 2> $OUT/r3m_bench.err; echo "bench exit $?"; cut -c1-300 $OUT/r3m_bench.json
