#!/bin/bash
# Slim end-of-session validation on the GPU box: GPU tests, smoke(), the default bench line (both arms) and the ncu
# launch list of the short bench command.  Outputs -> gpurun_out/<tag>_*.
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-others --no-cpu-baseline"
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --maxfail=25 -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; cut -c1-700 $OUT/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
echo "bench reference exit $?"
$CMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launches exit $?"
