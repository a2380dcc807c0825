#!/bin/bash
# Run on the GPU box (through gpurun): GPU tests, the default bench line, then the ncu launch list and one
# `--set full` capture of the dominant kernel for the same short bench command.  Outputs -> gpurun_out/.
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-others --no-cpu-baseline"
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --maxfail=25 -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -4 $OUT/${TAG}_pytest.log
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; cat $OUT/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
echo "bench reference exit $?"
$CMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launches exit $?"
$CMD > $OUT/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"spmm_(stream|rowgroup)_kernel" -s 4 -c 4 -o $OUT/${TAG}_spmm $CMD > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la $OUT
