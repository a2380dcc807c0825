#!/bin/bash
# Round-end pass on one B200: every GPU test, smoke(), the default bench line + the reference arm, then the ncu launch list
# and `--set full` captures of the production kernels (exported to CSV on the box: gpurun_out is capped at 64 MiB).
set -u
TAG=${1:-r2z}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv > $OUT/${TAG}_smi.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --maxfail=30 -p no:cacheprovider -rP > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest.log | cut -c1-300
grep "^\[parity\]" $OUT/${TAG}_pytest.log > $OUT/${TAG}_parity_errors.txt; wc -l $OUT/${TAG}_parity_errors.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/${TAG}_smoke.log
SECONDS=0
timeout 1200 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $? in $SECONDS s"; cat $OUT/${TAG}_bench.json | cut -c1-2500; grep -E "Error|error|Traceback" $OUT/${TAG}_bench.err | tail -5
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
echo "reference exit $?"
CMD="python bench.py --steps 2 --warmup 3 --no-others --no-cpu-baseline"
$CMD > $OUT/${TAG}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmm_rowgroup_kernel" -s 6 -c 2 -o /tmp/${TAG}_spmm $CMD > $OUT/${TAG}_ncu_spmm.log 2>&1
ncu -i /tmp/${TAG}_spmm.ncu-rep --page raw --csv > $OUT/${TAG}_spmm_raw.csv 2>/dev/null
python scripts/attn_probe.py mag 8 16 1 once > $OUT/${TAG}_attn_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gat_fwd_rg_kernel|gat_bwd_edges_kernel|gatv2_fwd_rg_kernel|gatv2_bwd_edges_kernel|gatv2_bwd_dst_stream_kernel" -c 16 -o /tmp/${TAG}_gat python scripts/attn_probe.py mag 8 16 1 once > $OUT/${TAG}_ncu_gat.log 2>&1
echo "ncu gat exit $?"
ncu -i /tmp/${TAG}_gat.ncu-rep --page raw --csv > $OUT/${TAG}_gat_raw.csv 2>/dev/null
timeout 300 python scripts/attn_probe.py mag 8 16 > $OUT/${TAG}_attn_mag_8_16.log 2>&1; cat $OUT/${TAG}_attn_mag_8_16.log
timeout 300 python scripts/attn_probe.py mag 2 64 > $OUT/${TAG}_attn_mag_2_64.log 2>&1
timeout 300 python scripts/attn_probe.py acm 8 64 > $OUT/${TAG}_attn_acm_8_64.log 2>&1
du -sh $OUT
