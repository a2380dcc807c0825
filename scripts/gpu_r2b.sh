#!/bin/bash
# Round-2 second GPU pass: full default bench line, operator timings (folded norm gradient A/B), ncu captures exported
# to CSV on the box (the .ncu-rep files stay there: gpurun_out is capped at 64 MiB).
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv > $OUT/r2b_smi.txt
timeout 900 python bench.py > $OUT/r2b_bench.json 2> $OUT/r2b_bench.err
echo "bench exit $?"; cat $OUT/r2b_bench.json; tail -3 $OUT/r2b_bench.err
timeout 300 python scripts/op_times.py 128 > $OUT/r2b_op_times.log 2>&1; cat $OUT/r2b_op_times.log
timeout 300 python scripts/op_times.py 64 > $OUT/r2b_op_times64.log 2>&1; cat $OUT/r2b_op_times64.log
CMD="python bench.py --steps 2 --warmup 3 --no-others --no-cpu-baseline"
$CMD > $OUT/r2b_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r2b_launches.csv $CMD > $OUT/r2b_ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmm_rowgroup_kernel" -s 6 -c 2 -o /tmp/r2b_spmm $CMD > $OUT/r2b_ncu_spmm.log 2>&1
echo "ncu spmm exit $?"
ncu -i /tmp/r2b_spmm.ncu-rep --page raw --csv > $OUT/r2b_spmm_raw.csv 2>/dev/null
ncu -i /tmp/r2b_spmm.ncu-rep --page source --csv > $OUT/r2b_spmm_source.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gat_fwd_kernel|gat_bwd_dst_kernel|gat_bwd_src_kernel" -c 3 -o /tmp/r2b_gat python scripts/attn_probe.py mag 8 16 1 once > $OUT/r2b_ncu_gat.log 2>&1
echo "ncu gat exit $?"
ncu -i /tmp/r2b_gat.ncu-rep --page raw --csv > $OUT/r2b_gat_raw.csv 2>/dev/null
ncu -i /tmp/r2b_gat.ncu-rep --page source --csv > $OUT/r2b_gat_source.csv 2>/dev/null
ls -la $OUT /tmp/*.ncu-rep
du -sh $OUT
