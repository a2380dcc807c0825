#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
for cfg in "mag 8 16" "acm 8 64"; do
  timeout 300 python scripts/attn_probe.py $cfg 2>&1 | grep -v gatv2 | tee "$OUT/r2e_attn_${cfg// /_}.log"
done
timeout 300 python scripts/op_times.py 128 2>&1 | grep -E "fused|rowdot|row-group|MHz" | tee $OUT/r2e_op_times.log
timeout 300 python scripts/narrow_probe.py 64 32 16 2>&1 | tail -1 | tee $OUT/r2e_narrow.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gat_bwd_edges_kernel" -c 1 -o /tmp/r2e_gat python scripts/attn_probe.py mag 8 16 1 once > $OUT/r2e_ncu_gat.log 2>&1
ncu -i /tmp/r2e_gat.ncu-rep --page raw --csv > $OUT/r2e_gat_raw.csv 2>/dev/null
ncu -i /tmp/r2e_gat.ncu-rep --page source --csv > $OUT/r2e_gat_source.csv 2>/dev/null
timeout 900 python -m pytest tests -m gpu -q --timeout 900 --maxfail=30 -p no:cacheprovider -k "propagate or narrow_row or column_slab or row_partitioned or mag_full_graph_regcn or many_relations or layer_golden" > $OUT/r2e_pytest.log 2>&1
echo "pytest exit $?"; tail -4 $OUT/r2e_pytest.log | cut -c1-300
