#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 --maxfail=30 -p no:cacheprovider -k "regat or gat or attention or layer_golden or model_golden or mag_full_graph_regat or tiny_graphs or row_partitioned or mag_regnn or mag_attention" > $OUT/r2d_pytest.log 2>&1
echo "pytest exit $?"; tail -8 $OUT/r2d_pytest.log | cut -c1-300
for cfg in "mag 8 16" "mag 2 64" "acm 8 64" "acm 8 16"; do
  timeout 300 python scripts/attn_probe.py $cfg > "$OUT/r2d_attn_${cfg// /_}.log" 2>&1; cat "$OUT/r2d_attn_${cfg// /_}.log"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gat_bwd_edges_kernel|gat_fwd_rg_kernel" -c 2 -o /tmp/r2d_gat python scripts/attn_probe.py mag 8 16 1 once > $OUT/r2d_ncu_gat.log 2>&1
echo "ncu gat exit $?"
ncu -i /tmp/r2d_gat.ncu-rep --page raw --csv > $OUT/r2d_gat_raw.csv 2>/dev/null
ncu -i /tmp/r2d_gat.ncu-rep --page source --csv > $OUT/r2d_gat_source.csv 2>/dev/null
