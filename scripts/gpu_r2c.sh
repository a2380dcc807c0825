#!/bin/bash
# Round-2 third GPU pass: new row-group REGAT kernels (parity + timings), A/B builds of the fused SpMM backward.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 --maxfail=30 -p no:cacheprovider -k "regat or gat or attention or layer_golden or model_golden or mag_full_graph_regat or narrow_row or column_slab or tiny_graphs" > $OUT/r2c_pytest.log 2>&1
echo "pytest exit $?"; tail -15 $OUT/r2c_pytest.log | cut -c1-300
for cfg in "mag 8 16" "mag 2 64" "acm 8 64" "acm 8 16"; do
  timeout 300 python scripts/attn_probe.py $cfg > "$OUT/r2c_attn_${cfg// /_}.log" 2>&1; cat "$OUT/r2c_attn_${cfg// /_}.log"
done
for v in base b5u4 b6u4 b6u2 b5u2 b3u4; do
  lib=variants/libregnn_$v.so; [ $v = base ] && lib=re_gnn_b200/lib/libregnn_b200.so
  echo "== $v"; REGNN_B200_LIB=$lib timeout 300 python scripts/narrow_probe.py 128 16 2>&1 | tail -1
done | tee $OUT/r2c_variants.log
