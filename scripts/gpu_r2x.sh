#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --maxfail=30 -p no:cacheprovider -k "gat or attention or layer_golden or model_golden or tiny or mag_regnn or regatv2 or v2" > $OUT/r2x_pytest.log 2>&1
echo "pytest exit $?"; tail -4 $OUT/r2x_pytest.log | cut -c1-300
for v in base e4 f3 ds4 ds6; do
  lib=variants/libregnn_$v.so; [ $v = base ] && lib=re_gnn_b200/lib/libregnn_b200.so
  echo "== $v"; REGNN_B200_LIB=$lib timeout 300 python scripts/attn_probe.py mag 8 16 2>&1 | grep -E "gatv2|per ABI" | tail -4
  REGNN_B200_LIB=$lib timeout 300 python scripts/attn_probe.py mag 2 64 2>&1 | grep -E "gatv2|per ABI" | tail -4
  REGNN_B200_LIB=$lib timeout 300 python scripts/attn_probe.py acm 8 64 2>&1 | grep -E "gatv2|per ABI" | tail -4
done | tee $OUT/r2x_probes.log
