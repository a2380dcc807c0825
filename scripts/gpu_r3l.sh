#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q --timeout 600 --maxfail=30 -p no:cacheprovider -k "regat or gat_layer or attention or layer_golden or model_golden or tiny or mag_regnn or named_shape or world1" > $OUT/r3l_pytest.log 2>&1
echo "pytest exit $?"; tail -3 $OUT/r3l_pytest.log | cut -c1-300
for v in base nofuse; do
  lib=variants/libregnn_$v.so; [ $v = base ] && lib=re_gnn_b200/lib/libregnn_b200.so
  echo "== $v"
  REGNN_B200_LIB=$lib timeout 300 python scripts/attn_probe.py mag 8 16 2>&1 | grep -E "gat_bwd|per ABI" | head -3
  REGNN_B200_LIB=$lib timeout 300 python scripts/attn_probe.py acm 8 64 2>&1 | grep -E "per ABI" | head -1
done | tee $OUT/r3l_reduce.log
