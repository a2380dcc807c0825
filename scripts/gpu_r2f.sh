#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
for v in base gb5 gb6; do
  lib=variants/libregnn_$v.so; [ $v = base ] && lib=re_gnn_b200/lib/libregnn_b200.so
  echo "== $v"; REGNN_B200_LIB=$lib timeout 300 python scripts/attn_probe.py mag 8 16 2>&1 | grep -E "gat_bwd|per ABI|gat_fwd "
  REGNN_B200_LIB=$lib timeout 300 python scripts/attn_probe.py acm 8 64 2>&1 | grep -E "gat_bwd|per ABI|gat_fwd "
done | tee $OUT/r2f_gat_variants.log
timeout 300 python scripts/attn_probe.py mag 2 64 2>&1 | grep -v gatv2 | tee $OUT/r2f_attn_mag_2_64.log
for v in base c1; do
  lib=variants/libregnn_$v.so; [ $v = base ] && lib=re_gnn_b200/lib/libregnn_b200.so
  echo "== $v"; REGNN_B200_LIB=$lib timeout 300 python scripts/op_times.py 128 2>&1 | grep -E "fused|row-group"
  REGNN_B200_LIB=$lib timeout 300 python scripts/op_times.py 64 2>&1 | grep -E "fused|row-group"
done | tee $OUT/r2f_spmm_variants.log
