#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
for rep in 1 2; do
for v in base s4 s4e; do
  lib=variants/libregnn_$v.so; [ $v = base ] && lib=re_gnn_b200/lib/libregnn_b200.so
  echo "== $v"; REGNN_B200_LIB=$lib timeout 300 python scripts/attn_probe.py mag 8 16 2>&1 | grep -E "per ABI" | tail -1
  REGNN_B200_LIB=$lib timeout 300 python scripts/attn_probe.py mag 2 64 2>&1 | grep -E "per ABI" | tail -1
done
done | tee $OUT/r2y_probes.log
