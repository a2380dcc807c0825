#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
for rep in 1 2; do
for v in base u2b4 u2b5 u2b6 u2b8; do
  lib=variants/libregnn_$v.so; [ $v = base ] && lib=re_gnn_b200/lib/libregnn_b200.so
  echo "== $v"
  for f in 128 64 16; do REGNN_B200_LIB=$lib timeout 300 python scripts/op_times.py $f 2>&1 | grep -E "fused" ; done
done
done | tee $OUT/r3e_op_times.log
