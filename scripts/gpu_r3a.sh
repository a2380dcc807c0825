#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --maxfail=30 -p no:cacheprovider -k "propagate or narrow_row or column_slab or row_partitioned or mag_full_graph or many_relations or layer_golden or model_golden or folded or regcn" > $OUT/r3a_pytest.log 2>&1
echo "pytest exit $?"; tail -3 $OUT/r3a_pytest.log | cut -c1-300
for rep in 1 2; do
for v in base nopipe; do
  lib=variants/libregnn_$v.so; [ $v = base ] && lib=re_gnn_b200/lib/libregnn_b200.so
  echo "== $v"
  for f in 128 64 16; do REGNN_B200_LIB=$lib timeout 300 python scripts/op_times.py $f 2>&1 | grep -E "fused|row-group|MHz" ; done
done
done | tee $OUT/r3a_op_times.log
