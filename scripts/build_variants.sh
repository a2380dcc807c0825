#!/bin/bash
# A/B builds of one translation unit: [VARIANT_SRC=spmm|gat] scripts/build_variants.sh "tag:-DFLAG=1 -DOTHER=2" ...
# Each variant links the regular objects (build/obj, from `python -m re_gnn_b200.build`) with its own <src>.o
# into variants/libregnn_<tag>.so; run with REGNN_B200_LIB=variants/libregnn_<tag>.so.
set -e
cd "$(dirname "$0")/.."
SRC=${VARIANT_SRC:-spmm}
python -m re_gnn_b200.build > /dev/null
mkdir -p variants build/var
for spec in "$@"; do
  tag="${spec%%:*}"; flags="${spec#*:}"
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC $flags \
    -c re_gnn_b200/csrc/$SRC.cu -o build/var/${SRC}_$tag.o &
done
wait
for spec in "$@"; do
  tag="${spec%%:*}"
  objs=""
  for o in api csr_build spmm gat sample grouped_linear; do
    if [ $o = $SRC ]; then objs="$objs build/var/${SRC}_$tag.o"; else objs="$objs build/obj/$o.o"; fi
  done
  nvcc -shared -o variants/libregnn_$tag.so $objs -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -cudart static
  echo variants/libregnn_$tag.so
done
