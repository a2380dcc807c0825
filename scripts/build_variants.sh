#!/bin/bash
# A/B builds of the SpMM translation unit: scripts/build_variants.sh "tag:-DFLAG=1 -DOTHER=2" ...
# Each variant links the regular objects (build/obj, from `python -m re_gnn_b200.build`) with its own spmm.o
# into variants/libregnn_<tag>.so; run with REGNN_B200_LIB=variants/libregnn_<tag>.so.
set -e
cd "$(dirname "$0")/.."
python -m re_gnn_b200.build > /dev/null
mkdir -p variants build/var
for spec in "$@"; do
  tag="${spec%%:*}"; flags="${spec#*:}"
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC $flags \
    -c re_gnn_b200/csrc/spmm.cu -o build/var/spmm_$tag.o &
done
wait
for spec in "$@"; do
  tag="${spec%%:*}"
  nvcc -shared -o variants/libregnn_$tag.so build/obj/api.o build/obj/csr_build.o build/var/spmm_$tag.o build/obj/gat.o \
    build/obj/sample.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -cudart static
  echo variants/libregnn_$tag.so
done
