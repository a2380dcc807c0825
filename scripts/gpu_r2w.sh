#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gatv2_bwd_edges_kernel|gatv2_bwd_dst_stream_kernel|colsum_stage" -c 3 -o /tmp/r2w_v2bwd python scripts/attn_probe.py mag 8 16 1 once > $OUT/r2w_ncu_bwd.log 2>&1
ncu -i /tmp/r2w_v2bwd.ncu-rep --page raw --csv > $OUT/r2w_v2bwd_raw.csv 2>/dev/null
ncu -i /tmp/r2w_v2bwd.ncu-rep --page source --csv > $OUT/r2w_v2bwd_source.csv 2>/dev/null
ls -la $OUT | tail -4
