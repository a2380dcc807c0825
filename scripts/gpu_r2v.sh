#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --maxfail=30 -p no:cacheprovider -k "gat or attention or layer_golden or model_golden or tiny or mag_regnn or regatv2 or v2" > $OUT/r2v_pytest.log 2>&1
echo "pytest exit $?"; tail -3 $OUT/r2v_pytest.log | cut -c1-300
timeout 300 python scripts/attn_probe.py mag 8 16 2>&1 | grep -E "gatv2|per ABI" | tail -4
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gatv2_bwd_edges_kernel|gatv2_bwd_dst_stream_kernel|gat_bwd_bins_kernel|colsum_stage" -c 4 -o /tmp/r2v_v2bwd python scripts/attn_probe.py mag 8 16 1 once > $OUT/r2v_ncu_bwd.log 2>&1
ncu -i /tmp/r2v_v2bwd.ncu-rep --page raw --csv > $OUT/r2v_v2bwd_raw.csv 2>/dev/null
ncu -i /tmp/r2v_v2bwd.ncu-rep --page source --csv > $OUT/r2v_v2bwd_source.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gatv2_fwd_rg_kernel" -s 3 -c 2 -o /tmp/r2v_v2fwd python scripts/attn_probe.py mag 8 16 1 once > $OUT/r2v_ncu_fwd.log 2>&1
ncu -i /tmp/r2v_v2fwd.ncu-rep --page raw --csv > $OUT/r2v_v2fwd_raw.csv 2>/dev/null
ncu -i /tmp/r2v_v2fwd.ncu-rep --page source --csv > $OUT/r2v_v2fwd_source.csv 2>/dev/null
ls -la $OUT | tail -8
