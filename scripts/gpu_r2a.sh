#!/bin/bash
# Round-2 first GPU pass: all GPU tests (new parity tests included), attention operators at HBM scale, ncu captures.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --maxfail=25 -p no:cacheprovider -rP > $OUT/r2a_pytest.log 2>&1
echo "pytest exit $?"; tail -5 $OUT/r2a_pytest.log
grep "MAG F=" $OUT/r2a_pytest.log
timeout 300 python scripts/attn_probe.py mag 8 16 > $OUT/r2a_attn_mag_h8d16.log 2>&1; cat $OUT/r2a_attn_mag_h8d16.log
timeout 300 python scripts/attn_probe.py mag 2 64 > $OUT/r2a_attn_mag_h2d64.log 2>&1; cat $OUT/r2a_attn_mag_h2d64.log
timeout 300 python scripts/attn_probe.py acm 8 64 > $OUT/r2a_attn_acm_h8d64.log 2>&1; cat $OUT/r2a_attn_acm_h8d64.log
timeout 300 python scripts/op_times.py 128 > $OUT/r2a_op_times.log 2>&1; cat $OUT/r2a_op_times.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gat" -o $OUT/r2a_attn_mag python scripts/attn_probe.py mag 8 16 1 once > $OUT/r2a_ncu_attn.log 2>&1
echo "ncu attn exit $?"
CMD="python bench.py --steps 2 --warmup 3 --no-others --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmm_rowgroup_kernel<true|spmm_rowgroup_kernel<\(bool\)1" -s 3 -c 1 -o $OUT/r2a_spmm_bwd $CMD > $OUT/r2a_ncu_bwd.log 2>&1
echo "ncu bwd exit $?"
ls -la $OUT | tail -12
