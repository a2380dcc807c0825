#!/usr/bin/env python
"""Turns the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py <tag> [<ncu-rep basename>]

  gpurun_out/<tag>_launches.csv   (ncu --metrics gpu__time_duration.sum)  -> profiles/<tag>_launches.md
  gpurun_out/<tag>_<rep>.ncu-rep  (ncu --set full)                        -> profiles/<tag>_<rep>_metrics.md
  gpurun_out/<tag>_bench.json                                             -> profiles/<tag>_bench.json
"""
import collections
import csv
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
           'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
           'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
           'launch__occupancy_limit_shared_mem', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
           'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
           'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
           'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
           'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
           'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def launches(tag):
    src = os.path.join(OUT, tag + '_launches.csv')
    if not os.path.exists(src):
        return
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi or not r[vi]:
            continue
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        a = agg.setdefault(r[ki], [0, 0.0, r[ui]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(PROF, tag + '_launches.md'), 'w') as f:
        f.write('# %s: per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none`\n\n' % tag)
        f.write('Command: `python bench.py --steps 2 --warmup 3 --no-others --no-cpu-baseline` (cold-cache, serialised '
                'launches: compare SHARES, not absolutes).\n\n| kernel | launches | total | avg | share |\n|---|---|---|---|---|\n')
        for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write('| `%s` | %d | %.1f %s | %.1f | %.3f |\n' % (n[:110], a[0], a[1], a[2], a[1] / a[0], a[1] / tot))


def full(tag, rep):
    """From gpurun_out/<tag>_<rep>.ncu-rep, or from gpurun_out/<tag>_<rep>_raw.csv when the report was exported with
    `ncu -i ... --page raw --csv` on the GPU box (gpurun_out is capped at 64 MiB, a source-level report is larger)."""
    src = os.path.join(OUT, '%s_%s.ncu-rep' % (tag, rep))
    pre = os.path.join(OUT, '%s_%s_raw.csv' % (tag, rep))
    if os.path.exists(src):
        raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    elif os.path.exists(pre):
        raw = open(pre).read()
    else:
        return
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(PROF, '%s_%s_metrics.md' % (tag, rep)), 'w') as f:
        f.write('# %s: `ncu --set full --clock-control none --import-source on` (%s.ncu-rep), per launch\n\n' % (tag, rep))
        for r in rows[2:]:
            f.write('## %s\n\n| metric | value | unit |\n|---|---|---|\n' % r[hdr.index('Kernel Name')])
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write('| %s | %s | %s |\n' % (m, r[i], units[i]))
            if 'dram__bytes_read.sum' in hdr:
                def val(name):
                    i = hdr.index(name)
                    x = float(r[i].replace(',', ''))
                    return x * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}.get(units[i], 1)
                f.write('| **traffic = dram read + write** | %.0f | byte |\n' % (val('dram__bytes_read.sum') + val('dram__bytes_write.sum')))
            f.write('\n')


if __name__ == '__main__':
    tag = sys.argv[1]
    rep = sys.argv[2] if len(sys.argv) > 2 else 'spmm'
    os.makedirs(PROF, exist_ok=True)
    launches(tag)
    full(tag, rep)
    for suffix in ('_bench.json', '_bench_reference.json'):
        b = os.path.join(OUT, tag + suffix)
        if os.path.exists(b) and os.path.getsize(b):
            shutil.copy(b, os.path.join(PROF, tag + suffix))
    print('profiles/ updated for', tag)
