#!/usr/bin/env python
"""Static evidence for profiles/: per production kernel the resource usage (cuobjdump -res-usage), the SASS mnemonic
histogram and the mnemonics that identify the memory / tensor paths (LDG.E.128 gathers, UBLKCP = TMA bulk copies,
HMMA...TF32 = tensor-pipe MMA, MUFU.EX2 = fast exp, SHFL, REDUX), plus the hot-loop excerpt around the first 128-bit gather.

    python scripts/sass_excerpt.py [tag]     ->  profiles/<tag>_sass_excerpt.md, profiles/<tag>_resource_usage.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 're_gnn_b200', 'lib', 'libregnn_b200.so')
KERNELS = [   # (label, substring of the mangled name)
    ('spmm_rowgroup_kernel<false,32> (REGCN forward, headline)', 'spmm_rowgroup_kernelILb0ELi32ELb0'),
    ('spmm_rowgroup_kernel<true,32,DNORM> (REGCN fused backward + norm gradient)', 'spmm_rowgroup_kernelILb1ELi32ELb1'),
    ('spmm_rowgroup_kernel<false,4> (column-slab forward, 8 GPUs)', 'spmm_rowgroup_kernelILb0ELi4ELb0'),
    ('spmm_stream_kernel<1,4,true> (F > 128 fused backward, TMA-staged tile)', 'spmm_stream_kernelILi1ELi4ELb1'),
    ('gat_fwd_rg_kernel<32,false> (REGAT forward)', 'gat_fwd_rg_kernelILi32ELb0'),
    ('gat_bwd_edges_kernel<32,4,false> (REGAT backward gather pass, D = 16)', 'gat_bwd_edges_kernelILi32ELi4ELb0'),
    ('gatv2_fwd_rg_kernel<32,4,false> (REGATv2 forward, D = 16)', 'gatv2_fwd_rg_kernelILi32ELi4ELb0'),
    ('gatv2_bwd_edges_kernel<32,4,false> (REGATv2 backward, the one gather pass)', 'gatv2_bwd_edges_kernelILi32ELi4ELb0'),
    ('gatv2_bwd_dst_stream_kernel<32> (REGATv2 backward, destination-major streaming pass)', 'gatv2_bwd_dst_stream_kernelILi32'),
    ('grouped_linear_fwd_kernel (3 x TF32 mma.sync)', 'grouped_linear_fwd_kernel'),
    ('rows_to_slabs_kernel (peer-memory push)', 'rows_to_slabs_kernel'),
]
KEY = ['LDG.E.128', 'LDG.E.CONSTANT', 'LDG.E.128.CONSTANT', 'STG.E.128', 'STG.E', 'LDS', 'STS', 'SHFL', 'MUFU.EX2', 'UBLKCP',
       'SYNCS', 'HMMA', 'FFMA', 'VOTE', 'CCTL', 'BAR', 'STL', 'LDL']


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r2'
    res = subprocess.run(['cuobjdump', '-res-usage', LIB], capture_output=True, text=True).stdout
    usage = {}
    name = None
    for line in res.splitlines():
        m = re.search(r'Function (\S+):', line)
        if m:
            name = m.group(1)
        elif 'REG:' in line and name:
            usage[name] = line.strip()
    with open(os.path.join(ROOT, 'profiles', tag + '_resource_usage.md'), 'w') as f:
        f.write('# %s: `cuobjdump -res-usage` of every kernel in libregnn_b200.so (sm_100a)\n\n| kernel | resources |\n|---|---|\n' % tag)
        for k in sorted(usage):
            dem = subprocess.run(['c++filt', k], capture_output=True, text=True).stdout.strip()
            f.write('| `%s` | %s |\n' % (dem[:140], usage[k]))
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    funcs = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and re.match(r'\s+/\*[0-9a-f]{4}\*/', line):
            body = re.sub(r'^\s+/\*[0-9a-f]{4}\*/\s+', '', line)
            funcs[cur].append(re.sub(r'\s*/\*\s*0x[0-9a-f]+\s*\*/\s*$', '', body).rstrip(' ;'))
    with open(os.path.join(ROOT, 'profiles', tag + '_sass_excerpt.md'), 'w') as f:
        f.write('# %s: SASS evidence for the production kernels (`cuobjdump -sass`, sm_100a)\n\n'
                'Gather kernels: 128-bit coalesced loads (`LDG.E.128`), warp shuffles, `MUFU.EX2` for the softmax, no tensor '
                'instructions. `UBLKCP` = TMA bulk copy (`cp.async.bulk`) in the whole-warp fused backward; `HMMA...TF32` = '
                'tensor-pipe MMA in the grouped input projection only.\n\n' % tag)
        for label, sub in KERNELS:
            hit = [k for k in funcs if sub in k]
            if not hit:
                f.write('## %s\n\n(not found: %s)\n\n' % (label, sub))
                continue
            k = hit[0]
            ins = funcs[k]
            ops = [re.sub(r'^@!?U?P\d+\s+', '', i).split()[0] for i in ins if i]
            hist = collections.Counter(ops)
            f.write('## %s\n\n`%s`\n\n%s — %d SASS instructions\n\n' % (label, k[:120], usage.get(k, ''), len(ins)))
            f.write('key mnemonics: ' + ', '.join('`%s` x%d' % (m, sum(v for o, v in hist.items() if o.startswith(m)))
                                                  for m in KEY if any(o.startswith(m) for o in hist)) + '\n\n')
            f.write('top mnemonics: ' + ', '.join('%s %d' % kv for kv in hist.most_common(12)) + '\n\n')
            first = next((i for i, s in enumerate(ins) if 'LDG.E.128' in s or 'HMMA' in s or 'UBLKCP' in s), None)
            if first is not None:
                f.write('excerpt around the first `%s`:\n\n```\n%s\n```\n\n' % (
                    'LDG.E.128 / HMMA / UBLKCP', '\n'.join(ins[max(0, first - 6):first + 18])))
    print('wrote profiles/%s_sass_excerpt.md and profiles/%s_resource_usage.md' % (tag, tag))


if __name__ == '__main__':
    main()
