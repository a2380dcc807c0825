"""CUDA-event timing of every operator of the REGCN RE-layer on the MAG-shaped graph (tuning aid)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from re_gnn_b200 import Graph, ops, synth  # noqa: E402

feat = int(sys.argv[1]) if len(sys.argv) > 1 else 128
d = synth.hetero_graph('mag')
dev = 'cuda:0'
g = Graph(d['src'], d['dst'], d['num_nodes']).to(dev)
et = torch.as_tensor(d['etype']).to(dev)
n, r, e = d['num_nodes'], d['num_relations'], d['src'].size
csr = g.csr()
etv = g.etype_views(et, r)
theta = (torch.rand(r, 1, device=dev) + 0.5) / 100.0
x = torch.randn(n, feat, device=dev)
gout = torch.randn(n, feat, device=dev)
deg, norm = ops.wdeg_norm_fwd(csr, etv[0], theta, 100.0, -0.5, counts=etv[2])
y = ops.spmm(csr['indptr'], csr['indices'], etv[0], theta, 100.0, norm, norm, x, split=csr.get('split'))
dx, dth, xdx = ops.spmm_bwd_fused(csr, etv[1], theta, 100.0, norm, x, gout, want_xdx=True)
dn = ops.rowdot_norm_bwd(norm, x, y, gout, dx, xdx=xdx)


def t(name, fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    print('%-34s %8.3f ms' % (name, a.elapsed_time(b) / iters))


import subprocess  # noqa: E402
print(subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu', '--format=csv,noheader'],
                     capture_output=True, text=True).stdout.strip())
print('N=%d E=%d F=%d' % (n, e, feat))
t('wdeg_norm_fwd (counts)', lambda: ops.wdeg_norm_fwd(csr, etv[0], theta, 100.0, -0.5, counts=etv[2]))
t('wdeg_norm_fwd (slots)', lambda: ops.wdeg_norm_fwd(csr, etv[0], theta, 100.0, -0.5))
t('spmm fwd (stream kernel)', lambda: ops.spmm(csr['indptr'], csr['indices'], etv[0], theta, 100.0, norm, norm, x, split=csr.get('split')))
t('spmm fwd (row-group kernel)', lambda: ops.spmm(csr['indptr'], csr['indices'], etv[0], theta, 100.0, norm, norm, x, split=csr.get('split'), order=ops.row_order(csr)))
t('spmm transposed (plain bwd_x)', lambda: ops.spmm(csr['indptr_t'], csr['indices_t'], etv[1], theta, 100.0, norm, norm, gout, split=csr.get('split_t')))
t('spmm_bwd_fused (no xdx)', lambda: ops.spmm_bwd_fused(csr, etv[1], theta, 100.0, norm, x, gout))
t('spmm_bwd_fused (+xdx)', lambda: ops.spmm_bwd_fused(csr, etv[1], theta, 100.0, norm, x, gout, want_xdx=True))
t('spmm_bwd_fused (+folded d_norm)', lambda: ops.spmm_bwd_fused(csr, etv[1], theta, 100.0, norm, x, gout, y=y, want_dnorm=True))
t('spmm_bwd_w (two-pass variant)', lambda: ops.spmm_bwd_w(csr, etv[0], theta, 100.0, norm, x, y, gout, dx, split=csr.get('split')))
t('rowdot_norm_bwd (Y,G,X,dX)', lambda: ops.rowdot_norm_bwd(norm, x, y, gout, dx))
t('rowdot_norm_bwd (xdx given)', lambda: ops.rowdot_norm_bwd(norm, x, y, gout, dx, xdx=xdx))
t('wdeg_norm_bwd (counts)', lambda: ops.wdeg_norm_bwd(csr, etv[0], theta, 100.0, -0.5, deg, dn, counts=etv[2]))
t('wdeg_norm_bwd (slots)', lambda: ops.wdeg_norm_bwd(csr, etv[0], theta, 100.0, -0.5, deg, dn))
t('copy [N,F] (torch)', lambda: y.copy_(x))
print(subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu', '--format=csv,noheader'],
                     capture_output=True, text=True).stdout.strip())
