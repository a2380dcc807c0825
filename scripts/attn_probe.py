"""CUDA-event timing of the fused attention operators on the MAG-shaped graph (HBM-scale: the gathered matrix is
993 MB at H*D = 128) or an HGB-shaped one -- tuning aid for csrc/gat.cu.

    python scripts/attn_probe.py [graph=mag] [heads=8] [dim=16] [iters=10] [once]

`once` runs every operator exactly one time after a warm-up (for an ncu capture of the same command)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from re_gnn_b200 import Graph, ops, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'mag'
H = int(sys.argv[2]) if len(sys.argv) > 2 else 8
D = int(sys.argv[3]) if len(sys.argv) > 3 else 16
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 10
once = 'once' in sys.argv
d = synth.hetero_graph(name)
dev = 'cuda:0'
g = Graph(d['src'], d['dst'], d['num_nodes']).to(dev)
et = torch.as_tensor(d['etype']).to(dev)
n, r, e = d['num_nodes'], d['num_relations'], d['src'].size
csr = g.csr()
etv = g.etype_views(et, r)
theta = (torch.rand(r, H, device=dev) + 0.5) / 100.0
feat = torch.randn(n, H, D, device=dev) * 0.5
fd = torch.randn(n, H, D, device=dev) * 0.5
attn = torch.randn(H * D, device=dev) * 0.3
al, ar = torch.randn(1, H, D, device=dev) * 0.3, torch.randn(1, H, D, device=dev) * 0.3
el, er = (feat * al).sum(-1), (feat * ar).sum(-1)
gout = torch.randn(n, H, D, device=dev)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)   # 256 MB > L2


def t(label, fn, bytes_=None):
    for _ in range(1 if once else 3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(1 if once else iters):
        if name != 'mag':
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    extra = '' if bytes_ is None else '  %7.0f GB/s algorithmic (%.2f of 6554)' % (bytes_ / ms / 1e6, bytes_ / ms / 1e6 / 6554.2)
    print('%-30s %8.3f ms%s' % (label, ms, extra))
    return ms


HD = H * D
print('graph=%s N=%d E=%d H=%d D=%d' % (name, n, e, H, D))
# SURVEY 8(d) gather-model bytes
b_gat_fwd = e * (4 * HD + 4 * H + 4 + 1) + n * (4 * HD + 4 * H + 4 + 8 * H)
b_gat_bwd_dst = e * (4 * HD + 4 * H + 5) + n * (4 * HD + 12 * H) + 2 * 4 * e * H
b_gat_bwd_src = e * (4 * HD + 4 * H + 4 + 4) + n * (4 * HD + 4 * H) + 4 * e * H
b_v2_fwd = e * (4 * HD + 4 + 1) + n * (2 * 4 * HD + 8 * H + 4) + 4 * HD
out, rowmax, rowsum, _ = ops.gat_fwd(csr, etv[0], theta, 100.0, feat, el, er, 0.2)
t('gat_fwd', lambda: ops.gat_fwd(csr, etv[0], theta, 100.0, feat, el, er, 0.2), b_gat_fwd)
# one-gather-pass backward: stats (2 N HD reads) + edges (E x (4HD + 16H + 9 + 4H write)) + reduce (2 x 4EH reads)
b_gat_bwd = n * (2 * 4 * HD + 16 * H) + e * (4 * HD + 16 * H + 9 + 4 * H) + n * 2 * 4 * HD + 2 * 4 * e * H
t('gat_bwd (stats+edges+reduce)', lambda: ops.gat_bwd(csr, etv[0], etv[1], theta, 100.0, feat, el, er, 0.2, None, out,
                                                       rowmax, rowsum, gout), b_gat_bwd)
t('gat_bwd (+score gradients)', lambda: ops.gat_bwd(csr, etv[0], etv[1], theta, 100.0, feat, el, er, 0.2, None, out,
                                                     rowmax, rowsum, gout, attn_l=al, attn_r=ar), b_gat_bwd + n * 3 * 4 * HD)
t('attn_scores_fwd', lambda: ops.attn_scores_fwd(feat, al, ar), n * (4 * HD + 8 * H))
if not once:
    from re_gnn_b200 import _lib
    with _lib.Trace() as tr:
        for _ in range(5):
            ops.gat_bwd(csr, etv[0], etv[1], theta, 100.0, feat, el, er, 0.2, None, out, rowmax, rowsum, gout, attn_l=al, attn_r=ar)
    print('  per ABI call (ms):', {k: round(v, 3) for k, v in tr.summary(5).items()})
o2, m2, s2, _, sv2 = ops.gatv2_fwd(csr, etv[0], theta, 100.0, feat, fd, attn, 0.2, save=True)
t('gatv2_fwd (inference)', lambda: ops.gatv2_fwd(csr, etv[0], theta, 100.0, feat, fd, attn, 0.2), b_v2_fwd)
t('gatv2_fwd (training: + logits, sign masks)', lambda: ops.gatv2_fwd(csr, etv[0], theta, 100.0, feat, fd, attn, 0.2, save=True),
  b_v2_fwd + e * (4 * H + 16 * ((HD + 127) // 128)))
# stats (2 N HD reads) + edges (E x (4HD + 16H + 4H + 16 HG + 8 + 4H write)) + dst stream (E x (4H + 16 HG) + 2 N HD) + bins
b_v2_bwd = (n * (2 * 4 * HD + 16 * H) + e * (4 * HD + 24 * H + 16 * ((HD + 127) // 128) + 8) + n * 4 * HD * 2
            + e * (4 * H + 16 * ((HD + 127) // 128)) + n * 2 * 4 * HD + e * (4 * H + 1))
t('gatv2_bwd (stats+edges+dst)', lambda: ops.gatv2_bwd(csr, etv[0], theta, 100.0, feat, fd, attn, 0.2, None, o2, m2, s2, sv2, gout),
  b_v2_bwd)
if not once:
    from re_gnn_b200 import _lib
    with _lib.Trace() as tr:
        for _ in range(5):
            ops.gatv2_bwd(csr, etv[0], theta, 100.0, feat, fd, attn, 0.2, None, o2, m2, s2, sv2, gout)
    print('  per ABI call (ms):', {k: round(v, 3) for k, v in tr.summary(5).items()})
t('el/er (eager torch, 2 x mul+sum)', lambda: ((feat * al).sum(-1), (feat * ar).sum(-1)))
