"""Times the forward SpMM and the fused backward on column slabs of the MAG-shaped graph (all rows x F columns),
the per-rank launches of partition.feature_sliced_propagate at P = 128 / F ranks."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, '.')
from re_gnn_b200 import Graph, ops, synth  # noqa: E402


def timed(fn, n=10, w=3):
    if os.environ.get('PROBE_ONCE'):   # one launch per kernel: the ncu capture run
        n, w = 1, 0
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    dev = torch.device('cuda:0')
    d = synth.hetero_graph('mag', seed=0)
    n, r, e = d['num_nodes'], d['num_relations'], d['src'].size
    g = Graph(d['src'], d['dst'], n).to(dev)
    csr = g.csr()
    etv = g.etype_views(torch.as_tensor(d['etype']).to(dev), r)
    th = (torch.rand(r, 1, device=dev) + 0.5) / 100
    _, nrm = ops.wdeg_norm_fwd(csr, etv[0], th, 100.0, -0.5, counts=etv[2])
    out = {}
    for f in [int(a) for a in sys.argv[1:]] or [128, 64, 32, 16]:
        x, gg = torch.randn(n, f, device=dev), torch.randn(n, f, device=dev)
        y, dx = torch.empty_like(x), torch.empty_like(x)
        order = ops.row_order(csr) if (f <= ops.NARROW_FEAT and not os.environ.get('PROBE_NO_ORDER')) else None
        t_f = timed(lambda: ops.spmm(csr['indptr'], csr['indices'], etv[0], th, 100.0, nrm, nrm, x, out=y,
                                     split=csr.get('split'), order=order))
        t_b = timed(lambda: ops.spmm_bwd_fused(csr, etv[1], th, 100.0, nrm, x, gg, out=dx, want_xdx=True))
        alg = e * (4 * f + 9) + n * (4 * f + 8)
        out[f] = {'fwd_ms': round(t_f, 3), 'bwd_fused_ms': round(t_b, 3), 'fwd_GBs': round(alg / t_f / 1e6, 1)}
        print(f, out[f], flush=True)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
