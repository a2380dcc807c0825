"""Runs forward+backward of one RE layer on a named synthetic graph a few times (profiling target for ncu).
    python scripts/layer_probe.py regat|regatv2|regcn|remixhop [iters]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import re_gnn_b200  # noqa: E402
from re_gnn_b200 import Graph, synth  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else 'regat'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
shape = {'regat': 'acm', 'regatv2': 'imdb', 'regcn': 'dblp', 'remixhop': 'imdb'}[kind]
d = synth.hetero_graph(shape)
dev = 'cuda:0'
g = Graph(d['src'], d['dst'], d['num_nodes']).to(dev)
et = torch.as_tensor(d['etype']).to(dev)
r = d['num_relations']
if kind == 'regat':
    mod, fin = re_gnn_b200.REGATConv(r, 100.0, 512, 64, 8, negative_slope=0.01, use_weight=False), 512
elif kind == 'regatv2':
    mod, fin = re_gnn_b200.REGATv2Conv(r, 100.0, 512, 64, 8, negative_slope=0.01, use_weight=False), 512
elif kind == 'regcn':
    mod, fin = re_gnn_b200.REGraphConv(r, 100.0, 64, 64, bias=False, weight=False), 64
else:
    mod, fin = re_gnn_b200.REMixHopConv(r, 100.0, 64, 64, p=[0, 1, 2]), 64
mod = mod.to(dev)
x = torch.randn(d['num_nodes'], fin, device=dev, requires_grad=True)
out = mod(g, x, et)
gout = torch.randn_like(out)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
for _ in range(iters):
    x.grad = None
    mod.zero_grad(set_to_none=True)
    flush.zero_()
    mod(g, x, et).backward(gout)
torch.cuda.synchronize()
print('ok', kind, d['num_nodes'], d['src'].size)
