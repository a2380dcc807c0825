"""Generates the golden fixtures in this directory from the REFERENCE'S OWN layer / model code.

Run in the build container (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden.py

How: ``oracle/dgl_stub`` (a pure-PyTorch restatement of the DGL primitives the reference calls) is put
on ``sys.path`` as ``dgl``, then the reference's unmodified ``layer/*.py`` and ``model/*.py`` are
imported from /root/reference and run in float64 on small seeded graphs.  Inputs, parameters,
outputs and all gradients are stored as ``<case>.npz``; constructor arguments go into the ``meta``
JSON string.  Tests rebuild the same module from ``meta``, load ``param::*`` and compare.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('REGNN_REFERENCE', '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'dgl_stub'))
sys.path.insert(0, REF)

import dgl  # noqa: E402  (the stub)
from layer import REGraphConv, REGATConv, REGATv2Conv, REMixHopConv, RESAGEConv, REGINConv  # noqa: E402  (the reference)
from model.REGCN import REGCN  # noqa: E402
from model.REGAT import REGAT  # noqa: E402
from model.REMixHop import REMixHop  # noqa: E402
from model.REGIN import REGIN  # noqa: E402

ACT = {'elu': F.elu, 'relu': F.relu, None: None}
ALPHA = 100.0


def small_graph(seed, n=48, e=260, r=7, self_loops=True, ntypes=3):
    """Random multigraph (duplicate edges allowed) with a few hub destinations and some nodes that
    have no in-edges before self loops; 1-based edge types; self loops get type r-ntypes+1+ntype."""
    rng = np.random.RandomState(seed)
    src = rng.randint(0, n, size=e)
    dst = np.where(rng.rand(e) < 0.35, rng.randint(0, 4, size=e), rng.randint(4, n - 6, size=e))
    et = rng.randint(1, r - ntypes + 1, size=e)
    keep = src != dst
    src, dst, et = src[keep], dst[keep], et[keep]
    if self_loops:
        loop = np.arange(n)
        ntype = loop * ntypes // n
        src = np.concatenate([src, loop])
        dst = np.concatenate([dst, loop])
        et = np.concatenate([et, r - ntypes + 1 + ntype])
    return src.astype(np.int64), dst.astype(np.int64), et.astype(np.int64), n


def make_graph(src, dst, n):
    return dgl.DGLGraph((src, dst), num_nodes=n)


def perturb_relations(module, rng):
    """Non-trivial relation embeddings: U(0.5,1.5)/alpha with one negative entry per table."""
    for name, p in module.named_parameters():
        if name.endswith('edge_weight'):
            v = torch.as_tensor(rng.uniform(0.5, 1.5, size=tuple(p.shape)) / ALPHA)
            v[0, 0] = -0.7 / ALPHA
            p.data.copy_(v)


ONLY = set(sys.argv[1:])   # `python tests/golden/make_golden.py <case> ...` rewrites only the named fixtures


def save_case(name, module, graph_arrays, inputs, run, meta):
    if ONLY and name not in ONLY:
        return
    src, dst, et, n = graph_arrays
    g = make_graph(src, dst, n)
    module = module.double()
    for t in inputs.values():
        t.requires_grad_(True)
    out = run(module, g, torch.as_tensor(et), inputs)
    outs = out if isinstance(out, tuple) else (out,)
    rng = np.random.RandomState(7)
    gout = torch.as_tensor(rng.randn(*outs[0].shape))
    outs[0].backward(gout)
    blob = dict(src=src, dst=dst, etype=et, num_nodes=np.int64(n), gout=gout.numpy(),
                meta=np.array(json.dumps(meta)))
    for i, o in enumerate(outs):
        blob['out%d' % i] = o.detach().numpy()
    for k, t in inputs.items():
        blob['in::' + k] = t.detach().numpy()
        blob['gin::' + k] = t.grad.numpy()
    for k, v in module.state_dict().items():      # parameters (shared ones under every alias) + buffers
        blob['param::' + k] = v.detach().numpy()
    for k, p in module.named_parameters():
        blob['grad::' + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **blob)
    print('%-28s out %s  params %d' % (name, tuple(outs[0].shape), sum(p.numel() for p in module.parameters())))


def layer_run(module, g, et, inputs):
    return module(g, inputs['x'], et)


def main():
    torch.manual_seed(123)
    rng = np.random.RandomState(123)
    R = 7
    G1 = small_graph(1)
    G2 = small_graph(2, self_loops=False)   # zero-in-degree rows present
    x16 = lambda n: torch.as_tensor(rng.randn(n, 16))  # noqa: E731

    # ---- REGraphConv -------------------------------------------------------------------------
    for name, kw, graph in [
        ('regcn_plain', dict(in_feats=16, out_feats=16, bias=False, activation=None, weight=False), G1),
        ('regcn_weight_bias_elu', dict(in_feats=16, out_feats=16, activation='elu'), G1),
        ('regcn_project_first', dict(in_feats=16, out_feats=8, activation=None), G1),
        ('regcn_nonorm_zero_indeg', dict(in_feats=16, out_feats=20, norm=False), G2),
        ('regcn_zero_indeg', dict(in_feats=16, out_feats=16, weight=False, bias=False), G2),
    ]:
        k = dict(kw)
        k['activation'] = ACT[k.get('activation')]
        m = REGraphConv(R, ALPHA, **k)
        perturb_relations(m, rng)
        save_case(name, m, graph, {'x': x16(graph[3])}, layer_run, dict(kind='REGraphConv', R=R, alpha=ALPHA, kw=kw))

    # ---- RESAGEConv / REGINConv ("next" row f-3: same SpMM, one-sided norm with exponent -1) ---
    for name, kw in [('resage_weight_elu', dict(in_feats=16, out_feats=16, activation='elu')),
                     ('resage_plain', dict(in_feats=16, out_feats=16, weight=False, bias=False))]:
        k = dict(kw)
        k['activation'] = ACT[k.get('activation')]
        m = RESAGEConv(R, ALPHA, **k)
        if m.weight_root is not None:
            m.weight_root.data.zero_()          # the reference leaves it uninitialised (and unused)
        perturb_relations(m, rng)
        save_case(name, m, G1, {'x': x16(G1[3])}, layer_run, dict(kind='RESAGEConv', R=R, alpha=ALPHA, kw=kw))
    m = REGINConv(R, ALPHA, torch.nn.Linear(16, 12), 'sum', 0, False, F.elu)
    perturb_relations(m, rng)
    save_case('regin_linear_elu', m, G1, {'x': x16(G1[3])}, layer_run,
              dict(kind='REGINConv', R=R, alpha=ALPHA, kw=dict(activation='elu'), apply_linear=[16, 12]))

    # ---- REMixHopConv ------------------------------------------------------------------------
    for name, kw in [('remixhop_p012', dict(in_feats=16, out_feats=8, p=[0, 1, 2])),
                     ('remixhop_p02_elu', dict(in_feats=16, out_feats=8, p=[0, 2], activation='elu'))]:
        k = dict(kw)
        k['activation'] = ACT[k.get('activation')]
        m = REMixHopConv(R, ALPHA, **k)
        perturb_relations(m, rng)
        save_case(name, m, G1, {'x': x16(G1[3])}, layer_run, dict(kind='REMixHopConv', R=R, alpha=ALPHA, kw=kw))

    # ---- REGATConv ---------------------------------------------------------------------------
    for name, kw, graph, fin in [
        ('regat_h4d8', dict(in_feats=16, out_feats=8, num_heads=4, negative_slope=0.01), G1, 16),
        ('regat_noweight_res', dict(in_feats=32, out_feats=8, num_heads=4, negative_slope=0.2, residual=True,
                                    use_weight=False, activation='elu'), G1, 32),
        ('regat_zero_indeg', dict(in_feats=16, out_feats=4, num_heads=2), G2, 16),
        ('regat_h1d64', dict(in_feats=16, out_feats=64, num_heads=1, negative_slope=0.01), G1, 16),
    ]:
        k = dict(kw)
        k['activation'] = ACT[k.get('activation')]
        m = REGATConv(R, ALPHA, **k)
        perturb_relations(m, rng)
        save_case(name, m, graph, {'x': torch.as_tensor(rng.randn(graph[3], fin))}, layer_run,
                  dict(kind='REGATConv', R=R, alpha=ALPHA, kw=kw))

    # REGATConv with edge_feats=None (no relation term)
    m = REGATConv(R, ALPHA, 16, 8, 2)
    save_case('regat_no_etype', m, G1, {'x': x16(G1[3])}, lambda mod, g, et, i: mod(g, i['x'], None),
              dict(kind='REGATConv', R=R, alpha=ALPHA, kw=dict(in_feats=16, out_feats=8, num_heads=2), no_etype=True))

    # ---- REGATv2Conv -------------------------------------------------------------------------
    for name, kw in [
        ('regatv2_h4d8', dict(in_feats=16, out_feats=8, num_heads=4, negative_slope=0.01)),
        ('regatv2_shared_res', dict(in_feats=16, out_feats=8, num_heads=2, share_weights=True, residual=True)),
        ('regatv2_noweight', dict(in_feats=32, out_feats=16, num_heads=2, use_weight=False, activation='elu')),
    ]:
        k = dict(kw)
        k['activation'] = ACT[k.get('activation')]
        m = REGATv2Conv(R, ALPHA, **k)
        perturb_relations(m, rng)
        # perturb the zero biases so their gradients are exercised
        for n_, p in m.named_parameters():
            if n_.endswith('.bias'):
                p.data.copy_(torch.as_tensor(rng.randn(*p.shape) * 0.1))
        fin = kw['in_feats']
        save_case(name, m, G1, {'x': torch.as_tensor(rng.randn(G1[3], fin))},
                  lambda mod, g, et, i: mod(g, i['x'], et, get_attention=True),
                  dict(kind='REGATv2Conv', R=R, alpha=ALPHA, kw=kw, get_attention=True))

    # ---- models (the three callers) ----------------------------------------------------------
    src, dst, et, n = G1
    sizes = [16, 16, 16]          # three node types of 16 nodes each (n = 48)
    dims = [5, 9, 7]
    feats = lambda: {'f%d' % i: torch.as_tensor(rng.randn(s, d)) for i, (s, d) in enumerate(zip(sizes, dims))}  # noqa: E731

    def model_run(mod, g, et_, inputs):
        mod.g = g
        return mod([inputs['f0'], inputs['f1'], inputs['f2']], et_)

    g0 = make_graph(src, dst, n)
    for name, ctor, meta in [
        ('model_regcn_2layer', lambda: REGCN(g0, R, ALPHA, 16, 16, 4, 2, F.elu, 0.0, dims),
         dict(kind='REGCN', args=[R, ALPHA, 16, 16, 4, 2, 'elu', 0.0, dims])),
        ('model_regcn_3layer', lambda: REGCN(g0, R, ALPHA, 16, 16, 4, 3, F.elu, 0.0, dims),
         dict(kind='REGCN', args=[R, ALPHA, 16, 16, 4, 3, 'elu', 0.0, dims])),
        ('model_regat', lambda: REGAT(g0, R, ALPHA, 2, 16, 16, 3, [2, 2, 1], F.elu, 0.0, 0.0, 0.01, False, dims),
         dict(kind='REGAT', args=[R, ALPHA, 2, 16, 16, 3, [2, 2, 1], 'elu', 0.0, 0.0, 0.01, False, dims])),
        ('model_regatv2', lambda: REGAT(g0, R, ALPHA, 2, 16, 16, 3, [2, 2, 1], F.elu, 0.0, 0.0, 0.01, False, dims,
                                        use_gatv2=True),
         dict(kind='REGAT', args=[R, ALPHA, 2, 16, 16, 3, [2, 2, 1], 'elu', 0.0, 0.0, 0.01, False, dims],
              use_gatv2=True)),
        ('model_remixhop', lambda: REMixHop(g0, R, ALPHA, 16, 16, 4, 2, dims, activation=F.elu),
         dict(kind='REMixHop', args=[R, ALPHA, 16, 16, 4, 2, dims], activation='elu')),
    ]:
        m = ctor()
        perturb_relations(m, rng)
        save_case(name, m, G1, feats(), model_run, meta)

    # the fourth caller (model/REGIN.py), added later with its own generator state so that it can be produced alone
    # (`python tests/golden/make_golden.py model_regin`) without touching the fixtures above
    torch.manual_seed(321)
    rng = np.random.RandomState(321)
    m = REGIN(g0, R, ALPHA, 16, 16, 4, 2, F.elu, 0.0, dims)
    perturb_relations(m, rng)
    save_case('model_regin', m, G1, feats(), model_run, dict(kind='REGIN', args=[R, ALPHA, 16, 16, 4, 2, 'elu', 0.0, dims]))


if __name__ == '__main__':
    main()
