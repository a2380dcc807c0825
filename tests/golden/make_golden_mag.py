"""Golden fixtures for the MAG-stack layers, generated from the REFERENCE'S OWN ``mag/regnn_layers.py``.

Run in the build container (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden_mag.py

How: ``oracle/pyg_stub`` (a pure-PyTorch restatement of the PyG / torch-scatter primitives the file uses) is put on
``sys.path``, then the reference's unmodified ``mag/regnn_layers.py`` (and the ``mag/utils.py`` it imports) run in
float64 on small seeded bipartite blocks of the shape PyG's NeighborSampler hands to the layers (targets are the
first ``n_dst`` sources).  Inputs, parameters, outputs and all gradients go to ``tests/golden/mag/<case>.npz``.
"""
import json
import zlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('REGNN_REFERENCE', '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'pyg_stub'))
sys.path.insert(0, os.path.join(REF, 'mag'))
torch.set_default_dtype(torch.float64)   # the reference builds its one-hot edge features with the default dtype

import regnn_layers as RL  # noqa: E402  (the reference)

NUM_NODE_TYPES, NUM_EDGE_TYPES, ALPHA = 3, 5, 100.0


def block(seed, n_src=60, n_dst=25, e=300, empty_targets=True):
    """Random bipartite multigraph: duplicate edges, two hub targets, (optionally) targets without in-edges."""
    rng = np.random.RandomState(seed)
    src = rng.randint(0, n_src, size=e)
    lo = 5 if empty_targets else 0
    dst = np.where(rng.rand(e) < 0.3, rng.randint(lo, lo + 2, size=e), rng.randint(lo, n_dst, size=e))
    et = rng.randint(0, NUM_EDGE_TYPES, size=e)
    tnt = rng.randint(0, NUM_NODE_TYPES, size=n_dst)
    return (np.stack([src, dst]).astype(np.int64), et.astype(np.int64), tnt.astype(np.int64), n_src, n_dst)


def save_case(name, kind, kw, blk, tuple_input=True):
    edge_index, et, tnt, n_src, n_dst = blk
    rng = np.random.RandomState(zlib.crc32(name.encode()) % (2 ** 31))
    torch.manual_seed(11)
    mod = getattr(RL, kind)(16, 8, NUM_NODE_TYPES, NUM_EDGE_TYPES, **kw)
    with torch.no_grad():      # non-trivial relation embeddings (one negative entry) and bias
        mod.relation_weight.copy_(torch.as_tensor(rng.uniform(0.5, 1.5, size=tuple(mod.relation_weight.shape)) / ALPHA))
        mod.relation_weight.view(-1)[0] = -0.7 / ALPHA
        mod.bias.copy_(torch.as_tensor(rng.randn(*mod.bias.shape) * 0.1))
    x_src = torch.as_tensor(rng.randn(n_src, 16)).requires_grad_(True)
    x_in = (x_src, x_src[:n_dst]) if tuple_input else x_src
    ew = None
    if kind == 'REGCNConv':   # also record the normalised edge weights the layer returns on request (:119-126, :137-138)
        out, ew, _ = mod(x_in, torch.as_tensor(edge_index), torch.as_tensor(et), torch.as_tensor(tnt), return_weights=True)
    else:
        out = mod(x_in, torch.as_tensor(edge_index), torch.as_tensor(et), torch.as_tensor(tnt))
    gout = torch.as_tensor(np.random.RandomState(7).randn(*out.shape))
    out.backward(gout)
    blob = dict(edge_index=edge_index, edge_type=et, target_node_type=tnt, n_src=np.int64(n_src), n_dst=np.int64(n_dst),
                x_src=x_src.detach().numpy(), gx_src=x_src.grad.numpy(), out=out.detach().numpy(), gout=gout.numpy(),
                meta=np.array(json.dumps(dict(kind=kind, kw=kw, tuple_input=tuple_input, in_channels=16, out_channels=8,
                                              num_node_types=NUM_NODE_TYPES, num_edge_types=NUM_EDGE_TYPES))))
    if ew is not None:
        blob['ew'] = ew.detach().numpy()
    for k, v in mod.state_dict().items():
        blob['param::' + k] = v.detach().numpy()
    for k, p in mod.named_parameters():
        blob['grad::' + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    np.savez_compressed(os.path.join(HERE, 'mag', name + '.npz'), **blob)
    print('%-30s out %s' % (name, tuple(out.shape)))


def reference_saint_conv():
    """mag/regnn_saint.py is a script (argument parsing and dataset loading at import), so its ``REGCNConv`` class is
    lifted out by source range -- the class body is the reference's text, unmodified -- and executed in a namespace
    holding exactly the names the script imports for it."""
    import ast
    import torch.nn.functional as F
    from torch.nn import Parameter, init
    from torch_geometric.nn import MessagePassing
    from utils import softmax, weighted_degree          # the reference's mag/utils.py
    path = os.path.join(REF, 'mag', 'regnn_saint.py')
    text = open(path).read()
    node = [n for n in ast.parse(text).body if isinstance(n, ast.ClassDef) and n.name == 'REGCNConv'][0]
    ns = dict(torch=torch, F=F, Parameter=Parameter, init=init, MessagePassing=MessagePassing, softmax=softmax,
              weighted_degree=weighted_degree)
    exec(compile(ast.get_source_segment(text, node), path, 'exec'), ns)
    return ns['REGCNConv']


def save_saint_case(name, n, e, seed, tuple_input, kw=None, train=False):
    """``kw``: extra constructor keywords (``use_softmax``, ``dropout``); ``train``: run in training mode, so that
    ``F.dropout`` acts on the edge weights -- the mask it drew is recovered from the returned ``ew`` (weights are
    positive, a dropped edge returns exactly 0)."""
    kw = kw or {}
    rng = np.random.RandomState(seed)
    dst = np.where(rng.rand(e) < 0.3, rng.randint(0, 2, size=e), rng.randint(0, n - 6, size=e))   # hubs; 6 rows w/o in-edges
    edge_index = np.stack([rng.randint(0, n, size=e), dst]).astype(np.int64)
    et = rng.randint(0, NUM_EDGE_TYPES, size=e).astype(np.int64)
    torch.manual_seed(5)
    mod = reference_saint_conv()(16, 8, NUM_NODE_TYPES, NUM_EDGE_TYPES, ALPHA, **kw)
    mod.train(train)
    with torch.no_grad():   # positive relation weights: the layer divides by the weighted in-degree without a clamp
        mod.relation_weight.copy_(torch.as_tensor(rng.uniform(0.3, 1.5, size=NUM_EDGE_TYPES) / ALPHA))
        mod.bias.copy_(torch.as_tensor(rng.randn(8) * 0.1))
    x = torch.as_tensor(rng.randn(n, 16)).requires_grad_(True)
    out, ew = mod((x, x) if tuple_input else x, torch.as_tensor(edge_index), torch.as_tensor(et), return_weights=True)
    gout = torch.as_tensor(np.random.RandomState(7).randn(*out.shape))
    out.backward(gout)
    blob = dict(edge_index=edge_index, edge_type=et, n_src=np.int64(n), n_dst=np.int64(n), x_src=x.detach().numpy(),
                gx_src=x.grad.numpy(), out=out.detach().numpy(), gout=gout.numpy(), ew=ew.detach().numpy(),
                meta=np.array(json.dumps(dict(kind='SaintREGCNConv', kw=kw, train=train, tuple_input=tuple_input, in_channels=16,
                                              out_channels=8, num_node_types=NUM_NODE_TYPES,
                                              num_edge_types=NUM_EDGE_TYPES))))
    for k, v in mod.state_dict().items():
        blob['param::' + k] = v.detach().numpy()
    for k, p in mod.named_parameters():
        blob['grad::' + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, 'mag', name + '.npz'), **blob)
    print('%-30s out %s' % (name, tuple(out.shape)))


def reference_regnn_model(model):
    """``REGNN`` of the mag/regnn_ns.py script, lifted out by source range (unmodified) and bound to the conv classes of
    mag/regnn_layers.py and to an ``args`` namespace standing in for the script's command line."""
    import ast
    import types
    import torch.nn.functional as F
    from torch.nn import Linear, ModuleDict, ModuleList, Parameter, ParameterDict
    path = os.path.join(REF, 'mag', 'regnn_ns.py')
    text = open(path).read()
    node = [n for n in ast.parse(text).body if isinstance(n, ast.ClassDef) and n.name == 'REGNN'][0]
    ns = dict(torch=torch, F=F, Linear=Linear, ModuleDict=ModuleDict, ModuleList=ModuleList, Parameter=Parameter,
              ParameterDict=ParameterDict, REGCNConv=RL.REGCNConv, REGATConv=RL.REGATConv, REGATv2Conv=RL.REGATv2Conv,
              args=types.SimpleNamespace(model=model, feats_type=3, self_loop_type=2, no_re=False))
    exec(compile(ast.get_source_segment(text, node), path, 'exec'), ns)
    return ns['REGNN']


def save_model_case(name, model, heads, residual):
    """Two sampled blocks in PyG's ``adjs`` form (outermost first; targets are the first rows of the sources)."""
    rng = np.random.RandomState(zlib.crc32(name.encode()) % (2 ** 31))
    n_total, sizes = 90, [(70, 30), (30, 10)]           # n_id has 70 entries; layer 1: 70 -> 30, layer 2: 30 -> 10
    feat_dims = {0: 12, 1: 7, 2: 9}
    node_type = rng.randint(0, NUM_NODE_TYPES, size=n_total)
    local_idx = np.zeros(n_total, dtype=np.int64)
    for t in range(NUM_NODE_TYPES):
        local_idx[node_type == t] = np.arange((node_type == t).sum())
    x_dict = {t: rng.randn(int((node_type == t).sum()), d) for t, d in feat_dims.items()}
    n_id = rng.permutation(n_total)[:sizes[0][0]]
    e_total = 400
    edge_type_all = rng.randint(0, NUM_EDGE_TYPES, size=e_total)
    adjs, off = [], 0
    for (n_src, n_dst), e in zip(sizes, (260, 140)):
        ei = np.stack([rng.randint(0, n_src, size=e), rng.randint(0, n_dst, size=e)]).astype(np.int64)
        adjs.append((ei, np.arange(off, off + e), (n_src, n_dst)))
        off += e
    torch.manual_seed(3)
    net = reference_regnn_model(model)(16, 8, 5, heads, 2, ALPHA, 0.0, feat_dims, NUM_EDGE_TYPES, residual, False)
    with torch.no_grad():
        for conv in net.convs:
            conv.relation_weight.copy_(torch.as_tensor(rng.uniform(0.5, 1.5, size=tuple(conv.relation_weight.shape)) / ALPHA))
            conv.bias.copy_(torch.as_tensor(rng.randn(*conv.bias.shape) * 0.1))
    net.eval()
    xd = {t: torch.as_tensor(v) for t, v in x_dict.items()}
    out = net(torch.as_tensor(n_id), xd, [(torch.as_tensor(a), torch.as_tensor(b), c) for a, b, c in adjs],
              torch.as_tensor(edge_type_all), torch.as_tensor(node_type), torch.as_tensor(local_idx))
    gout = torch.as_tensor(np.random.RandomState(7).randn(*out.shape))
    out.backward(gout)
    blob = dict(n_id=n_id, node_type=node_type, local_node_idx=local_idx, edge_type=edge_type_all, out=out.detach().numpy(),
                gout=gout.numpy(),
                meta=np.array(json.dumps(dict(kind='REGNN', model=model, heads=heads, residual=residual, in_channels=16,
                                              hidden_channels=8, out_channels=5, num_layers=2,
                                              feat_dims={str(k): v for k, v in feat_dims.items()},
                                              num_edge_types=NUM_EDGE_TYPES, sizes=sizes))))
    for t, v in x_dict.items():
        blob['x::%d' % t] = v
    for i, (ei, eid, _) in enumerate(adjs):
        blob['adj%d::edge_index' % i], blob['adj%d::e_id' % i] = ei, eid
    for k, v in net.state_dict().items():
        blob['param::' + k] = v.detach().numpy()
    for k, p in net.named_parameters():
        blob['grad::' + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    np.savez_compressed(os.path.join(HERE, 'mag', name + '.npz'), **blob)
    print('%-30s out %s' % (name, tuple(out.shape)))


def main():
    save_model_case('model_regnn_regcn', 'regcn', 1, True)
    save_model_case('model_regnn_regat_h2', 'regat', 2, True)
    save_model_case('model_regnn_regatv2_h2', 'regatv2', 2, False)
    save_saint_case('saint_regcn_tensor_input', 50, 320, 21, False)
    save_saint_case('saint_regcn_tuple_input', 36, 200, 22, True)
    save_saint_case('saint_regcn_softmax', 50, 320, 23, False, dict(use_softmax=True))
    save_saint_case('saint_regcn_dropout', 50, 320, 24, False, dict(dropout=0.4), train=True)
    save_saint_case('saint_regcn_softmax_dropout', 44, 260, 25, True, dict(use_softmax=True, dropout=0.25), train=True)
    b1, b2 = block(1), block(2, empty_targets=False)
    square = block(3, n_src=40, n_dst=40, e=260)       # full-graph inference shape: targets == sources
    for name, kw, blk in [
        ('regcn_loop2', dict(self_loop_type=2), b1),
        ('regcn_loop1_residual', dict(self_loop_type=1, residual=True), b1),     # targets without in-edges: mean of nothing
        ('regcn_loop3', dict(self_loop_type=3), b2),
        ('regcn_loop2_residual_square', dict(self_loop_type=2, residual=True), square),
    ]:
        save_case(name, 'REGCNConv', kw, blk)
    for kind, tag in [('REGATConv', 'regat'), ('REGATv2Conv', 'regatv2')]:
        for name, kw, blk, tup in [
            (tag + '_h2_loop2', dict(heads=2, self_loop_type=2), b1, True),
            (tag + '_h4_loop1_residual', dict(heads=4, self_loop_type=1, residual=True), b1, True),
            (tag + '_h2_mean', dict(heads=2, self_loop_type=2, concat=False), b2, True),
            (tag + '_h1_tensor_input', dict(heads=1, self_loop_type=2, negative_slope=0.05), square, False),
        ]:
            save_case(name, kind, kw, blk, tup)


if __name__ == '__main__':
    main()
