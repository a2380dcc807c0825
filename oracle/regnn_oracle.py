"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch (CPU, fp32 or fp64) restatement of the arithmetic of RE-GNN's relation-embedded
message-passing layers, written against explicit ``(src, dst, etype, num_nodes)`` arrays in
edge-id order instead of a DGLGraph.  Gradients come from autograd.

Who may import this: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- as the checker or the timed CPU baseline, never as a
fallback for ``re_gnn_b200`` (the product path raises when its CUDA library is missing).

Parity status: the reference ships no tests, golden vectors or fixtures, and its arithmetic lives in
``dgl==0.7.1`` which cannot be installed here.  What pins this oracle instead:
  * ``tests/golden/*.npz`` were produced by running the reference's own, unmodified
    ``layer/*.py`` / ``model/*.py`` over ``oracle/dgl_stub`` (a restatement of the DGL *primitives*
    only: SURVEY.md Appendix B) with ``tests/golden/make_golden.py``;
    ``tests/test_oracle_golden.py`` checks every function below against them;
  * ``tests/test_oracle_dense.py`` checks it against an independent dense-adjacency formulation
    in float64 and against analytic known answers.
  * the MAG-stack functions (``mag_*``) are pinned the same way through ``oracle/pyg_stub`` and
    ``tests/golden/mag/*.npz`` (see the comment above them).
The DGL primitive semantics themselves (u_mul_e/copy_u/u_add_v + sum, edge_softmax) remain
"parity unpinned": no DGL binary is available to confirm them.

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
import torch
import torch.nn.functional as F

RELATION_SLOPE = 0.01  # nn.LeakyReLU() default used for the relation table in every layer


def relation_weight(edge_weight, alpha):
    """``w = LeakyReLU_0.01(edge_weight * alpha)``  -- layer/REGraphConv.py:58-60,
    layer/REGATConv.py:72-74, layer/REGATv2Conv.py:140-142, layer/REMixHopConv.py:50-52."""
    return F.leaky_relu(edge_weight * alpha, RELATION_SLOPE)


def edge_relation(edge_weight, alpha, etype):
    """``ew = w[e_feat - 1]`` (1-based edge types) -- layer/REGraphConv.py:61."""
    return relation_weight(edge_weight, alpha)[etype - 1]


def segment_sum(values, dst, num_nodes):
    """``update_all(..., fn.sum)``: sum of per-edge messages over the in-edges of each node."""
    out = torch.zeros((num_nodes,) + tuple(values.shape[1:]), dtype=values.dtype)
    return out.index_add(0, dst, values)


def weighted_degree_norm(ew, dst, num_nodes, exponent=-0.5):
    """``norm = clamp(sum_in ew, min=1) ** exponent`` -- layer/REGraphConv.py:66-73,
    layer/REMixHopConv.py:58-62 (exponent -0.5); layer/RESAGEConv.py:78 uses -1."""
    deg = segment_sum(ew.reshape(-1, 1), dst, num_nodes).squeeze(1)
    return torch.pow(deg.clamp(min=1), exponent)


def edge_softmax(logits, dst, num_nodes):
    """dgl edge_softmax (norm_by='dst'): per destination, per trailing index; the per-destination
    maximum is subtracted; no epsilon -- call sites layer/REGATConv.py:88, layer/REGATv2Conv.py:148."""
    shape = (num_nodes,) + tuple(logits.shape[1:])
    idx = dst.view((-1,) + (1,) * (logits.dim() - 1)).expand_as(logits)
    mx = torch.full(shape, float('-inf'), dtype=logits.dtype).scatter_reduce(
        0, idx, logits.detach(), 'amax', include_self=True)
    ex = torch.exp(logits - mx[dst])
    return ex / segment_sum(ex, dst, num_nodes)[dst]


# ------------------------------------------------------------------------------------------------
def regraphconv_forward(src, dst, etype, num_nodes, feat, edge_weight, alpha, weight=None, bias=None,
                        activation=None, norm=True, in_feats=None, out_feats=None):
    """layer/REGraphConv.py:52-106 with dropout 0.  ``weight`` is ``[in, out]`` or None."""
    ew = edge_relation(edge_weight, alpha, etype)                       # :58-62, [E,1]
    if norm:
        nrm = weighted_degree_norm(ew, dst, num_nodes).unsqueeze(1)     # :66-75
        feat = feat * nrm                                               # :76
    in_feats = feat.shape[1] if in_feats is None else in_feats
    out_feats = (weight.shape[1] if weight is not None else in_feats) if out_feats is None else out_feats
    if in_feats > out_feats:                                            # :78-87
        if weight is not None:
            feat = feat @ weight
        rst = segment_sum(feat[src] * ew, dst, num_nodes)
    else:                                                               # :88-95
        rst = segment_sum(feat[src] * ew, dst, num_nodes)
        if weight is not None:
            rst = rst @ weight
    if norm:
        rst = rst * nrm                                                 # :97-98
    if bias is not None:
        rst = rst + bias                                                # :100-101
    if activation is not None:
        rst = activation(rst)                                           # :103-104
    return rst


def regat_forward(src, dst, etype, num_nodes, feat, attn_l, attn_r, edge_weight, alpha,
                  negative_slope=0.2, fc_weight=None, res_weight=None, residual_identity=False,
                  activation=None, return_attention=False):
    """layer/REGATConv.py:64-100 with dropout 0.  ``fc_weight`` is nn.Linear's ``[H*D, in]`` or None
    (``use_weight=False`` -> Identity).  ``etype=None`` drops the relation term (:71,:83)."""
    H, D = attn_l.shape[1], attn_l.shape[2]
    h = feat
    f = (h @ fc_weight.t() if fc_weight is not None else h).view(-1, H, D)   # :67
    el = (f * attn_l).sum(-1, keepdim=True)                                  # :68
    er = (f * attn_r).sum(-1, keepdim=True)                                  # :69
    e = el[src] + er[dst]                                                    # :80
    if etype is not None:
        e = e + edge_relation(edge_weight, alpha, etype).reshape(-1, H, 1)   # :72-75,:84
    e = F.leaky_relu(e, negative_slope)                                      # :86
    a = edge_softmax(e, dst, num_nodes)                                      # :88
    rst = segment_sum(f[src] * a, dst, num_nodes)                            # :90-92
    if res_weight is not None:                                               # :94-96
        rst = rst + (h @ res_weight.t()).view(h.shape[0], -1, D)
    elif residual_identity:
        rst = rst + h.view(h.shape[0], -1, D)
    if activation is not None:
        rst = activation(rst)
    return (rst, a) if return_attention else rst


def regatv2_forward(src, dst, etype, num_nodes, feat, attn, edge_weight, alpha, negative_slope=0.2,
                    fc_src=None, fc_dst=None, res_fc=None, residual_identity=False,
                    activation=None, return_attention=False):
    """layer/REGATv2Conv.py:103-164 with dropout 0.  ``fc_src`` / ``fc_dst`` / ``res_fc`` are
    ``(weight[H*D, in], bias[H*D] | None)`` tuples or None (Identity); pass ``fc_dst=fc_src`` for
    ``share_weights``."""
    H, D = attn.shape[1], attn.shape[2]

    def lin(p, x):
        if p is None:
            return x
        y = x @ p[0].t()
        return y + p[1] if p[1] is not None else y

    h = feat
    fs = lin(fc_src, h).view(-1, H, D)                                       # :123-124
    fd = lin(fc_dst, h).view(-1, H, D)                                       # :125-130
    e = F.leaky_relu(fs[src] + fd[dst], negative_slope)                      # :135-136
    e = (e * attn).sum(-1).unsqueeze(2)                                      # :137
    if etype is not None:
        e = e + edge_relation(edge_weight, alpha, etype).reshape(-1, H, 1)   # :139-145
    a = edge_softmax(e, dst, num_nodes)                                      # :148
    rst = segment_sum(fs[src] * a, dst, num_nodes)                           # :150-152
    if res_fc is not None:                                                   # :154-156
        rst = rst + lin(res_fc, h).view(h.shape[0], -1, D)
    elif residual_identity:
        rst = rst + h.view(h.shape[0], -1, D)
    if activation is not None:
        rst = activation(rst)
    return (rst, a) if return_attention else rst


def remixhop_forward(src, dst, etype, num_nodes, feats, edge_weight, alpha, weights, p=(0, 1, 2),
                     activation=None, bn=None):
    """layer/REMixHopConv.py:48-94 with dropout 0.  ``weights`` maps power j -> nn.Linear weight
    ``[out, in]``.  The propagation after the highest power is dead code in the reference
    (:70-82) and is skipped here; ``bn`` is an optional callable (BatchNorm1d)."""
    ew = edge_relation(edge_weight, alpha, etype)                            # :50-55
    nrm = weighted_degree_norm(ew, dst, num_nodes).unsqueeze(1)              # :58-64
    outputs = []
    for j in range(max(p) + 1):                                              # :70
        if j in p:
            outputs.append(feats @ weights[j].t())                          # :74-76
        if j < max(p):
            feats = feats * nrm                                              # :78
            feats = segment_sum(feats[src], dst, num_nodes)                  # :79-81 (copy_u: unweighted)
            feats = feats * nrm                                              # :82
    final = torch.cat(outputs, dim=1)                                        # :84
    if bn is not None:
        final = bn(final)
    if activation is not None:
        final = activation(final)
    return final


def resage_forward(src, dst, etype, num_nodes, feat, edge_weight, alpha, weight=None, bias=None,
                   activation=None, norm=True):
    """layer/RESAGEConv.py:55-114 (dropout 0; ``in_feats <= out_feats`` branch order is irrelevant
    to the maths): source-side norm with exponent -1, no destination-side norm, ``+ feat_root``."""
    feat_root = feat @ weight if weight is not None else feat               # :60-63 (weight_root unused, Q9)
    ew = edge_relation(edge_weight, alpha, etype)
    if norm:
        feat = feat * weighted_degree_norm(ew, dst, num_nodes, -1.0).unsqueeze(1)   # :72-82
    rst = segment_sum(feat[src] * ew, dst, num_nodes)
    if weight is not None:
        rst = rst @ weight
    rst = rst + feat_root                                                    # :106
    if bias is not None:
        rst = rst + bias
    if activation is not None:
        rst = activation(rst)
    return rst


# ------------------------------------------------------------------------------------------------
# Callers (model/*.py) restated on top of the layer functions: used for the epoch-time CPU baseline.
def regcn_model_forward(src, dst, etype, num_nodes, features_list, params, alpha, n_layers,
                        activation=F.elu):
    """model/REGCN.py:35-46 with dropout 0.  ``params``: dict with ``fc`` (list of (W,b)),
    ``layers`` (list of dicts edge_weight/weight/bias), ``out`` (W,b)."""
    h = torch.cat([x @ w.t() + b for (w, b), x in zip(params['fc'], features_list)], 0)
    for l in range(n_layers):
        lp = params['layers'][l]
        act = activation if 0 < l < n_layers - 1 else None
        h = regraphconv_forward(src, dst, etype, num_nodes, h, lp['edge_weight'], alpha,
                                lp.get('weight'), lp.get('bias'), act)
    w, b = params['out']
    return h @ w.t() + b, h


# ------------------------------------------------------------------------------------------------
# MAG stack (PyG API).  PyG / torch_scatter cannot be installed here; these restatements of mag/regnn_layers.py are
# pinned to fixtures recorded from the reference's own, unmodified mag/regnn_layers.py + mag/utils.py run over
# oracle/pyg_stub (tests/golden/make_golden_mag.py -> tests/golden/mag/*.npz, checked by tests/test_oracle_golden.py).
# The PyG MessagePassing / torch_scatter primitive semantics themselves remain "parity unpinned" (no PyG binary).
# mag/regnn_saint.py is a script (importing it parses arguments and loads the dataset): make_golden_mag.py lifts its
# REGCNConv class out by source range, unmodified, and saint_regcn_forward below is pinned to that class' outputs too.
def mag_regcn_forward(x_src, x_target, edge_index, edge_type, target_node_type, weight, bias, relation_weight,
                      scaling_factor, num_edge_types, self_loop_type=2, residual=False):
    """mag/regnn_layers.py:80-150 (``REGCNConv.forward``; ``aggr='mean'``, bias added in ``update``)."""
    src, dst = edge_index[0], edge_index[1]
    n_dst = x_target.shape[0]
    if self_loop_type == 2:                                                       # :90-96
        loop = torch.arange(n_dst, dtype=src.dtype)
        src, dst = torch.cat([src, loop]), torch.cat([dst, loop])
        edge_type = torch.cat([edge_type, target_node_type + num_edge_types])
    xs = x_src @ weight                                                           # :102
    w = F.leaky_relu(relation_weight * scaling_factor, RELATION_SLOPE)[edge_type]  # :110-113
    msg = w.view(-1, 1) * xs[src]                                                 # message(): ew * x_j, ew = edge_weight (:129)
    out = torch.zeros((n_dst, xs.shape[1]), dtype=xs.dtype).index_add(0, dst, msg)
    cnt = torch.zeros(n_dst, dtype=xs.dtype).index_add(0, dst, torch.ones(dst.numel(), dtype=xs.dtype))
    out = out / cnt.clamp(min=1).view(-1, 1) + bias                               # aggr='mean', update(): + bias
    if residual:
        out = out + x_target @ weight                                             # weight_root aliases weight (:50)
    return out


def _mag_self_loops(edge_index, edge_type, target_node_type, num_edge_types, self_loop_type, n_dst):
    """mag/regnn_layers.py:90-96 / :237-243 / :365-371: self_loop_type 2 appends one loop per target whose type
    is num_edge_types + node type; types 1 and 3 leave the edge list alone."""
    src, dst = edge_index[0], edge_index[1]
    if self_loop_type == 2:
        loop = torch.arange(n_dst, dtype=src.dtype)
        src, dst = torch.cat([src, loop]), torch.cat([dst, loop])
        edge_type = torch.cat([edge_type, target_node_type + num_edge_types])
    return src, dst, edge_type


def _mag_softmax(logits, dst, n_dst):
    """mag/utils.py:28-57: softmax over the in-edges of every target, stabilised with the GLOBAL maximum (all edges,
    all heads) and with 1e-16 added to the denominator."""
    out = (logits - logits.max()).exp()
    den = torch.zeros((n_dst,) + tuple(out.shape[1:]), dtype=out.dtype).index_add(0, dst, out)
    return out / (den[dst] + 1e-16)


def _mag_attention_tail(out, x_dst, bias, heads, out_channels, concat, residual):
    """mag/regnn_layers.py:275-284 / :413-422 (use_norm None)."""
    out = out.reshape(-1, heads * out_channels) if concat else out.mean(dim=1)
    out = out + bias
    if residual:
        out = out + x_dst.reshape(-1, heads * out_channels)
    return out


def mag_regat_forward(x_src, x_target, edge_index, edge_type, target_node_type, lin_weight, att_src, att_dst, bias,
                      relation_weight, scaling_factor, num_edge_types, heads, out_channels, negative_slope=0.2,
                      self_loop_type=2, residual=False, concat=True):
    """mag/regnn_layers.py:221-296 (``REGATConv.forward``; ``lin_dst`` aliases ``lin_src``, :188; aggr='add')."""
    h, c = heads, out_channels
    xs = (x_src @ lin_weight.t()).view(-1, h, c)                                   # :226-234
    xd = (x_target @ lin_weight.t()).view(-1, h, c)
    src, dst, edge_type = _mag_self_loops(edge_index, edge_type, target_node_type, num_edge_types, self_loop_type,
                                          xd.shape[0])
    a_src = (xs * att_src).sum(-1)                                                 # :254-255
    a_dst = (xd * att_dst).sum(-1)
    w = F.leaky_relu(relation_weight * scaling_factor, RELATION_SLOPE)[edge_type]  # :258-261, [E,H]
    logits = F.leaky_relu(w + a_src[src] + a_dst[dst], negative_slope)             # :263-267
    ew = _mag_softmax(logits, dst, xd.shape[0])                                    # :269
    out = torch.zeros_like(xd).index_add(0, dst, ew.unsqueeze(-1) * xs[src])       # :273, message :294-296
    return _mag_attention_tail(out, xd, bias, h, c, concat, residual)


def mag_regatv2_forward(x_src, x_target, edge_index, edge_type, target_node_type, lin_weight, att, bias,
                        relation_weight, scaling_factor, num_edge_types, heads, out_channels, negative_slope=0.2,
                        self_loop_type=2, residual=False, concat=True):
    """mag/regnn_layers.py:354-433 (``REGATv2Conv.forward``): the relation term is added AFTER the attention
    product and there is no outer LeakyReLU."""
    h, c = heads, out_channels
    xs = (x_src @ lin_weight.t()).view(-1, h, c)
    xd = (x_target @ lin_weight.t()).view(-1, h, c)
    src, dst, edge_type = _mag_self_loops(edge_index, edge_type, target_node_type, num_edge_types, self_loop_type,
                                          xd.shape[0])
    alpha = (F.leaky_relu(xs[src] + xd[dst], negative_slope) * att).sum(-1)        # :388-393
    w = F.leaky_relu(relation_weight * scaling_factor, RELATION_SLOPE)[edge_type]  # :396-399
    ew = _mag_softmax(w + alpha, dst, xd.shape[0])                                 # :401-403
    out = torch.zeros_like(xd).index_add(0, dst, ew.unsqueeze(-1) * xs[src])
    return _mag_attention_tail(out, xd, bias, h, c, concat, residual)


def mag_softmax(src, index, num_nodes):
    """mag/utils.py:28-57 ``softmax``: the GLOBAL maximum is subtracted and 1e-16 added to every denominator."""
    out = torch.exp(src - src.max()) if src.numel() else src
    den = torch.zeros((num_nodes,) + tuple(out.shape[1:]), dtype=out.dtype).index_add(0, index, out)
    return out / (den[index] + 1e-16)


def saint_regcn_forward(x, edge_index, edge_type, weight, bias, relation_weight, scaling_factor, use_softmax=False,
                        edge_keep=None, dropout=0.0, return_weights=False):
    """mag/regnn_saint.py:224-275 (``REGCNConv.forward``, ``aggr='add'``).  ``edge_keep`` (bool [E]) with ``dropout`` p
    restates ``F.dropout(ew, p, training=True)`` (:258) for a given mask: kept weights are scaled by 1/(1-p)."""
    src, dst = edge_index[0], edge_index[1]
    n = x.shape[0]
    xs = x @ weight                                                               # :234-236
    w = F.leaky_relu(relation_weight * scaling_factor, RELATION_SLOPE)[edge_type]  # :239-242
    if use_softmax:
        ew = mag_softmax(w, dst, n)                                               # :249-250
    else:
        deg = torch.zeros(n, dtype=x.dtype).index_add(0, dst, w)                  # :253 weighted_degree
        ew = w * deg.pow(-1.0)[dst]                                               # :254-256
    if edge_keep is not None:
        ew = ew * edge_keep.to(ew.dtype) / (1.0 - dropout)                        # :258
    out = torch.zeros((n, xs.shape[1]), dtype=x.dtype).index_add(0, dst, ew.view(-1, 1) * xs[src]) + bias
    return (out, ew) if return_weights else out