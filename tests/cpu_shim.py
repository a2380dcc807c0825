"""TEST-ONLY emulation of ``re_gnn_b200.ops`` (one function per CUDA operator, same signatures and
the same slot-order conventions) in plain PyTorch on the CPU, plus CPU versions of
``Graph.csr`` / ``Graph.etype_views`` built on oracle/csr_oracle.py.

Purpose: run the package's host logic -- autograd Functions, layer modules, the reference's model
files on top of them -- in the GPU-less build container.  Installed by the ``cpu_ops`` fixture via
monkeypatch; never imported by the package.
"""
import numpy as np
import torch

from oracle import csr_oracle

SLOPE = 0.01


def _leaky(x, s):
    return torch.where(x > 0, x, x * s)


def _lgrad(x, s):
    return torch.where(x > 0, torch.ones_like(x), torch.full_like(x, s))


def _rows_of(indptr):
    n = indptr.numel() - 1
    return torch.repeat_interleave(torch.arange(n), (indptr[1:] - indptr[:-1]).long())


def _w(theta, alpha):
    return _leaky(theta * alpha, SLOPE)


def graph_csr(self):
    if self._csr is None:
        c = csr_oracle.csr_build(self._src.numpy(), self._dst.numpy(), self._n)
        self._csr = {k: torch.as_tensor(v) for k, v in c.items()}
    return self._csr


def graph_etype_views(self, e_feat, num_relations):
    c = self.csr()
    et = e_feat.numpy()
    if et.size and (et.min() < 1 or et.max() > num_relations):
        raise RuntimeError('edge type outside [1, %d]' % num_relations)
    a, b = csr_oracle.etype_permute(et, c['eid'].numpy(), c['slot_t'].numpy())
    return torch.as_tensor(a), torch.as_tensor(b), None


def wdeg_norm_fwd(csr, et_csr, theta, alpha, exponent, rows=None, counts=None, clamp_min=1.0):
    w = _w(theta.detach().view(-1), alpha)
    n = csr['indptr'].numel() - 1
    deg = torch.zeros(n, dtype=w.dtype).index_add(0, csr['row'].long(), w[et_csr.long()])
    if clamp_min > 0:
        return deg, deg.clamp(min=clamp_min) ** exponent
    safe = torch.where(deg == 0, torch.ones_like(deg), deg)
    return deg, torch.where(deg == 0, torch.zeros_like(deg), safe ** exponent)


def wdeg_norm_bwd(csr, et_csr, theta, alpha, exponent, deg, d_norm, rows=None, counts=None, clamp_min=1.0):
    th = theta.detach().view(-1)
    if clamp_min > 0:
        dd = torch.where(deg >= clamp_min, exponent * deg.clamp(min=clamp_min) ** (exponent - 1) * d_norm,
                         torch.zeros_like(deg))
    else:
        safe = torch.where(deg == 0, torch.ones_like(deg), deg)
        dd = torch.where(deg != 0, exponent * safe ** (exponent - 1) * d_norm, torch.zeros_like(deg))
    dw = torch.zeros_like(th).index_add(0, et_csr.long(), dd[csr['row'].long()])
    return dw * alpha * _lgrad(th * alpha, SLOPE)


def spmm(indptr, indices, etype, theta, alpha, norm_src, norm_dst, x, rows=None, out=None, split=None, order=None):
    x = x.detach()
    row, col = _rows_of(indptr), indices.long()
    coef = torch.ones(col.numel(), dtype=x.dtype)
    if theta is not None:
        coef = _w(theta.detach().view(-1), alpha)[etype.long()]
    if norm_src is not None:
        coef = coef * norm_src.detach()[col]
    y = torch.zeros((indptr.numel() - 1, x.shape[1]), dtype=x.dtype).index_add(0, row, coef[:, None] * x[col])
    if norm_dst is not None:
        y = y * norm_dst.detach()[:, None]
    if rows is not None:          # only the owned row block is written (partitioned runs)
        keep = torch.zeros(y.shape[0], 1, dtype=y.dtype)
        keep[rows[0]:rows[1]] = 1
        y = y * keep
    if out is not None:
        if rows is None:
            out[:y.shape[0]].copy_(y)   # `out` may carry padding rows
        else:
            out[rows[0]:rows[1]] = y[rows[0]:rows[1]]
        return out
    return y


def spmm_bwd_w(csr, et_csr, theta, alpha, norm, x, y, g, dx, rows=None, sides=3, split=None):
    x, y, g, dx = x.detach(), y.detach(), g.detach(), dx.detach()
    d_norm = None
    if norm is None:
        sides = 0
    n = csr['indptr'].numel() - 1
    own = torch.zeros(n, dtype=x.dtype)
    own[slice(*rows) if rows is not None else slice(0, n)] = 1
    if norm is not None:
        d_norm = ((y * g).sum(1) * bool(sides & 2) + (x * dx).sum(1) * bool(sides & 1)) / norm.detach()
        d_norm = torch.where(own > 0, d_norm, torch.zeros_like(d_norm))   # foreign rows hold garbage
    d_theta = None
    if theta is not None:
        th = theta.detach().view(-1)
        row, col = csr['row'].long(), csr['indices'].long()
        dwe = (x[col] * g[row]).sum(1)
        if sides & 1:
            dwe = dwe * norm.detach()[col]
        if sides & 2:
            dwe = dwe * norm.detach()[row]
        dw = torch.zeros_like(th).index_add(0, et_csr.long(), torch.where(own[row] > 0, dwe, torch.zeros_like(dwe)))
        d_theta = dw * alpha * _lgrad(th * alpha, SLOPE)
    return d_theta, d_norm


def spmm_bwd_fused(csr, et_t, theta, alpha, norm, x, g, rows=None, sides=3, out=None, want_xdx=False, y=None,
                   want_dnorm=False):
    ns = norm if (norm is not None and sides & 2) else None
    nd = norm if (norm is not None and sides & 1) else None
    dx = spmm(csr['indptr_t'], csr['indices_t'], et_t, theta, alpha, ns, nd, g, rows=rows, out=out)
    x, g = x.detach(), g.detach()
    th = theta.detach().view(-1)
    src = _rows_of(csr['indptr_t'])
    dstn = csr['indices_t'].long()
    dwe = (x[src] * g[dstn]).sum(1)
    if norm is not None and sides & 1:
        dwe = dwe * norm.detach()[src]
    if norm is not None and sides & 2:
        dwe = dwe * norm.detach()[dstn]
    if rows is not None:          # a rank owns the edges whose SOURCE row it owns
        dwe = torch.where((src >= rows[0]) & (src < rows[1]), dwe, torch.zeros_like(dwe))
    dw = torch.zeros_like(th).index_add(0, et_t.long(), dwe)
    xdx = None
    if want_xdx:
        xdx = (x * dx.detach()).sum(1)
        if rows is not None:
            own = torch.zeros(xdx.numel(), dtype=torch.bool)
            own[rows[0]:rows[1]] = True
            xdx = torch.where(own, xdx, torch.zeros_like(xdx))
    if want_dnorm:               # the folded norm gradient of the lane-group kernel
        n = norm.numel()
        d = (x[:n] * dx.detach()[:n]).sum(1) * bool(sides & 1)
        if sides & 2:
            d = d + (y.detach()[:n] * g[:n]).sum(1)
        xdx = d / norm.detach()
    return dx, dw * alpha * _lgrad(th * alpha, SLOPE), xdx


def rowdot_norm_bwd(norm, x, y, g, dx, rows=None, sides=3, xdx=None):
    n = norm.numel()  # row-indexed buffers may carry padding rows past n (even all-gather layout)
    x, y, g, dx = x.detach()[:n], y.detach()[:n], g.detach()[:n], dx.detach()[:n]
    xd = xdx[:n] if xdx is not None else (x * dx).sum(1)
    d = ((y * g).sum(1) * bool(sides & 2) + xd * bool(sides & 1)) / norm.detach()
    if rows is not None:
        own = torch.zeros(n, dtype=torch.bool)
        own[rows[0]:rows[1]] = True
        d = torch.where(own, d, torch.zeros_like(d))
    return d


def _softmax_rows(l, row, n):
    mx = torch.full((n, l.shape[1]), float('-inf'), dtype=l.dtype).scatter_reduce(
        0, row[:, None].expand_as(l), l, 'amax', include_self=True)
    ex = torch.exp(l - mx[row])
    sm = torch.zeros((n, l.shape[1]), dtype=l.dtype).index_add(0, row, ex)
    mx = torch.where(torch.isinf(mx), torch.zeros_like(mx), mx)
    return ex / sm[row], mx, sm


def _keep_csr(keep, csr):
    return None if keep is None else keep.detach()[csr['eid'].long()]


def _to_edge_order(v_csr, csr):
    out = torch.empty_like(v_csr)
    out[csr['eid'].long()] = v_csr
    return out


def _gat_logits(csr, et_csr, theta, alpha, el, er, slope):
    row, col = csr['row'].long(), csr['indices'].long()
    pre = el.detach()[col] + er.detach()[row]
    if theta is not None and et_csr is not None:
        pre = pre + _w(theta.detach(), alpha)[et_csr.long()]
    return pre, _leaky(pre, slope)


def gat_fwd(csr, et_csr, theta, alpha, feat, el, er, slope, keep=None, want_attn=False, rows=None):
    feat = feat.detach()
    n = feat.shape[0]
    row, col = csr['row'].long(), csr['indices'].long()
    _, l = _gat_logits(csr, et_csr, theta, alpha, el, er, slope)
    a, mx, sm = _softmax_rows(l, row, n)
    kc = _keep_csr(keep, csr)
    at = a if kc is None else a * kc
    out = torch.zeros_like(feat).index_add(0, row, at[:, :, None] * feat[col])
    return out, mx, sm, (_to_edge_order(at, csr) if want_attn else None)


def gat_bwd(csr, et_csr, et_t, theta, alpha, feat, el, er, slope, keep, out, rowmax, rowsum, g, attn_l=None,
            attn_r=None, rows=None):
    feat, g, out = feat.detach(), g.detach(), out.detach()
    row, col = csr['row'].long(), csr['indices'].long()
    pre, l = _gat_logits(csr, et_csr, theta, alpha, el, er, slope)
    sm = rowsum[row]
    a = torch.exp(l - rowmax[row]) * torch.where(sm > 0, 1.0 / sm, torch.zeros_like(sm))
    kc = _keep_csr(keep, csr)
    at = a if kc is None else a * kc
    da = (feat[col] * g[row]).sum(-1)
    S = (out * g).sum(-1)
    dpre = (at * da - a * S[row]) * _lgrad(pre, slope)
    d_er = torch.zeros_like(rowmax).index_add(0, row, dpre)
    d_el = torch.zeros_like(rowmax).index_add(0, col, dpre)
    d_feat = torch.zeros_like(g).index_add(0, col, at[:, :, None] * g[row])
    d_theta = None
    if theta is not None and et_csr is not None:
        th = theta.detach()
        dw = torch.zeros_like(th).index_add(0, et_csr.long(), dpre)
        d_theta = dw * alpha * _lgrad(th * alpha, SLOPE)
    d_al = d_ar = None
    if attn_l is not None:      # gradients through el = <feat, attn_l>, er = <feat, attn_r>
        h, d = g.shape[1], g.shape[2]
        d_feat = d_feat + d_el[:, :, None] * attn_l.detach().view(1, h, d) + d_er[:, :, None] * attn_r.detach().view(1, h, d)
        d_al = (d_el[:, :, None] * feat).sum(0).reshape(-1)
        d_ar = (d_er[:, :, None] * feat).sum(0).reshape(-1)
    return d_feat, d_el, d_er, d_theta, d_al, d_ar


def attn_scores_fwd(feat, attn_l, attn_r):
    f = feat.detach()
    h, d = f.shape[1], f.shape[2]
    return (f * attn_l.detach().view(1, h, d)).sum(-1), (f * attn_r.detach().view(1, h, d)).sum(-1)




def _v2_logits(csr, et_csr, theta, alpha, fs, fd, attn, slope):
    row, col = csr['row'].long(), csr['indices'].long()
    q = fs.detach()[col] + fd.detach()[row]
    lr = _leaky(q, slope)
    l = (lr * attn.detach().view(1, fs.shape[1], fs.shape[2])).sum(-1)
    if theta is not None and et_csr is not None:
        l = l + _w(theta.detach(), alpha)[et_csr.long()]
    return q, lr, l


def gatv2_fwd(csr, et_csr, theta, alpha, fs, fd, attn, slope, keep=None, want_attn=False, rows=None, save=False):
    n = fd.shape[0]
    row, col = csr['row'].long(), csr['indices'].long()
    q, _, l = _v2_logits(csr, et_csr, theta, alpha, fs, fd, attn, slope)
    a, mx, sm = _softmax_rows(l, row, n)
    kc = _keep_csr(keep, csr)
    at = a if kc is None else a * kc
    out = torch.zeros_like(fd.detach()).index_add(0, row, at[:, :, None] * fs.detach()[col])
    saved = (l, q > 0) if save else None    # the sign mask as a bool [E,H,D] tensor in slot order
    return out, mx, sm, (_to_edge_order(at, csr) if want_attn else None), saved


def gatv2_bwd(csr, et_csr, theta, alpha, fs, fd, attn, slope, keep, out, rowmax, rowsum, saved, g, rows=None):
    """The formulas of the kernels (stored logits + sign bits; d_attn as the two node-side sums)."""
    g, out, fs, fd = g.detach(), out.detach(), fs.detach(), fd.detach()
    l, pos = saved
    row, col = csr['row'].long(), csr['indices'].long()
    a = torch.exp(l - rowmax[row]) / rowsum[row]
    kc = _keep_csr(keep, csr)
    at = a if kc is None else a * kc
    da = (fs[col] * g[row]).sum(-1)
    S = (out * g).sum(-1)
    dl = at * da - a * S[row]
    av = attn.detach().view(1, fs.shape[1], fs.shape[2])
    phi = torch.where(pos, torch.ones((), dtype=g.dtype), torch.full((), slope, dtype=g.dtype))
    t = dl[:, :, None] * phi
    t_dst = torch.zeros_like(fd).index_add(0, row, t)
    t_src = torch.zeros_like(fs).index_add(0, col, t)
    d_fd = av * t_dst
    d_fs = torch.zeros_like(fs).index_add(0, col, at[:, :, None] * g[row]) + av * t_src
    d_attn = ((fs * t_src).sum(0) + (fd * t_dst).sum(0)).reshape(-1)
    d_theta = None
    if theta is not None and et_csr is not None:
        th = theta.detach()
        d_theta = torch.zeros_like(th).index_add(0, et_csr.long(), dl) * alpha * _lgrad(th * alpha, SLOPE)
    return d_fs, d_fd, d_attn, d_theta


def _grouped_rows(tables, seg_ptr, perm, local_idx, num_rows):
    """(type of every output row, row of its table) from the sorted-segment description of the C ABI."""
    seg = seg_ptr.long()
    pos_type = torch.repeat_interleave(torch.arange(len(tables)), seg[1:] - seg[:-1])
    rows = perm.long() if perm is not None else torch.arange(num_rows)
    typ = torch.empty(num_rows, dtype=torch.long)
    typ[rows] = pos_type
    if local_idx is not None:
        src = local_idx.long()
    else:
        src = torch.empty(num_rows, dtype=torch.long)
        src[rows] = torch.arange(num_rows) - seg[pos_type]
    return typ, src


def grouped_linear_fwd(tables, weights, biases, seg_ptr, perm, local_idx, num_rows):
    typ, src = _grouped_rows(tables, seg_ptr, perm, local_idx, num_rows)
    out = torch.zeros((num_rows, weights[0].shape[0]), dtype=tables[0].dtype)
    for t, (x, w, b) in enumerate(zip(tables, weights, biases)):
        m = typ == t
        y = x.detach()[src[m]] @ w.detach().t()
        out[m] = y + b.detach() if b is not None else y
    return out


def grouped_linear_bwd(tables, n_out, seg_ptr, perm, local_idx, dout, want_bias):
    typ, src = _grouped_rows(tables, seg_ptr, perm, local_idx, dout.shape[0])
    dws, dbs = [], []
    for t, x in enumerate(tables):
        m = typ == t
        dws.append(dout[m].t() @ x.detach()[src[m]])
        dbs.append(dout[m].sum(0))
    return dws, (dbs if want_bias else None)


def sampler_sample_slots(self, targets, fanout, key):
    from oracle import sampler_oracle
    return torch.as_tensor(sampler_oracle.sample_slots(self.graph.csr()['indptr'].numpy(), targets.numpy(), fanout, key))


def saint_walks(self, key):
    from oracle import sampler_oracle
    csr = {k: v.numpy() for k, v in self.graph.csr().items() if torch.is_tensor(v)}
    return torch.as_tensor(sampler_oracle.random_walks(csr, self.graph.number_of_nodes(), self.roots, self.walk_length, key))


def install(monkeypatch):
    from re_gnn_b200 import graph as G, ops, sampling
    monkeypatch.setattr(sampling.NeighborSampler, 'sample_slots', sampler_sample_slots)
    monkeypatch.setattr(sampling.SaintRandomWalkSampler, 'walks', saint_walks)
    monkeypatch.setattr(G.Graph, 'csr', graph_csr)
    monkeypatch.setattr(G.Graph, 'etype_views', graph_etype_views)
    for name in ('wdeg_norm_fwd', 'wdeg_norm_bwd', 'spmm', 'spmm_bwd_w', 'spmm_bwd_fused', 'rowdot_norm_bwd', 'gat_fwd', 'gat_bwd',
                 'gatv2_fwd', 'gatv2_bwd', 'attn_scores_fwd', 'grouped_linear_fwd', 'grouped_linear_bwd'):
        monkeypatch.setattr(ops, name, globals()[name])
