"""Raw operator wrappers: one Python function per C-ABI entry point of ``libregnn_b200.so``.

Each wrapper allocates its outputs with the PyTorch caching allocator, passes raw device pointers
and the current CUDA stream, and raises on any error.  No arithmetic happens here and there is no
CPU branch: tensors must live on a CUDA device.
"""
import ctypes
import os

import torch

from . import _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('re_gnn_b200 operators need CUDA tensors (got %s); there is no CPU path' % t.device)
    return t.detach().to(torch.float32).contiguous()


def _rows(rows, n):
    return (0, n) if rows is None else (int(rows[0]), int(rows[1]))


def wdeg_norm_fwd(csr, et_csr, theta, alpha, exponent, rows=None, counts=None, clamp_min=1.0):
    """deg[v] = sum_in w[etype], norm = max(deg,1)^exponent -> (deg[N], norm[N]).
    ``counts``: the per-(row, relation) in-edge counts of this e_feat (Graph.etype_views()[2])."""
    theta = _f32(theta).view(-1)
    n = csr['indptr'].numel() - 1
    rb, re = _rows(rows, n)
    deg = torch.empty(n, dtype=torch.float32, device=theta.device)
    norm = torch.empty(n, dtype=torch.float32, device=theta.device)
    with torch.cuda.device(theta.device):
        _lib.call('regnn_wdeg_norm_fwd', _ptr(csr['indptr']), _ptr(et_csr), _ptr(counts), _ptr(theta), float(alpha),
                  theta.numel(), float(exponent), float(clamp_min), rb, re, _ptr(deg), _ptr(norm), _stream())
        _lib.count_launches(1)
    return deg, norm


def wdeg_norm_bwd(csr, et_csr, theta, alpha, exponent, deg, d_norm, rows=None, counts=None, clamp_min=1.0):
    theta = _f32(theta).view(-1)
    d_norm = _f32(d_norm)
    n = csr['indptr'].numel() - 1
    rb, re = _rows(rows, n)
    r = theta.numel()
    partials = torch.empty(_lib.partial_blocks(re - rb) * r, dtype=torch.float64, device=theta.device)
    d_theta = torch.empty(r, dtype=torch.float32, device=theta.device)
    with torch.cuda.device(theta.device):
        _lib.call('regnn_wdeg_norm_bwd', _ptr(csr['indptr']), _ptr(et_csr), _ptr(counts), _ptr(theta), float(alpha), r,
                  float(exponent), float(clamp_min), rb, re, _ptr(deg), _ptr(d_norm), _ptr(partials), _ptr(d_theta),
                  _stream())
        _lib.count_launches(2)
    return d_theta


def _split_args(split, feat, device):
    if split is None:
        return None, None, 0
    ws = torch.empty(split['struct'].num_frags * feat, dtype=torch.float32, device=device)
    return ctypes.byref(split['struct']), ws, 1


# widths up to this run the lane-group kernel when the degree-sorted row order is supplied
NARROW_FEAT = int(os.environ.get('REGNN_NARROW_FEAT', '128'))


def row_order(csr, transposed=False):
    """Rows of one CSR view that are not cut into fragments, by descending slot count (stable): the work list of
    the narrow-row SpMM kernels (``row_order`` of regnn_spmm_fwd / regnn_spmm_bwd_fused).  Index bookkeeping with
    torch ops on the device, built on first use and cached in the csr dict."""
    key = 'order_t' if transposed else 'order'
    if key not in csr:
        indptr = csr['indptr_t' if transposed else 'indptr']
        split = csr.get('split_t' if transposed else 'split')
        deg = indptr[1:] - indptr[:-1]
        order = torch.sort(deg, descending=True, stable=True).indices
        # long rows (more slots than the split threshold) sort first: drop them
        csr[key] = order[(split['struct'].num_long if split is not None else 0):].to(torch.int32).contiguous()
    return csr[key]


def spmm(indptr, indices, etype, theta, alpha, norm_src, norm_dst, x, rows=None, out=None, split=None, order=None):
    """Y[v] = norm_dst[v] * sum_s w[etype[s]] * norm_src[indices[s]] * X[indices[s]] over rows.
    ``split``: the long-row decomposition of this CSR view (Graph.csr()['split' | 'split_t']);
    ``order``: ``row_order`` of the same view (used for widths <= NARROW_FEAT on the full row range)."""
    x = _f32(x)
    n = indptr.numel() - 1
    rb, re = _rows(rows, n)
    f = x.shape[1]
    theta = _f32(theta).view(-1) if theta is not None else None
    if out is None:
        out = torch.empty((n, f), dtype=torch.float32, device=x.device)
        if (rb, re) != (0, n):
            out.zero_()
    sp, ws, extra = _split_args(split, f, x.device)
    with torch.cuda.device(x.device):
        _lib.call('regnn_spmm_fwd', _ptr(indptr), _ptr(indices), _ptr(etype) if theta is not None else None,
                  _ptr(theta), float(alpha), theta.numel() if theta is not None else 0, _ptr(norm_src),
                  _ptr(norm_dst), _ptr(x), x.stride(0), _ptr(out), out.stride(0), rb, re, f, sp, _ptr(ws),
                  _ptr(order) if (order is not None and f <= NARROW_FEAT and (rb, re) == (0, n)) else None, _stream())
        _lib.count_launches(1 + extra)
    return out


def spmm_bwd_w(csr, et_csr, theta, alpha, norm, x, y, g, dx, rows=None, sides=3, split=None):
    """-> (d_theta[R] or None, d_norm[N] or None).  sides: bit0 source side scaled, bit1 destination."""
    x, y, g, dx = _f32(x), _f32(y), _f32(g), _f32(dx)
    n = csr['indptr'].numel() - 1
    rb, re = _rows(rows, n)
    weighted = theta is not None
    theta = _f32(theta).view(-1) if weighted else None
    r = theta.numel() if weighted else 0
    partials = torch.empty(max(_lib.partial_blocks(re - rb) * r, 1), dtype=torch.float64, device=x.device)
    d_theta = torch.empty(r, dtype=torch.float32, device=x.device) if weighted else None
    d_norm = torch.zeros(n, dtype=torch.float32, device=x.device) if norm is not None else None
    with torch.cuda.device(x.device):
        _lib.call('regnn_spmm_bwd_w', _ptr(csr['indptr']), _ptr(csr['indices']), _ptr(et_csr) if weighted else None,
                  _ptr(theta), float(alpha), r, _ptr(norm), int(sides), _ptr(x), x.stride(0), _ptr(y), y.stride(0),
                  _ptr(g), g.stride(0), _ptr(dx), dx.stride(0), rb, re, x.shape[1], _ptr(partials),
                  _ptr(d_theta), _ptr(d_norm), ctypes.byref(split['struct']) if split is not None else None, _stream())
        _lib.count_launches(2 if weighted else 1)
    return d_theta, d_norm


def spmm_bwd_fused(csr, et_t, theta, alpha, norm, x, g, rows=None, sides=3, out=None, want_xdx=False, y=None,
                   want_dnorm=False):
    """One gather pass over the transposed view -> (dX[N,F], d_theta[R], xdx[N] | d_norm[N] | None).
    ``want_dnorm`` (needs ``y`` = the forward result when the destination side is scaled, and a shape the lane-group
    kernel covers: ``fused_dnorm_fits``): the third result is the whole norm gradient, no row-dot pass needed;
    ``want_xdx``: it is <X[u],dX[u]> only (input of ``rowdot_norm_bwd``).  See regnn_spmm_bwd_fused."""
    x, g = _f32(x), _f32(g)
    theta = _f32(theta).view(-1)
    n = csr['indptr_t'].numel() - 1
    rb, re = _rows(rows, n)
    f, r = x.shape[1], theta.numel()
    if out is None:
        out = torch.empty((n, f), dtype=torch.float32, device=x.device)
        if (rb, re) != (0, n):
            out.zero_()
    partials = torch.empty(_lib.partial_blocks(re - rb) * r, dtype=torch.float64, device=x.device)
    d_theta = torch.empty(r, dtype=torch.float32, device=x.device)
    sp, ws, extra = _split_args(csr.get('split_t'), f, x.device)
    order = row_order(csr, True) if (f <= NARROW_FEAT and (rb, re) == (0, n)) else None
    d_norm = xdx = None
    if want_dnorm:
        if order is None or norm is None:
            raise RuntimeError('the folded norm gradient needs the lane-group kernel (F <= %d, full row range)' % NARROW_FEAT)
        y = _f32(y)
        d_norm = torch.empty(n, dtype=torch.float32, device=x.device)
    elif want_xdx:
        xdx = torch.zeros(n, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call('regnn_spmm_bwd_fused', _ptr(csr['indptr_t']), _ptr(csr['indices_t']), _ptr(et_t), _ptr(theta),
                  float(alpha), r, _ptr(norm), int(sides), _ptr(x), x.stride(0), _ptr(g), g.stride(0), _ptr(out),
                  out.stride(0), rb, re, f, _ptr(partials), _ptr(d_theta), _ptr(xdx), _ptr(y) if want_dnorm else None,
                  y.stride(0) if (want_dnorm and y is not None) else 0, _ptr(d_norm), sp, _ptr(ws), _ptr(order), _stream())
        _lib.count_launches(2 + extra * (2 if (want_xdx or want_dnorm) else 1))
    return out, d_theta, (d_norm if want_dnorm else xdx)


def fused_dnorm_fits(feat, rows_full=True):
    """True when regnn_spmm_bwd_fused can also produce d_norm (lane-group kernel: F <= NARROW_FEAT, F % 4 == 0)."""
    return rows_full and feat <= NARROW_FEAT and feat % 4 == 0


SMEM_BUDGET = 227 * 1024   # dynamic shared memory one block can opt in to on sm_100a


def fused_bwd_fits(feat, num_relations):
    """True when regnn_spmm_bwd_fused has a kernel for this shape: the lane-local relation bins (1088 bytes per
    relation) and -- for the whole-warp kernel, F > NARROW_FEAT -- the 8 x 8-row X tile must fit in shared memory.
    Otherwise the caller takes the two-pass backward (transposed regnn_spmm_fwd + regnn_spmm_bwd_w)."""
    bins = 1088 * int(num_relations)
    tile = 0 if feat <= NARROW_FEAT and feat % 4 == 0 else 64 + 256 * ((int(feat) + 3) & ~3)
    return feat <= 1024 and bins + tile <= SMEM_BUDGET


def rowdot_norm_bwd(norm, x, y, g, dx, rows=None, sides=3, xdx=None):
    """d_norm[v] = ([sides&2]<Y,G> + [sides&1]<X,dX>) / norm, rows outside the range are zero."""
    x, y, g, dx = _f32(x), _f32(y), _f32(g), _f32(dx)
    n = norm.numel()
    rb, re = _rows(rows, n)
    d_norm = torch.empty(n, dtype=torch.float32, device=x.device) if (rb, re) == (0, n) else \
        torch.zeros(n, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call('regnn_rowdot_norm_bwd', _ptr(norm), int(sides), _ptr(x), x.stride(0), _ptr(y), y.stride(0),
                  _ptr(g), g.stride(0), _ptr(dx), dx.stride(0), _ptr(xdx), rb, re, x.shape[1], _ptr(d_norm), _stream())
        _lib.count_launches(1)
    return d_norm


def rows_to_slabs(x_own, num_ranks, row_offset, peer_slab_ptrs):
    """Pushes this rank's row block ``x_own`` [rows, F] into every rank's column slab (regnn_rows_to_slabs).
    ``peer_slab_ptrs``: int64 device tensor [P] of peer-mapped slab base addresses."""
    x_own = _f32(x_own)
    with torch.cuda.device(x_own.device):
        _lib.call('regnn_rows_to_slabs', _ptr(x_own), x_own.stride(0), x_own.shape[0], x_own.shape[1], int(num_ranks),
                  int(row_offset), _ptr(peer_slab_ptrs), _stream())
        _lib.count_launches(1)


def slabs_to_rows(slab, num_rows, peers):
    """Pushes rows [0, num_rows) of this rank's column slab ``slab`` [>= num_rows, F/P] to their owner ranks' row blocks
    (regnn_slabs_to_rows; ``peers``: a ``_lib.PeerRows``)."""
    slab = _f32(slab)
    with torch.cuda.device(slab.device):
        _lib.call('regnn_slabs_to_rows', _ptr(slab), slab.stride(0), int(num_rows), slab.shape[1], ctypes.byref(peers),
                  _stream())
        _lib.count_launches(1)


def spmm_scatter(csr, etype, theta, alpha, norm_src, norm_dst, x_cols, peers, y_local=None, transposed=False):
    """regnn_spmm_fwd_scatter: forward SpMM of a column slab, result rows stored to their owner ranks (``peers``: a
    ``_lib.PeerRows``) and, with ``y_local`` ([rows >= N, F/P] fp32), also kept as a local slab.  ``transposed``: over
    the transposed view (the backward of the un-weighted propagation; ``etype`` is then the transposed-order array)."""
    x_cols = _f32(x_cols)
    sfx = '_t' if transposed else ''
    n = csr['indptr' + sfx].numel() - 1
    f = x_cols.shape[1]
    theta = _f32(theta).view(-1) if theta is not None else None
    sp, ws, extra = _split_args(csr.get('split' + sfx), f, x_cols.device)
    with torch.cuda.device(x_cols.device):
        _lib.call('regnn_spmm_fwd_scatter', _ptr(csr['indptr' + sfx]), _ptr(csr['indices' + sfx]),
                  _ptr(etype) if theta is not None else None, _ptr(theta), float(alpha),
                  theta.numel() if theta is not None else 0, _ptr(norm_src), _ptr(norm_dst), _ptr(x_cols),
                  x_cols.stride(0), n, f, sp, _ptr(ws), _ptr(row_order(csr, transposed)), ctypes.byref(peers), _ptr(y_local),
                  y_local.stride(0) if y_local is not None else 0, _stream())
        _lib.count_launches(1 + extra)


def spmm_bwd_fused_scatter(csr, et_t, theta, alpha, norm, x_cols, g_cols, peers, sides=3, y_cols=None):
    """regnn_spmm_bwd_fused_scatter -> (d_theta[R], d_norm[N] | None), both for this rank's columns only; the dX rows go
    to their owner ranks.  ``y_cols`` (the local result slab of ``spmm_scatter``) switches the folded norm gradient on."""
    x_cols, g_cols = _f32(x_cols), _f32(g_cols)
    theta = _f32(theta).view(-1)
    n = csr['indptr_t'].numel() - 1
    f, r = x_cols.shape[1], theta.numel()
    partials = torch.empty(_lib.partial_blocks(n) * r, dtype=torch.float64, device=x_cols.device)
    d_theta = torch.empty(r, dtype=torch.float32, device=x_cols.device)
    d_norm = torch.empty(n, dtype=torch.float32, device=x_cols.device) if (y_cols is not None and norm is not None) else None
    sp, ws, extra = _split_args(csr.get('split_t'), f, x_cols.device)
    with torch.cuda.device(x_cols.device):
        _lib.call('regnn_spmm_bwd_fused_scatter', _ptr(csr['indptr_t']), _ptr(csr['indices_t']), _ptr(et_t), _ptr(theta),
                  float(alpha), r, _ptr(norm), int(sides), _ptr(x_cols), x_cols.stride(0), _ptr(g_cols),
                  g_cols.stride(0), n, f, _ptr(partials), _ptr(d_theta), None,
                  _ptr(y_cols) if d_norm is not None else None, y_cols.stride(0) if d_norm is not None else 0,
                  _ptr(d_norm), sp, _ptr(ws), _ptr(row_order(csr, True)), ctypes.byref(peers), _stream())
        _lib.count_launches(2 + extra)
    return d_theta, d_norm


def _attn_split(split, h, d, device):
    """(byref(struct) | None, workspace | None, extra finalize launches flag) for the attention kernels."""
    if split is None:
        return None, None
    ws = torch.empty(split['struct'].num_frags * (h * d + 2 * h), dtype=torch.float32, device=device)
    return ctypes.byref(split['struct']), ws


def _rel(theta, et_csr):
    if theta is None or et_csr is None:
        return None, None, 0
    theta = _f32(theta)
    return theta, et_csr, theta.shape[0]


def _full(rb, re, n):
    return (rb, re) == (0, n)


def gat_fwd(csr, et_csr, theta, alpha, feat, el, er, slope, keep=None, want_attn=False, rows=None):
    feat, el, er, keep = _f32(feat), _f32(el), _f32(er), _f32(keep)
    n, h, d = feat.shape
    rb, re = _rows(rows, n)
    theta, et_csr, r = _rel(theta, et_csr)
    dev = feat.device
    out = torch.empty_like(feat) if _full(rb, re, n) else torch.zeros_like(feat)
    rowmax = torch.zeros((n, h), dtype=torch.float32, device=dev)
    rowsum = torch.zeros((n, h), dtype=torch.float32, device=dev)
    e = csr['indices'].numel()
    attn = torch.zeros((max(e, 1), h), dtype=torch.float32, device=dev)[:e] if want_attn else None
    sp, ws = _attn_split(csr.get('split'), h, d, dev)
    order = row_order(csr) if _full(rb, re, n) else None
    with torch.cuda.device(dev):
        _lib.call('regnn_gat_fwd', _ptr(csr['indptr']), _ptr(csr['indices']), _ptr(csr['eid']), _ptr(et_csr),
                  _ptr(theta), float(alpha), r, _ptr(feat), _ptr(el), _ptr(er), float(slope), _ptr(keep), h, d,
                  rb, re, _ptr(out), _ptr(rowmax), _ptr(rowsum), _ptr(attn), sp, _ptr(ws), _ptr(order), _stream())
        _lib.count_launches(1 + (sp is not None))
    return out, rowmax, rowsum, attn


def gat_bwd(csr, et_csr, et_t, theta, alpha, feat, el, er, slope, keep, out, rowmax, rowsum, g, attn_l=None,
            attn_r=None, rows=None):
    """Backward of ``gat_fwd`` as one gather pass (regnn_gat_bwd_stats -> regnn_gat_bwd_edges -> regnn_gat_bwd_reduce)
    -> (d_feat, d_el, d_er, d_theta | None, d_attn_l | None, d_attn_r | None).
    With ``attn_l`` / ``attn_r`` ([H*D], the projection-score vectors of layer/REGATConv.py:68-69, el = <feat, attn_l>,
    er = <feat, attn_r>) the gradients through the scores are folded in: ``d_feat`` is the total feature gradient and
    the two parameter gradients are returned (regnn_attn_scores_bwd)."""
    feat, el, er, keep, g, out = _f32(feat), _f32(el), _f32(er), _f32(keep), _f32(g), _f32(out)
    n, h, d = feat.shape
    rb, re = _rows(rows, n)
    theta, et_csr, r = _rel(theta, et_csr)
    dev = feat.device
    e = csr['indices'].numel()
    full = _full(rb, re, n)
    fold = attn_l is not None
    if fold:
        attn_l, attn_r = _f32(attn_l).view(-1), _f32(attn_r).view(-1)
    stats = torch.empty((n, h, 4), dtype=torch.float32, device=dev)
    dpre_csr = torch.empty((max(e, 1), h), dtype=torch.float32, device=dev)[:e]
    d_feat = torch.empty_like(g) if full else torch.zeros_like(g)
    d_el = torch.zeros((n, h), dtype=torch.float32, device=dev)
    d_er = torch.zeros((n, h), dtype=torch.float32, device=dev)
    partials = torch.empty(max(_lib.partial_blocks(n) * max(r * h, 2 * h * d if fold else 0), 1), dtype=torch.float64,
                           device=dev)
    d_theta = torch.zeros((r, h), dtype=torch.float32, device=dev) if r else None   # stays 0 when the graph has no edges
    sp_t, ws_t = _attn_split(csr.get('split_t'), h, d, dev)
    sp, ws = _attn_split(csr.get('split'), h, 0, dev)
    with torch.cuda.device(dev):
        _lib.call('regnn_gat_bwd_stats', _ptr(out), _ptr(g), _ptr(er), _ptr(rowmax), _ptr(rowsum), n, h, d, _ptr(stats),
                  _stream())
        _lib.call('regnn_gat_bwd_edges', _ptr(csr['indptr_t']), _ptr(csr['indices_t']), _ptr(csr['slot_t']),
                  _ptr(et_t) if r else None, _ptr(csr['eid']), _ptr(theta), float(alpha), r, _ptr(feat), _ptr(el),
                  _ptr(stats), float(slope), _ptr(keep), _ptr(g), h, d, rb, re, _ptr(d_feat), _ptr(d_el), _ptr(dpre_csr),
                  _ptr(attn_l) if fold else None, sp_t, _ptr(ws_t), _ptr(row_order(csr, True)) if full else None, _stream())
        _lib.call('regnn_gat_bwd_reduce', _ptr(csr['indptr']), _ptr(et_csr) if r else None, _ptr(theta), float(alpha), r,
                  _ptr(dpre_csr), h, rb, re, _ptr(d_er), _ptr(partials), _ptr(d_theta), sp, _ptr(ws), _stream())
        _lib.count_launches(3 + ((1 if re - rb >= 65536 else 2) if r else 0) + (3 if sp_t is not None else 0) + (sp is not None))
        d_al = d_ar = None
        if fold:
            d_al = torch.empty(h * d, dtype=torch.float32, device=dev)
            d_ar = torch.empty(h * d, dtype=torch.float32, device=dev)
            _lib.call('regnn_attn_scores_bwd', _ptr(feat), _ptr(d_el), _ptr(d_er), n, h, d, _ptr(partials), _ptr(d_al),
                      _ptr(d_ar), _ptr(d_feat), _ptr(attn_r), _stream())
            _lib.count_launches(3)
    return d_feat, d_el, d_er, d_theta, d_al, d_ar


def attn_scores_fwd(feat, attn_l, attn_r):
    """el[n,h] = <feat[n,h,:], attn_l[h,:]>, er likewise (regnn_attn_scores_fwd)."""
    feat = _f32(feat)
    n, h, d = feat.shape
    attn_l, attn_r = _f32(attn_l).view(-1), _f32(attn_r).view(-1)
    el = torch.empty((n, h), dtype=torch.float32, device=feat.device)
    er = torch.empty((n, h), dtype=torch.float32, device=feat.device)
    with torch.cuda.device(feat.device):
        _lib.call('regnn_attn_scores_fwd', _ptr(feat), _ptr(attn_l), _ptr(attn_r), n, h, d, _ptr(el), _ptr(er), _stream())
        _lib.count_launches(1)
    return el, er


def gatv2_fwd(csr, et_csr, theta, alpha, fs, fd, attn, slope, keep=None, want_attn=False, rows=None, save=False):
    """-> (out, rowmax, rowsum, attention | None, saved | None).  ``save``: also store what the backward needs, the raw
    logit of every (CSR slot, head) and the 128-bit sign mask of fs[src]+fd[dst] per (slot, 128-float slice):
    ``saved = (logit_csr [E,H], qmask [E, ceil(H*D/128), 4] int32)``."""
    fs, fd, keep = _f32(fs), _f32(fd), _f32(keep)
    attn = _f32(attn).view(-1)
    n, h, d = fd.shape
    rb, re = _rows(rows, n)
    theta, et_csr, r = _rel(theta, et_csr)
    dev = fs.device
    out = torch.empty_like(fd) if (rb, re) == (0, n) else torch.zeros_like(fd)
    rowmax = torch.zeros((n, h), dtype=torch.float32, device=dev)
    rowsum = torch.zeros((n, h), dtype=torch.float32, device=dev)
    e = csr['indices'].numel()
    att = torch.zeros((max(e, 1), h), dtype=torch.float32, device=dev)[:e] if want_attn else None
    saved = None
    if save and e * max(h, (h * d + 127) // 128) >= 2 ** 32:
        raise ValueError('gatv2_fwd(save=True): E * max(H, ceil(H*D/128)) = %d exceeds the 32-bit slot offsets of the kernel'
                         % (e * max(h, (h * d + 127) // 128)))
    if save:
        saved = (torch.empty((max(e, 1), h), dtype=torch.float32, device=dev),    # max(E, 1): never a NULL pointer
                 torch.empty((max(e, 1), (h * d + 127) // 128, 4), dtype=torch.int32, device=dev))
    sp, ws = _attn_split(csr.get('split'), h, d, dev)
    order = row_order(csr) if _full(rb, re, n) else None
    with torch.cuda.device(dev):
        _lib.call('regnn_gatv2_fwd', _ptr(csr['indptr']), _ptr(csr['indices']), _ptr(csr['eid']), _ptr(et_csr),
                  _ptr(theta), float(alpha), r, _ptr(fs), _ptr(fd), _ptr(attn), float(slope), _ptr(keep), h, d,
                  rb, re, _ptr(out), _ptr(rowmax), _ptr(rowsum), _ptr(att), _ptr(saved[0]) if save else None,
                  _ptr(saved[1]) if save else None, sp, _ptr(ws), _ptr(order), _stream())
        _lib.count_launches(1 + (sp is not None))
    return out, rowmax, rowsum, att, saved


V2_STAGE_CHUNKS = 592   # REGNN_V2_STAGE_CHUNKS (csrc/gat.cu)


def gatv2_bwd(csr, et_csr, theta, alpha, fs, fd, attn, slope, keep, out, rowmax, rowsum, saved, g, rows=None):
    """Backward of ``gatv2_fwd(..., save=True)``: regnn_gat_bwd_stats -> regnn_gatv2_bwd_edges (the one gather pass) ->
    regnn_gatv2_bwd_dst (streaming) -> (d_fs, d_fd, d_attn [H*D], d_theta | None)."""
    fs, fd, keep, g, out = _f32(fs), _f32(fd), _f32(keep), _f32(g), _f32(out)
    attn = _f32(attn).view(-1)
    logit_csr, qmask = saved
    n, h, d = fd.shape
    rb, re = _rows(rows, n)
    theta, et_csr, r = _rel(theta, et_csr)
    dev = fs.device
    e = csr['indices'].numel()
    full = _full(rb, re, n)
    stats = torch.empty((n, h, 4), dtype=torch.float32, device=dev)
    dl_csr = torch.empty((max(e, 1), h), dtype=torch.float32, device=dev)
    d_fs = torch.empty_like(g) if full else torch.zeros_like(g)
    d_fd = torch.empty_like(fd) if full else torch.zeros_like(fd)
    d_attn_src = torch.empty(h * d, dtype=torch.float32, device=dev)
    d_attn_dst = torch.empty(h * d, dtype=torch.float32, device=dev)
    d_theta = torch.zeros((r, h), dtype=torch.float32, device=dev) if r else None   # stays 0 when the graph has no edges
    split_t, split = csr.get('split_t'), csr.get('split')
    sp_t, ws_t = _attn_split(split_t, h, d, dev)
    sp, ws = _attn_split(split, h, d, dev)
    order_t = row_order(csr, True) if full else None
    nblocks = _lib.load().regnn_gatv2_bwd_edges_blocks(re - rb, split_t['struct'].num_frags if split_t else 0,
                                                       split_t['struct'].num_long if split_t else 0, h, d,
                                                       int(order_t is not None))
    block_partials = torch.empty(max(nblocks, 1) * h * d, dtype=torch.float32, device=dev)
    partials = torch.empty(max(_lib.partial_blocks() * h * d, V2_STAGE_CHUNKS * max(r * h, h * d)), dtype=torch.float64,
                           device=dev)
    with torch.cuda.device(dev):
        _lib.call('regnn_gat_bwd_stats', _ptr(out), _ptr(g), None, _ptr(rowmax), _ptr(rowsum), n, h, d, _ptr(stats),
                  _stream())
        _lib.call('regnn_gatv2_bwd_edges', _ptr(csr['indptr_t']), _ptr(csr['indices_t']), _ptr(csr['slot_t']),
                  _ptr(csr['eid']), _ptr(fs), _ptr(stats), _ptr(logit_csr), _ptr(qmask), _ptr(attn), float(slope),
                  _ptr(keep), _ptr(g), h, d, rb, re, _ptr(d_fs), _ptr(dl_csr), _ptr(d_attn_src), _ptr(block_partials),
                  _ptr(partials), sp_t, _ptr(ws_t), _ptr(order_t), _stream())
        _lib.call('regnn_gatv2_bwd_dst', _ptr(csr['indptr']), _ptr(et_csr) if r else None, _ptr(theta), float(alpha), r,
                  _ptr(fd), _ptr(dl_csr), _ptr(qmask), _ptr(attn), float(slope), h, d, rb, re, _ptr(d_fd),
                  _ptr(d_attn_dst), _ptr(partials), _ptr(d_theta), sp, _ptr(ws), _stream())
        _lib.count_launches(6 + (2 if r else 0) + (sp_t is not None) + (sp is not None))
    return d_fs, d_fd, d_attn_src + d_attn_dst, d_theta


# ---- grouped per-node-type input projection ---------------------------------------------------------------------
GROUPED_BWD_SPLITS = 8


def _host_ptrs(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    return arr


def _grouped_common(tables):
    t = len(tables)
    ldx = (ctypes.c_int64 * t)(*[x.stride(0) for x in tables])
    k = (ctypes.c_int * t)(*[x.shape[1] for x in tables])
    return ldx, k


def grouped_linear_fwd(tables, weights, biases, seg_ptr, perm, local_idx, num_rows):
    """out[i] = W[t(i)] x_{t(i)}[row(i)] + b[t(i)] for all node types in one launch (regnn_grouped_linear_fwd).
    ``tables`` / ``weights`` / ``biases``: per-type lists ([n_t, K_t] fp32; nn.Linear [n_out, K_t]; [n_out] or None);
    ``seg_ptr`` int32 [T+1] on the device; ``perm`` / ``local_idx`` int64 device tensors or None."""
    tables = [_f32(x) for x in tables]
    weights = [_f32(w) for w in weights]
    biases = [_f32(b) if b is not None else None for b in biases]
    n_out = weights[0].shape[0]
    dev = tables[0].device
    out = torch.empty((num_rows, n_out), dtype=torch.float32, device=dev)
    ldx, k = _grouped_common(tables)
    with torch.cuda.device(dev):
        _lib.call('regnn_grouped_linear_fwd', len(tables), _host_ptrs(tables), ldx, k, _host_ptrs(weights),
                  _host_ptrs(biases), n_out, int(num_rows), _ptr(seg_ptr), _ptr(perm), _ptr(local_idx), _ptr(out),
                  out.stride(0), _stream())
        _lib.count_launches(1)
    return out


def grouped_linear_bwd(tables, n_out, seg_ptr, perm, local_idx, dout, want_bias):
    """-> (per-type weight gradients [n_out, K_t], per-type bias gradients [n_out] | None): regnn_grouped_linear_bwd
    writes ``GROUPED_BWD_SPLITS`` deterministic partials per type, added here in split order."""
    tables = [_f32(x) for x in tables]
    dout = _f32(dout)
    dev = dout.device
    s = GROUPED_BWD_SPLITS
    dwp = [torch.empty((s, n_out, x.shape[1]), dtype=torch.float32, device=dev) for x in tables]
    dbp = [torch.empty((s, n_out), dtype=torch.float32, device=dev) for _ in tables] if want_bias else None
    ldx, k = _grouped_common(tables)
    with torch.cuda.device(dev):
        _lib.call('regnn_grouped_linear_bwd', len(tables), _host_ptrs(tables), ldx, k, n_out, dout.shape[0], _ptr(seg_ptr),
                  _ptr(perm), _ptr(local_idx), _ptr(dout), dout.stride(0), s, _host_ptrs(dwp),
                  _host_ptrs(dbp) if want_bias else None, _stream())
        _lib.count_launches(1)
    return [p.sum(dim=0) for p in dwp], ([p.sum(dim=0) for p in dbp] if want_bias else None)
