"""Destination-row-block partitioning of the full-batch REGCN aggregation across ranks
(BASELINE config 4: ogbn-mag-shaped graph over 2/4/8 B200).  The reference has no multi-GPU code
(SURVEY.md 2.4); numerics must equal the single-device result.

Scheme (owner computes, no floating-point reduction across ranks except the tiny relation-gradient
table):
  * rank p owns the contiguous destination rows [r_p, r_{p+1}), balanced by in-edge count;
  * every rank holds the whole CSR / transposed view (int32 structures: ~0.4 GB at MAG scale) and
    computes the (cheap, E-byte) relation-weighted degree norm for all rows;
  * forward:  NCCL all-gather of the owned source-feature rows -> full X; fused SpMM on owned rows;
  * backward: NCCL all-gather of the owned dL/dY rows -> full G; transposed SpMM on owned SOURCE rows
    gives the owned dX rows (no reduce-scatter, no atomics); the relation-gradient partials of the
    owned rows are summed with one all-reduce of R floats.

``feature_sliced_propagate`` is the second scheme with the same row-block-in / row-block-out contract: the
aggregation runs on column slabs (all rows, F/P columns per rank) between two all-to-alls, which moves P times
fewer bytes over NVLink than the all-gathers above.
"""
import torch
import torch.distributed as dist

from . import ops


def row_blocks(indptr, parts, balance='edges'):
    """Row boundaries [r_0=0, ..., r_P=N].  ``balance='edges'``: ~equal in-edge counts per block (prefix sums of
    indptr), uneven row counts.  ``balance='rows'``: ceil(N/P) rows per block -- compute is less balanced, but the
    exchange becomes ONE even ``all_gather_into_tensor`` straight into the row-indexed buffer, which is what
    matters when the step is exchange-bound."""
    n = indptr.numel() - 1
    if balance == 'rows':
        per = (n + parts - 1) // parts
        return [min(n, p * per) for p in range(parts)] + [n]
    e = int(indptr[-1].item())
    targets = torch.arange(1, parts, device=indptr.device, dtype=torch.int64) * e // parts
    cuts = torch.searchsorted(indptr.to(torch.int64), targets)
    bounds = [0] + [int(c) for c in cuts.clamp(max=n).tolist()] + [n]
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds


def _uniform_rows(bounds):
    parts = len(bounds) - 1
    per = bounds[1] - bounds[0]
    return per if per > 0 and all(bounds[p] == min(bounds[-1], p * per) for p in range(parts)) else 0


def _alloc_full(n, f, bounds, like):
    """[N, F] buffer indexed by global row id; over-allocated to P*ceil(N/P) rows for the even all-gather."""
    per = _uniform_rows(bounds)
    rows = max(n, per * (len(bounds) - 1)) if per else n
    return torch.empty((rows, f), dtype=like.dtype, device=like.device)


def _all_gather_rows(full, own, bounds, group):
    """Gathers every rank's row block into ``full`` (global row ids).  Equal-size blocks: one
    ``all_gather_into_tensor``.  Uneven blocks (balanced by edges): ProcessGroupNCCL's grouped broadcasts."""
    per = _uniform_rows(bounds)
    if per and dist.get_backend(group) == 'nccl' and full.shape[0] >= per * (len(bounds) - 1):
        send = own.contiguous()
        if send.shape[0] != per:      # last block: pad to the common size
            send = torch.cat([send, send.new_zeros((per - send.shape[0], send.shape[1]))])
        dist.all_gather_into_tensor(full[:per * (len(bounds) - 1)], send, group=group)
        return full
    views = [full[bounds[p]:bounds[p + 1]] for p in range(len(bounds) - 1)]
    if dist.get_backend(group) == 'nccl':
        dist.all_gather(views, own.contiguous(), group=group)
    else:  # gloo (CPU tests) has no uneven all-gather: one broadcast per owner
        rank = dist.get_rank(group)
        views[rank].copy_(own)
        for p, v in enumerate(views):
            if v.numel():
                dist.broadcast(v, src=dist.get_global_rank(group, p) if group is not None else p, group=group)
    return full


class _PartitionedPropagate(torch.autograd.Function):
    """Rows [rb, re) of  Y = norm (.) A_w (norm (.) X)  with X row-partitioned across ranks."""

    @staticmethod
    def forward(ctx, graph, etv, x_own, theta, alpha, norm, bounds, rank, group):
        csr = graph.csr()
        n = graph.number_of_nodes()
        rb, re = bounds[rank], bounds[rank + 1]
        x_full = _alloc_full(n, x_own.shape[1], bounds, x_own)
        _all_gather_rows(x_full, x_own, bounds, group)
        y_full = torch.empty_like(x_full)
        ops.spmm(csr['indptr'], csr['indices'], etv[0], theta, alpha, norm, norm, x_full, rows=(rb, re), out=y_full,
                 split=csr.get('split'))
        ctx.graph, ctx.etv, ctx.alpha, ctx.bounds, ctx.rank, ctx.group = graph, etv, alpha, bounds, rank, group
        ctx.save_for_backward(x_full, y_full, theta, norm)
        return y_full[rb:re]

    @staticmethod
    def backward(ctx, g_own):
        x_full, y_full, theta, norm = ctx.saved_tensors
        csr = ctx.graph.csr()
        bounds, rank = ctx.bounds, ctx.rank
        rb, re = bounds[rank], bounds[rank + 1]
        g_full = torch.empty_like(x_full)
        _all_gather_rows(g_full, g_own, bounds, ctx.group)
        dx_full = torch.empty_like(x_full)
        # owner computes: the transposed rows [rb, re) are this rank's SOURCE rows; one fused gather pass gives
        # their dX and this rank's share of the relation gradient (each edge belongs to exactly one source row)
        _, d_theta, xdx = ops.spmm_bwd_fused(csr, ctx.etv[1], theta, ctx.alpha, norm, x_full, g_full, rows=(rb, re),
                                             out=dx_full, want_xdx=True)
        d_norm = ops.rowdot_norm_bwd(norm, x_full, y_full, g_full, dx_full, rows=(rb, re), xdx=xdx)
        # d_theta / d_norm hold this rank's rows only; the caller all-reduces the parameter gradient once.
        return None, None, dx_full[rb:re], d_theta.view_as(theta), None, d_norm, None, None, None


def partitioned_propagate(graph, etv, x_own, theta, alpha, norm, bounds, rank, group=None):
    return _PartitionedPropagate.apply(graph, etv, x_own, theta, alpha, norm, bounds, rank, group)


def _rows_to_columns(own, per, parts, group):
    """[rows_own, F] row block (all columns)  ->  [P*per, F/P] column slab (all rows, this rank's columns).
    One all-to-all: block q of the send buffer holds this rank's rows restricted to rank q's columns, and
    lands on rank q at row offset rank*per -- the receive buffer IS the slab, no unpack pass."""
    fc = own.shape[1] // parts
    pad = own if own.shape[0] == per else torch.cat([own, own.new_zeros((per - own.shape[0], own.shape[1]))])
    send = pad.view(per, parts, fc).permute(1, 0, 2).contiguous()
    slab = torch.empty((parts * per, fc), dtype=own.dtype, device=own.device)
    dist.all_to_all_single(slab, send.view(parts * per, fc), group=group)
    return slab


def _columns_to_rows(slab, per, parts, rows_own, group):
    """Inverse of ``_rows_to_columns``: the slab is sent as is (block q = rank q's rows), the received
    [P, per, F/P] pieces are interleaved back into [rows_own, F]."""
    fc = slab.shape[1]
    recv = torch.empty_like(slab)
    dist.all_to_all_single(recv, slab, group=group)
    return recv.view(parts, per, fc).permute(1, 0, 2).reshape(per, parts * fc)[:rows_own]


class _ColumnSlabPropagate(torch.autograd.Function):
    """Same contract as ``_PartitionedPropagate`` (row block in, row block out), but the aggregation itself runs
    feature-sliced: every rank propagates ALL rows for F/P of the columns, so the SpMM and its backward need
    no remote rows at all.  The row<->column re-partition is an all-to-all that moves (P-1)/P of ONE row block
    per rank (N*F/P floats) instead of the all-gather's (P-1) row blocks -- P times less NVLink traffic; the
    only extra exchange is the all-reduce of the N-float norm gradient (its row dot products span all columns)."""

    @staticmethod
    def forward(ctx, graph, etv, x_own, theta, alpha, norm, bounds, rank, group):
        csr = graph.csr()
        parts = len(bounds) - 1
        per = _uniform_rows(bounds)
        if not per or x_own.shape[1] % parts:
            raise ValueError('feature-sliced propagation needs equal row blocks (row_blocks(..., balance="rows")) '
                             'and a feature width divisible by the number of ranks')
        x_cols = _rows_to_columns(x_own, per, parts, group)
        y_cols = torch.empty_like(x_cols)
        ops.spmm(csr['indptr'], csr['indices'], etv[0], theta, alpha, norm, norm, x_cols, out=y_cols,
                 split=csr.get('split'), order=ops.row_order(csr) if x_cols.shape[1] <= ops.NARROW_FEAT else None)
        ctx.graph, ctx.etv, ctx.alpha, ctx.bounds, ctx.rank, ctx.group = graph, etv, alpha, bounds, rank, group
        ctx.save_for_backward(x_cols, y_cols, theta, norm)
        return _columns_to_rows(y_cols, per, parts, x_own.shape[0], group)

    @staticmethod
    def backward(ctx, g_own):
        x_cols, y_cols, theta, norm = ctx.saved_tensors
        csr = ctx.graph.csr()
        bounds, rank = ctx.bounds, ctx.rank
        parts, per = len(bounds) - 1, _uniform_rows(bounds)
        rb, re = bounds[rank], bounds[rank + 1]
        g_cols = _rows_to_columns(g_own.contiguous(), per, parts, ctx.group)
        dx_cols = torch.empty_like(x_cols)
        _, d_theta, xdx = ops.spmm_bwd_fused(csr, ctx.etv[1], theta, ctx.alpha, norm, x_cols, g_cols, out=dx_cols,
                                             want_xdx=True)
        # row dot products over this rank's columns only; the sum over ranks completes them
        d_norm = ops.rowdot_norm_bwd(norm, x_cols, y_cols, g_cols, dx_cols, xdx=xdx)
        dist.all_reduce(d_norm, op=dist.ReduceOp.SUM, group=ctx.group)
        # keep the owned rows: the norm's own backward then yields a per-rank share of the relation gradient,
        # like d_theta above (this rank's columns), and the caller all-reduces the parameter gradient once
        own = torch.zeros_like(d_norm)
        own[rb:re] = d_norm[rb:re]
        dx_own = _columns_to_rows(dx_cols, per, parts, re - rb, ctx.group)
        return None, None, dx_own, d_theta.view_as(theta), None, own, None, None, None


def feature_sliced_propagate(graph, etv, x_own, theta, alpha, norm, bounds, rank, group=None):
    return _ColumnSlabPropagate.apply(graph, etv, x_own, theta, alpha, norm, bounds, rank, group)


def allreduce_relation_grads(params, group=None):
    """Sums the per-rank relation-embedding gradients (R x H floats each) -- the only cross-rank
    floating-point reduction of the partitioned layer."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
