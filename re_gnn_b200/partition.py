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
