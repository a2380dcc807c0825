"""Multi-GPU execution of the full-batch REGCN aggregation (BASELINE config 4: ogbn-mag-shaped graph over 2/4/8
B200).  The reference has no multi-GPU code; numerics must equal the single-device result (Y and dX
are bit-identical in every scheme below: all kernels add a row's slots in slot order).

Common contract: rank p holds the row block [r_p, r_{p+1}) of X and dL/dY and gets back the same rows of Y and dX;
every rank holds the whole CSR / transposed view (int32 structures, ~0.4 GB at MAG scale) and computes the cheap
relation-degree norm for all rows; the caller sums the R-float relation gradient once (``allreduce_relation_grads``).

1. ``partitioned_propagate`` -- destination-row blocks (owner computes):
   forward:  NCCL all-gather of the owned source rows -> full X; fused SpMM on the owned rows;
   backward: NCCL all-gather of the owned dL/dY rows -> full G; transposed SpMM on the owned SOURCE rows gives the
   owned dX rows (no reduce-scatter, no atomics).  Exchange-bound: (P-1)/P of the [N,F] matrix per rank, twice.
2. ``feature_sliced_propagate`` -- column slabs: every rank propagates ALL rows for F/P columns, so the SpMM and its
   backward need no remote rows; a row<->slab re-partition (N*F/P floats per rank, P times less) sits on each side:
   * ``exchange=None``: ``all_to_all_single`` (NCCL; gloo in the CPU tests) plus pack / unpack copies;
   * ``exchange=SlabExchange(...)``: our own kernels over NVLink peer memory -- ``regnn_rows_to_slabs`` pushes row
     blocks into the peers' slabs, the ``regnn_spmm_*_scatter`` epilogues store finished rows into their owners'
     row blocks; torch symmetric memory supplies the mapped buffers and the device-side barrier.
   The norm gradient's row dots span all columns, but the row owner has complete rows: ``d_norm`` is a local pass.
"""
import torch
import torch.distributed as dist

from . import _lib, ops


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


def _own_rows_norm_grad(norm, x_own, y_own, g_own, dx_own, rb, re):
    """d_norm (length N, zero outside [rb, re)) from the owner's complete rows."""
    d_norm = torch.zeros_like(norm)
    if re > rb:
        d_norm[rb:re] = ops.rowdot_norm_bwd(norm[rb:re].contiguous(), x_own, y_own, g_own, dx_own)
    return d_norm


class _ColumnSlabPropagate(torch.autograd.Function):
    """Same contract as ``_PartitionedPropagate`` (row block in, row block out), but the aggregation itself runs
    feature-sliced: every rank propagates ALL rows for F/P of the columns, so the SpMM and its backward need
    no remote rows at all.  The row<->column re-partition is an all-to-all that moves (P-1)/P of ONE row block
    per rank (N*F/P floats) instead of the all-gather's (P-1) row blocks -- P times less NVLink traffic; the
    norm gradient's row dot products span all columns, but the row owner holds complete rows: a local pass."""

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
        weighted = theta is not None
        ops.spmm(csr['indptr'], csr['indices'], etv[0] if weighted else None, theta, alpha, norm, norm, x_cols, out=y_cols,
                 split=csr.get('split'), order=ops.row_order(csr) if x_cols.shape[1] <= ops.NARROW_FEAT else None)
        ctx.graph, ctx.etv, ctx.alpha, ctx.bounds, ctx.rank, ctx.group = graph, etv, alpha, bounds, rank, group
        ctx.weighted = weighted
        y_own = _columns_to_rows(y_cols, per, parts, x_own.shape[0], group)
        ctx.x_own, ctx.y_own = x_own.detach(), y_own
        ctx.save_for_backward(x_cols, theta, norm)
        return y_own

    @staticmethod
    def backward(ctx, g_own):
        x_cols, theta, norm = ctx.saved_tensors
        csr = ctx.graph.csr()
        bounds, rank = ctx.bounds, ctx.rank
        parts, per = len(bounds) - 1, _uniform_rows(bounds)
        rb, re = bounds[rank], bounds[rank + 1]
        g_own = g_own.contiguous()
        g_cols = _rows_to_columns(g_own, per, parts, ctx.group)
        dx_cols = torch.empty_like(x_cols)
        d_theta = None
        if ctx.weighted:
            _, d_theta, _ = ops.spmm_bwd_fused(csr, ctx.etv[1], theta, ctx.alpha, norm, x_cols, g_cols, out=dx_cols)
            d_theta = d_theta.view_as(theta)
        else:   # un-weighted propagation (REMixHop's copy_u): the transposed SpMM alone
            ops.spmm(csr['indptr_t'], csr['indices_t'], None, None, ctx.alpha, norm, norm, g_cols, out=dx_cols,
                     split=csr.get('split_t'), order=ops.row_order(csr, True) if g_cols.shape[1] <= ops.NARROW_FEAT else None)
        dx_own = _columns_to_rows(dx_cols, per, parts, re - rb, ctx.group)
        # the row dot products of d_norm span all columns: the row owner has them all (x, y, dL/dY, dX row blocks),
        # so this is a local pass; the norm's own backward then yields a per-rank share of the relation gradient,
        # like d_theta above (this rank's columns), and the caller all-reduces the parameter gradient once
        d_norm = _own_rows_norm_grad(norm, ctx.x_own, ctx.y_own, g_own, dx_own, rb, re) if norm is not None else None
        return None, None, dx_own, d_theta, None, d_norm, None, None, None


class SlabExchange:
    """Peer-mapped (torch symmetric memory over NVLink) buffers of ONE feature-sliced aggregation call site: the
    two column slabs (X and dL/dY, [P*per, F/P]) that peers push their row blocks into, the two row blocks
    (Y and dX, [per, F]) that the SpMM kernels of all ranks store finished rows into, and a [P, 256]-float table
    every rank writes its share of the relation gradient into (``allreduce_relation_grads``).  ``y_cols`` is a plain
    local slab: this rank's columns of Y, kept for the folded norm gradient of the backward kernel.  Create it once
    per layer; a forward must be followed by its backward before the next forward (the buffers are reused)."""

    GRAD_SLOTS = 256

    def __init__(self, feat, bounds, rank, device, group=None):
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.parts, self.per, self.rank, self.feat = len(bounds) - 1, _uniform_rows(bounds), rank, feat
        if not self.per or feat % (4 * self.parts):
            raise ValueError('SlabExchange needs equal row blocks and F divisible by 4 * ranks')
        self.fc = feat // self.parts
        self._handles = []

        def alloc(rows, cols):
            t = symm.empty((rows, cols), dtype=torch.float32, device=device)
            h = symm.rendezvous(t, group)
            self._handles.append(h)
            return t, torch.tensor([int(p) for p in h.buffer_ptrs], dtype=torch.int64, device=device)

        self.x_cols, self.x_ptrs = alloc(self.parts * self.per, self.fc)
        self.g_cols, self.g_ptrs = alloc(self.parts * self.per, self.fc)
        self.y_rows, self.y_ptrs = alloc(self.per, feat)
        self.dx_rows, self.dx_ptrs = alloc(self.per, feat)
        self.grad_tab, self.grad_ptrs = alloc(self.parts, self.GRAD_SLOTS)
        self.y_cols = torch.zeros((self.parts * self.per, self.fc), dtype=torch.float32, device=device)
        self.x_cols.zero_()
        self.g_cols.zero_()
        self.grad_tab.zero_()
        self.peer_y = _lib.PeerRows(self.y_ptrs.data_ptr(), self.parts, self.per, feat, rank * self.fc)
        self.peer_dx = _lib.PeerRows(self.dx_ptrs.data_ptr(), self.parts, self.per, feat, rank * self.fc)
        self.barrier()

    def barrier(self):
        """Device-side barrier across the ranks on the current stream: everything the ranks wrote into each
        other's buffers before it is visible after it."""
        with _lib.phase('peer_barrier'):
            self._handles[0].barrier()

    def allreduce_small(self, flat):
        """Deterministic sum of a small fp32 vector (<= GRAD_SLOTS) over the ranks: every rank pushes its share into
        row ``rank`` of every peer's table (regnn_rows_to_slabs with a one-row block), a barrier, then each rank
        reduces the P rows with one fixed-shape reduction -- the same order, hence the same bits, on every rank and
        for every topology (no ring / tree of a collective library decides the association)."""
        n = flat.numel()
        if n > self.GRAD_SLOTS:
            raise ValueError('at most %d floats' % self.GRAD_SLOTS)
        if not hasattr(self, '_grad_row'):
            self._grad_row = torch.zeros((1, self.GRAD_SLOTS * self.parts), dtype=torch.float32, device=flat.device)
        self._grad_row.view(self.parts, self.GRAD_SLOTS)[:, :n] = flat.view(1, -1)     # the same share goes to every peer
        ops.rows_to_slabs(self._grad_row, self.parts, self.rank, self.grad_ptrs)
        self.barrier()
        return self.grad_tab[:, :n].sum(dim=0)


class _PeerSlabPropagate(torch.autograd.Function):
    """``_ColumnSlabPropagate`` with both re-partitions done by our own kernels over peer memory: row blocks are
    pushed into the peers' column slabs (regnn_rows_to_slabs), and the SpMM kernels store every finished row
    straight into its owner's row block from their epilogue (regnn_spmm_*_scatter) -- no collective library
    call, no pack / unpack pass.  The norm gradient is folded into the backward kernel per column slab (its
    backward is linear, so the shares add up at the R-float level): no row-dot pass and no N-float exchange.
    ``alias=True`` returns views of the exchange buffers (valid until the next step) instead of copies."""

    @staticmethod
    def forward(ctx, graph, etv, x_own, theta, alpha, norm, bounds, rank, xch, alias):
        csr = graph.csr()
        rb, re = bounds[rank], bounds[rank + 1]
        x_own = x_own.contiguous()
        ops.rows_to_slabs(x_own, xch.parts, rank * xch.per, xch.x_ptrs)
        xch.barrier()            # every rank's slice of X has landed in my slab
        weighted = theta is not None
        ops.spmm_scatter(csr, etv[0] if weighted else None, theta, alpha, norm, norm, xch.x_cols, xch.peer_y,
                         y_local=xch.y_cols if (norm is not None and weighted) else None)
        xch.barrier()            # every rank's columns of my rows have landed in my row block
        y_own = xch.y_rows[:re - rb]
        if not alias or not weighted:
            y_own = y_own.clone()
        ctx.graph, ctx.etv, ctx.alpha, ctx.bounds, ctx.rank, ctx.xch, ctx.alias = graph, etv, alpha, bounds, rank, xch, alias
        ctx.weighted = weighted
        if not weighted:   # REMixHop's copy_u propagation: no fused backward to fold the norm gradient into
            ctx.x_own, ctx.y_own = x_own.detach(), y_own
        ctx.save_for_backward(theta, norm)
        return y_own

    @staticmethod
    def backward(ctx, g_own):
        theta, norm = ctx.saved_tensors
        csr, xch, rank = ctx.graph.csr(), ctx.xch, ctx.rank
        rb, re = ctx.bounds[rank], ctx.bounds[rank + 1]
        g_own = g_own.contiguous()
        ops.rows_to_slabs(g_own, xch.parts, rank * xch.per, xch.g_ptrs)
        xch.barrier()
        if not ctx.weighted:   # the transposed SpMM alone, rows scattered to their owners; the owner's row dots are local
            ops.spmm_scatter(csr, None, None, ctx.alpha, norm, norm, xch.g_cols, xch.peer_dx, transposed=True)
            xch.barrier()
            dx_own = xch.dx_rows[:re - rb].clone()
            d_norm = _own_rows_norm_grad(norm, ctx.x_own, ctx.y_own, g_own, dx_own, rb, re) if norm is not None else None
            return None, None, dx_own, None, None, d_norm, None, None, None, None
        # d_norm: every row's norm gradient restricted to this rank's columns (all N rows, a share of the total)
        d_theta, d_norm = ops.spmm_bwd_fused_scatter(csr, ctx.etv[1], theta, ctx.alpha, norm, xch.x_cols, xch.g_cols,
                                                     xch.peer_dx, y_cols=xch.y_cols if norm is not None else None)
        xch.barrier()
        dx_own = xch.dx_rows[:re - rb]
        if not ctx.alias:
            dx_own = dx_own.clone()
        return None, None, dx_own, d_theta.view_as(theta), None, d_norm, None, None, None, None


def feature_sliced_propagate(graph, etv, x_own, theta, alpha, norm, bounds, rank, group=None, exchange=None,
                             alias=False):
    """Row block in, row block out; the aggregation runs on column slabs.  ``exchange`` (a ``SlabExchange``):
    re-partition over peer memory inside our own kernels; None: NCCL / gloo all-to-alls."""
    if exchange is not None:
        return _PeerSlabPropagate.apply(graph, etv, x_own, theta, alpha, norm, bounds, rank, exchange, alias)
    return _ColumnSlabPropagate.apply(graph, etv, x_own, theta, alpha, norm, bounds, rank, group)


class _HeadSlicedGat(torch.autograd.Function):
    """REGAT core (projection scores + fused logits / edge softmax / aggregation) sharded over the HEADS: attention
    heads are independent (el / er, the softmax statistics and the aggregation never mix heads), so rank q runs all
    rows for heads [q*H/P, (q+1)*H/P) -- the same column-slab scheme as the REGCN aggregation, with the row <-> slab
    re-partition done by one all-to-all on each side (NCCL; gloo in the CPU tests).  Row block in, row block out;
    parameter gradients (attn_l, attn_r, relation embedding) come back with this rank's head slice filled and zeros
    elsewhere: the caller sums them across ranks (``allreduce_relation_grads``)."""

    @staticmethod
    def forward(ctx, graph, etv, f_own, attn_l, attn_r, theta, alpha, slope, bounds, rank, group):
        csr = graph.csr()
        n = graph.number_of_nodes()
        parts, per = len(bounds) - 1, _uniform_rows(bounds)
        rows_own, heads, dim = f_own.shape
        if not per or heads % parts:
            raise ValueError('head-sliced attention needs equal row blocks and num_heads divisible by the number of ranks')
        hs = heads // parts
        sl = slice(rank * hs, (rank + 1) * hs)
        f_cols = _rows_to_columns(f_own.reshape(rows_own, heads * dim), per, parts, group)[:n].view(n, hs, dim)
        al, ar = attn_l[:, sl].contiguous(), attn_r[:, sl].contiguous()
        th = theta[:, sl].contiguous() if etv is not None else None
        et = etv[0] if etv is not None else None
        el, er = ops.attn_scores_fwd(f_cols, al, ar)
        out, rowmax, rowsum, _ = ops.gat_fwd(csr, et, th, alpha, f_cols, el, er, slope)
        ctx.graph, ctx.etv, ctx.alpha, ctx.slope, ctx.bounds, ctx.rank, ctx.group = graph, etv, alpha, slope, bounds, rank, group
        ctx.shape, ctx.sl = (rows_own, heads, dim), sl
        ctx.save_for_backward(f_cols, el, er, al, ar, th, out, rowmax, rowsum, attn_l, attn_r, theta)
        pad = torch.cat([out.view(n, hs * dim), out.new_zeros((parts * per - n, hs * dim))]) if parts * per > n else out.view(n, hs * dim)
        return _columns_to_rows(pad.contiguous(), per, parts, rows_own, group).view(rows_own, heads, dim)

    @staticmethod
    def backward(ctx, g_own):
        f_cols, el, er, al, ar, th, out, rowmax, rowsum, attn_l, attn_r, theta = ctx.saved_tensors
        csr = ctx.graph.csr()
        rows_own, heads, dim = ctx.shape
        n = ctx.graph.number_of_nodes()
        parts, per = len(ctx.bounds) - 1, _uniform_rows(ctx.bounds)
        hs = heads // parts
        g_cols = _rows_to_columns(g_own.reshape(rows_own, heads * dim).contiguous(), per, parts, ctx.group)[:n].view(n, hs, dim)
        et, et_t = (ctx.etv[0], ctx.etv[1]) if ctx.etv is not None else (None, None)
        d_f, _, _, d_th, d_al, d_ar = ops.gat_bwd(csr, et, et_t, th, ctx.alpha, f_cols, el, er, ctx.slope, None, out,
                                                   rowmax, rowsum, g_cols.contiguous(), attn_l=al, attn_r=ar)
        pad = torch.cat([d_f.view(n, hs * dim), d_f.new_zeros((parts * per - n, hs * dim))]) if parts * per > n else d_f.view(n, hs * dim)
        d_f_own = _columns_to_rows(pad.contiguous(), per, parts, rows_own, ctx.group).view(rows_own, heads, dim)
        g_al, g_ar = torch.zeros_like(attn_l), torch.zeros_like(attn_r)
        g_al[:, ctx.sl] = d_al.view(1, hs, dim)
        g_ar[:, ctx.sl] = d_ar.view(1, hs, dim)
        g_th = None
        if d_th is not None:
            g_th = torch.zeros_like(theta)
            g_th[:, ctx.sl] = d_th
        return None, None, d_f_own, g_al, g_ar, g_th, None, None, None, None, None


class _PeerHeadSlicedGat(torch.autograd.Function):
    """``_HeadSlicedGat`` with the four re-partitions done by our own kernels over peer memory (a ``SlabExchange`` of
    width H*D): row blocks are pushed into the peers' head slabs (regnn_rows_to_slabs), the finished head slab is pushed
    back to the row owners (regnn_slabs_to_rows) -- no collective library call, no pack / unpack pass.  The head slab of
    the features stays in the exchange buffer for the backward pass."""

    @staticmethod
    def forward(ctx, graph, etv, f_own, attn_l, attn_r, theta, alpha, slope, bounds, rank, xch):
        csr = graph.csr()
        n = graph.number_of_nodes()
        parts = xch.parts
        rows_own, heads, dim = f_own.shape
        if heads % parts or xch.feat != heads * dim:
            raise ValueError('peer head-sliced attention needs num_heads divisible by the number of ranks and a '
                             'SlabExchange of width num_heads * head_dim')
        hs = heads // parts
        sl = slice(rank * hs, (rank + 1) * hs)
        ops.rows_to_slabs(f_own.reshape(rows_own, heads * dim).contiguous(), parts, rank * xch.per, xch.x_ptrs)
        xch.barrier()            # every rank's rows of my heads have landed in my slab
        f_cols = xch.x_cols[:n].view(n, hs, dim)
        al, ar = attn_l[:, sl].contiguous(), attn_r[:, sl].contiguous()
        th = theta[:, sl].contiguous() if etv is not None else None
        et = etv[0] if etv is not None else None
        el, er = ops.attn_scores_fwd(f_cols, al, ar)
        out, rowmax, rowsum, _ = ops.gat_fwd(csr, et, th, alpha, f_cols, el, er, slope)
        ops.slabs_to_rows(out.view(n, hs * dim), n, xch.peer_y)
        xch.barrier()            # every rank's heads of my rows have landed in my row block
        ctx.graph, ctx.etv, ctx.alpha, ctx.slope, ctx.rank, ctx.xch = graph, etv, alpha, slope, rank, xch
        ctx.shape, ctx.sl = (rows_own, heads, dim), sl
        ctx.save_for_backward(el, er, al, ar, th, out, rowmax, rowsum, attn_l, attn_r, theta)
        return xch.y_rows[:rows_own].clone().view(rows_own, heads, dim)

    @staticmethod
    def backward(ctx, g_own):
        el, er, al, ar, th, out, rowmax, rowsum, attn_l, attn_r, theta = ctx.saved_tensors
        xch, rank = ctx.xch, ctx.rank
        csr = ctx.graph.csr()
        rows_own, heads, dim = ctx.shape
        n = ctx.graph.number_of_nodes()
        parts = xch.parts
        hs = heads // parts
        ops.rows_to_slabs(g_own.reshape(rows_own, heads * dim).contiguous(), parts, rank * xch.per, xch.g_ptrs)
        xch.barrier()
        f_cols = xch.x_cols[:n].view(n, hs, dim)
        g_cols = xch.g_cols[:n].view(n, hs, dim)
        et, et_t = (ctx.etv[0], ctx.etv[1]) if ctx.etv is not None else (None, None)
        d_f, _, _, d_th, d_al, d_ar = ops.gat_bwd(csr, et, et_t, th, ctx.alpha, f_cols, el, er, ctx.slope, None, out,
                                                   rowmax, rowsum, g_cols, attn_l=al, attn_r=ar)
        ops.slabs_to_rows(d_f.view(n, hs * dim), n, xch.peer_dx)
        xch.barrier()
        d_f_own = xch.dx_rows[:rows_own].clone().view(rows_own, heads, dim)
        g_al, g_ar = torch.zeros_like(attn_l), torch.zeros_like(attn_r)
        g_al[:, ctx.sl] = d_al.view(1, hs, dim)
        g_ar[:, ctx.sl] = d_ar.view(1, hs, dim)
        g_th = None
        if d_th is not None:
            g_th = torch.zeros_like(theta)
            g_th[:, ctx.sl] = d_th
        return None, None, d_f_own, g_al, g_ar, g_th, None, None, None, None, None


def head_sliced_gat(graph, etv, f_own, attn_l, attn_r, theta, alpha, slope, bounds, rank, group=None, exchange=None):
    """f_own [rows_own, H, D] -> out rows [rows_own, H, D]; see ``_HeadSlicedGat``.  ``exchange`` (a ``SlabExchange`` of
    width H*D): the re-partitions run over peer memory inside our own kernels instead of NCCL / gloo all-to-alls."""
    if exchange is not None:
        return _PeerHeadSlicedGat.apply(graph, etv, f_own, attn_l, attn_r, theta, alpha, slope, bounds, rank, exchange)
    return _HeadSlicedGat.apply(graph, etv, f_own, attn_l, attn_r, theta, alpha, slope, bounds, rank, group)


def allreduce_relation_grads(params, group=None, exchange=None):
    """Sums the per-rank relation-embedding gradients (R x H floats each) -- the only cross-rank floating-point
    reduction of the partitioned layer.  ``exchange`` (a ``SlabExchange``): all-gather over peer memory and a sum in
    rank order (bit-identical on every rank and topology); otherwise all-gather with the collective library (NCCL /
    gloo) followed by the same fixed-order sum."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    if exchange is not None and flat.numel() <= exchange.GRAD_SLOTS:
        total = exchange.allreduce_small(flat)
    else:
        with _lib.phase('allgather_relation_grads'):
            world = dist.get_world_size(group)
            parts = [torch.empty_like(flat) for _ in range(world)]
            dist.all_gather(parts, flat, group=group)
        total = parts[0].clone()
        for q in range(1, world):
            total = total + parts[q]
    off = 0
    for g in grads:
        g.copy_(total[off:off + g.numel()].view_as(g))
        off += g.numel()
