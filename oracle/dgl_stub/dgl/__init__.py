"""TEST INFRASTRUCTURE ONLY -- a pure-PyTorch stand-in for the slice of DGL 0.7.1 that
the reference's ``layer/*.py`` and ``model/*.py`` touch.

Why it exists: ``dgl==0.7.1`` (reference ``requirements.txt:6``) is not installable in this
image, so the reference layers cannot be imported as they are.  With this package first on
``sys.path`` under the name ``dgl`` the reference's *unmodified* layer and model files import
and run on CPU, which lets ``tests/golden/make_golden.py`` record golden input/output/gradient
vectors from the reference's own code.  Only the DGL primitives are restated here (semantics:
SURVEY.md Appendix B); every line of layer logic that is exercised is the reference's.

Nothing in ``re_gnn_b200`` imports this package.  Only ``tests/`` and the golden generator do.
"""
import contextlib

import numpy as np
import torch

from . import function  # noqa: F401
from .base import DGLError  # noqa: F401


class _Frame(dict):
    def update(self, other=(), **kw):  # DGL frames accept dict updates
        super().update(other, **kw)


class DGLGraph:
    """Homogeneous graph: edge k is ``src[k] -> dst[k]`` (edge-id order preserved).

    ``DGLGraph(scipy_sparse)``: a stored entry ``A[i, j]`` is the edge ``i -> j`` (row = source,
    column = destination) in the matrix's COO order -- the convention ``run_regnn.py:84`` relies on.
    """

    is_block = False

    def __init__(self, data=None, num_nodes=None):
        if data is None:
            src = dst = torch.zeros(0, dtype=torch.int64)
            n = int(num_nodes or 0)
        elif isinstance(data, tuple):
            src = torch.as_tensor(np.asarray(data[0]), dtype=torch.int64)
            dst = torch.as_tensor(np.asarray(data[1]), dtype=torch.int64)
            n = int(num_nodes) if num_nodes is not None else int(max(src.max(), dst.max())) + 1
        else:  # scipy sparse matrix
            coo = data.tocoo()
            src = torch.as_tensor(coo.row.astype(np.int64))
            dst = torch.as_tensor(coo.col.astype(np.int64))
            n = int(coo.shape[0])
        self._src, self._dst, self._n = src, dst, n
        self.ndata, self.edata = _Frame(), _Frame()

    # homogeneous graph: source and destination frames are the node frame
    @property
    def srcdata(self):
        return self.ndata

    @property
    def dstdata(self):
        return self.ndata

    def edges(self):
        return self._src, self._dst

    def number_of_nodes(self):
        return self._n

    num_nodes = number_of_nodes
    number_of_dst_nodes = number_of_nodes
    number_of_src_nodes = number_of_nodes

    def number_of_edges(self):
        return int(self._src.numel())

    num_edges = number_of_edges

    def in_degrees(self):
        return torch.bincount(self._dst, minlength=self._n)

    def to(self, device):
        return self

    def local_var(self):
        g = DGLGraph.__new__(DGLGraph)
        g._src, g._dst, g._n = self._src, self._dst, self._n
        g.ndata, g.edata = _Frame(self.ndata), _Frame(self.edata)
        return g

    @contextlib.contextmanager
    def local_scope(self):
        nd, ed = self.ndata, self.edata
        self.ndata, self.edata = _Frame(nd), _Frame(ed)
        try:
            yield
        finally:
            self.ndata, self.edata = nd, ed

    # ---- message passing ------------------------------------------------------------------
    def _message(self, m):
        if m.kind == 'u_mul_e':
            return self.ndata[m.lhs][self._src] * self.edata[m.rhs]
        if m.kind == 'copy_u':
            return self.ndata[m.lhs][self._src]
        if m.kind == 'u_add_v':
            return self.ndata[m.lhs][self._src] + self.ndata[m.rhs][self._dst]
        raise NotImplementedError(m.kind)

    def update_all(self, message_func, reduce_func):
        msg = self._message(message_func)
        if reduce_func.kind != 'sum':
            raise NotImplementedError(reduce_func.kind)
        out = torch.zeros((self._n,) + tuple(msg.shape[1:]), dtype=msg.dtype)
        self.ndata[reduce_func.out] = out.index_add(0, self._dst, msg)

    def apply_edges(self, func):
        self.edata[func.out] = self._message(func)


def remove_self_loop(g):
    keep = g._src != g._dst
    return DGLGraph((g._src[keep], g._dst[keep]), num_nodes=g._n)


def add_self_loop(g):
    loop = torch.arange(g._n, dtype=torch.int64)
    return DGLGraph((torch.cat([g._src, loop]), torch.cat([g._dst, loop])), num_nodes=g._n)


def graph(data, num_nodes=None):
    return DGLGraph(data, num_nodes=num_nodes)
