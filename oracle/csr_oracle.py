"""CPU ORACLE (numpy) for the integer side of the path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Bit-exact definition of the graph structures the CUDA library builds (``regnn_csr_build`` and
``regnn_etype_permute`` in include/regnn_b200.h).  The reference leaves this to DGL: a
``DGLGraph`` lazily builds an in-CSR (destination-indexed) for ``update_all`` and an out-CSR for the
backward pass (call sites layer/REGraphConv.py:69-92).  The conventions fixed
here -- and checked against the reference's edge-id alignment ``for u, v in zip(*g.edges())``
(run_regnn.py:94-99) -- are:

  * edge k is ``src[k] -> dst[k]``; ``etype[k]`` is 1-based and aligned to edge ids;
  * slots of the dst-sorted CSR are ordered by (dst, edge id)  -- a *stable* sort by dst;
  * entries of the transposed (src-sorted) view are ordered by (src, CSR slot).
"""
import numpy as np


def csr_build(src, dst, num_nodes):
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    eid = np.argsort(dst, kind='stable').astype(np.int32)
    indices = src[eid].astype(np.int32)
    row = dst[eid].astype(np.int32)
    indptr = np.zeros(num_nodes + 1, dtype=np.int32)
    np.cumsum(np.bincount(dst, minlength=num_nodes), out=indptr[1:])
    slot_t = np.argsort(indices, kind='stable').astype(np.int32)
    indices_t = row[slot_t]
    indptr_t = np.zeros(num_nodes + 1, dtype=np.int32)
    np.cumsum(np.bincount(src, minlength=num_nodes), out=indptr_t[1:])
    return dict(indptr=indptr, indices=indices, eid=eid, row=row,
                indptr_t=indptr_t, indices_t=indices_t, slot_t=slot_t)


def etype_permute(etype_1based, eid, slot_t):
    """uint8 0-based edge types in CSR-slot order and in transposed-entry order."""
    et = (np.asarray(etype_1based, dtype=np.int64) - 1).astype(np.uint8)
    et_csr = et[eid]
    return et_csr, et_csr[slot_t]


def in_degrees(dst, num_nodes):
    return np.bincount(np.asarray(dst, dtype=np.int64), minlength=num_nodes).astype(np.int64)
