"""GPU neighbour sampler for the sampled-minibatch path (BASELINE config 5).

Stands in for ``NeighborSampler(edge_index, node_idx=..., sizes=[25, 20], batch_size=...)`` of
mag/regnn_ns.py:206-214: layer-wise sampling from the seed nodes outwards, up to ``size`` distinct in-edges
per target without replacement, target nodes first in every frontier, blocks returned outermost-first like
PyG's ``adjs``.  The random choice is counter-based (``regnn_sample_neighbors``): a batch is a pure function
of (seed, epoch, rank, batch index), so data-parallel ranks draw independent batches and the CPU oracle
(oracle/sampler_oracle.py) reproduces every sampled edge set bit for bit.  Relabelling uses torch
sort/unique on the device (plumbing).
"""
import ctypes

import torch

from . import _lib
from .graph import Graph

M32 = 0xFFFFFFFF


def _mix32(z):
    z &= M32
    z ^= z >> 16
    z = (z * 0x7feb352d) & M32
    z ^= z >> 15
    z = (z * 0x846ca68b) & M32
    z ^= z >> 16
    return z


def layer_key(seed, epoch, rank, batch, layer):
    lo = _mix32((seed * 0x9E3779B1 + epoch * 0x85EBCA6B + layer * 0xC2B2AE35) & M32)
    hi = _mix32((rank * 0x27D4EB2F + batch * 0x165667B1 + 0x5bd1e995) & M32)
    return (hi << 32) | lo


class Block:
    """One bipartite message-passing block: edges ``src_local -> dst_local`` between the ``n_src`` nodes of
    the outer frontier and its first ``n_dst`` nodes (the targets); ``eid`` are edge ids of the full graph."""

    def __init__(self, src_local, dst_local, eid, n_src, n_dst):
        self.src, self.dst, self.eid, self.n_src, self.n_dst = src_local, dst_local, eid, int(n_src), int(n_dst)

    @property
    def edge_index(self):
        return torch.stack([self.src, self.dst])

    @property
    def size(self):
        return (self.n_src, self.n_dst)


class NeighborSampler:
    def __init__(self, graph: Graph, fanouts, seed=0, rank=0):
        self.graph, self.fanouts, self.seed, self.rank = graph, list(fanouts), int(seed), int(rank)

    def sample_slots(self, targets, fanout, key):
        csr = self.graph.csr()
        t = targets.numel()
        out = torch.empty(max(t * fanout, 1), dtype=torch.int32, device=targets.device)[:t * fanout]
        with torch.cuda.device(targets.device):
            _lib.call('regnn_sample_neighbors', ctypes.c_void_p(csr['indptr'].data_ptr()),
                      ctypes.c_void_p(targets.data_ptr()), t, int(fanout), ctypes.c_uint64(key),
                      ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            _lib.count_launches(1)
        return out.view(t, fanout)

    def sample(self, seeds, epoch=0, batch=0):
        """-> (n_id [int64, global ids of the outermost frontier], [Block, ...] outermost first)."""
        csr = self.graph.csr()
        n_id = seeds.to(torch.int64).contiguous()
        blocks = []
        for layer, fanout in enumerate(self.fanouts):
            slots = self.sample_slots(n_id, fanout, layer_key(self.seed, epoch, self.rank, batch, layer))
            valid = slots >= 0
            t = n_id.numel()
            dst_local = torch.arange(t, device=n_id.device).view(-1, 1).expand_as(slots)[valid]
            flat = slots[valid].to(torch.int64)
            src_global = csr['indices'][flat].to(torch.int64)
            eid = csr['eid'][flat].to(torch.int64)
            uniq, inv = torch.unique(torch.cat([n_id, src_global]), return_inverse=True)
            is_target = torch.zeros(uniq.numel(), dtype=torch.bool, device=n_id.device)
            is_target[inv[:t]] = True
            rest = uniq[~is_target]
            pos = torch.empty(uniq.numel(), dtype=torch.int64, device=n_id.device)
            pos[inv[:t]] = torch.arange(t, device=n_id.device)
            pos[~is_target] = t + torch.arange(rest.numel(), device=n_id.device)
            new_n_id = torch.cat([n_id, rest])
            blocks.append(Block(pos[inv[t:]], dst_local, eid, new_n_id.numel(), t))
            n_id = new_n_id
        return n_id, blocks[::-1]


class SaintRandomWalkSampler:
    """GraphSAINT random-walk sampler (``GraphSAINTRandomWalkSampler(data, batch_size=roots, walk_length=...)``,
    mag/regnn_saint.py:185-190; ``sample_coverage=0`` so no normalisation statistics): ``roots`` random start
    nodes, ``walk_length`` steps along out-edges (``regnn_random_walk``), then the subgraph INDUCED by the visited
    nodes.  Returns ``(n_id, edge_index_local [2,E'], eid)`` with ``n_id`` sorted ascending."""

    def __init__(self, graph: Graph, roots, walk_length, seed=0, rank=0):
        self.graph, self.roots, self.walk_length = graph, int(roots), int(walk_length)
        self.seed, self.rank = int(seed), int(rank)

    def walks(self, key):
        csr = self.graph.csr()
        dev = csr['indptr'].device
        out = torch.empty((self.roots, self.walk_length + 1), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _lib.call('regnn_random_walk', ctypes.c_void_p(csr['indptr_t'].data_ptr()),
                      ctypes.c_void_p(csr['indices_t'].data_ptr()), self.graph.number_of_nodes(), self.roots,
                      self.walk_length, ctypes.c_uint64(key), ctypes.c_void_p(out.data_ptr()),
                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            _lib.count_launches(1)
        return out

    def sample(self, epoch=0, batch=0):
        csr = self.graph.csr()
        n = self.graph.number_of_nodes()
        dev = csr['indptr'].device
        n_id = torch.unique(self.walks(layer_key(self.seed, epoch, self.rank, batch, 0x5A1)).view(-1))
        relabel = torch.full((n,), -1, dtype=torch.int64, device=dev)
        relabel[n_id] = torch.arange(n_id.numel(), device=dev)
        indptr = csr['indptr'].to(torch.int64)
        starts, degs = indptr[n_id], indptr[n_id + 1] - indptr[n_id]
        total = int(degs.sum().item())
        first = torch.cumsum(degs, 0) - degs
        slot = torch.repeat_interleave(starts - first, degs) + torch.arange(total, device=dev)
        dst_local = torch.repeat_interleave(torch.arange(n_id.numel(), device=dev), degs)
        src_local = relabel[csr['indices'][slot].to(torch.int64)]
        keep = src_local >= 0
        return n_id, torch.stack([src_local[keep], dst_local[keep]]), csr['eid'][slot[keep]].to(torch.int64)
