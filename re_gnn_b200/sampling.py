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
