"""CPU ORACLE (numpy) for the neighbour sampler -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Bit-exact specification of ``regnn_sample_neighbors`` + the block relabelling in
``re_gnn_b200/sampling.py``.  The reference samples with PyG's ``NeighborSampler`` /
``torch_sparse.sample_adj`` (mag/regnn_ns.py:206-214), whose RNG stream is neither pinned nor available
here (torch_sparse is not installable), so the random choice itself is OUR specification; what is kept
from the reference is the contract: up to ``size`` distinct in-neighbours per target without replacement
(all of them when the in-degree is smaller), targets first in the next frontier, layers sampled from the
seeds outwards and consumed outermost-first.
"""
import numpy as np

M32 = 0xFFFFFFFF


def mix32(z):
    z &= M32
    z ^= z >> 16
    z = (z * 0x7feb352d) & M32
    z ^= z >> 15
    z = (z * 0x846ca68b) & M32
    z ^= z >> 16
    return z


def feistel_perm(x, deg, key):
    b = 2
    while (1 << b) < deg:
        b += 2
    h = b >> 1
    mask = (1 << h) - 1
    y = x
    while True:
        L, R = y >> h, y & mask
        for r in range(4):
            f = mix32((R * 0x9E3779B1 + key + r * 0x85EBCA6B) & M32) & mask
            L, R = R, L ^ f
        y = (L << h) | R
        if y < deg:
            return y


def sample_slots(indptr, targets, fanout, key):
    """-> int32 [T, fanout] CSR slots (-1 = none)."""
    key_lo, key_hi = key & M32, (key >> 32) & M32
    out = np.full((len(targets), fanout), -1, dtype=np.int32)
    for i, t in enumerate(np.asarray(targets).tolist()):
        s0, deg = int(indptr[t]), int(indptr[t + 1] - indptr[t])
        if deg <= fanout:
            out[i, :deg] = s0 + np.arange(deg)
        else:
            k = mix32(key_lo ^ mix32((t + key_hi) & M32))
            out[i] = [s0 + feistel_perm(j, deg, k) for j in range(fanout)]
    return out


def layer_key(seed, epoch, rank, batch, layer):
    """64-bit counter-based key; must match re_gnn_b200.sampling.layer_key."""
    lo = mix32((seed * 0x9E3779B1 + epoch * 0x85EBCA6B + layer * 0xC2B2AE35) & M32)
    hi = mix32((rank * 0x27D4EB2F + batch * 0x165667B1 + 0x5bd1e995) & M32)
    return (hi << 32) | lo


def sample_blocks(csr, seeds, fanouts, seed=0, epoch=0, rank=0, batch=0):
    """Layer-wise sampling from the seeds outwards.  Returns (n_id, blocks) with blocks ordered OUTERMOST
    FIRST like PyG's ``adjs``; each block is (src_local, dst_local, eid, n_src, n_dst)."""
    indptr, indices, eid = (np.asarray(csr[k]) for k in ('indptr', 'indices', 'eid'))
    n_id = np.asarray(seeds, dtype=np.int64)
    blocks = []
    for layer, fanout in enumerate(fanouts):
        slots = sample_slots(indptr, n_id, fanout, layer_key(seed, epoch, rank, batch, layer))
        valid = slots >= 0
        dst_local = np.repeat(np.arange(len(n_id), dtype=np.int64), fanout).reshape(slots.shape)[valid]
        flat = slots[valid].astype(np.int64)
        src_global = indices[flat].astype(np.int64)
        e = eid[flat].astype(np.int64)
        t = len(n_id)
        uniq, inv = np.unique(np.concatenate([n_id, src_global]), return_inverse=True)
        is_target = np.zeros(len(uniq), dtype=bool)
        is_target[inv[:t]] = True
        rest = uniq[~is_target]
        pos = np.empty(len(uniq), dtype=np.int64)
        pos[inv[:t]] = np.arange(t)
        pos[~is_target] = t + np.arange(len(rest))
        src_local = pos[inv[t:]]
        new_n_id = np.concatenate([n_id, rest])
        blocks.append((src_local, dst_local, e, len(new_n_id), t))
        n_id = new_n_id
    return n_id, blocks[::-1]


# ---- GraphSAINT random-walk sampler (mag/regnn_saint.py:185-190) ---------------------------------------------
def random_walks(csr, num_nodes, num_roots, walk_length, key):
    """int64 [num_roots, walk_length+1]; bit-exact spec of regnn_random_walk."""
    key_lo, key_hi = key & M32, (key >> 32) & M32
    indptr_t, indices_t = np.asarray(csr['indptr_t']), np.asarray(csr['indices_t'])
    out = np.zeros((num_roots, walk_length + 1), dtype=np.int64)
    for i in range(num_roots):
        cur = mix32(key_lo ^ mix32((i * 0x9E3779B1 + key_hi) & M32)) % num_nodes
        out[i, 0] = cur
        for step in range(1, walk_length + 1):
            t0, deg = int(indptr_t[cur]), int(indptr_t[cur + 1] - indptr_t[cur])
            if deg > 0:
                r = mix32(key_hi ^ mix32((i * 0x85EBCA6B + step * 0xC2B2AE35 + key_lo) & M32))
                cur = int(indices_t[t0 + r % deg])
            out[i, step] = cur
    return out


def saint_subgraph(csr, num_nodes, num_roots, walk_length, seed=0, epoch=0, rank=0, batch=0):
    """Node-induced subgraph of the nodes visited by the walks: (n_id sorted ascending, src_local, dst_local, eid),
    edges ordered by (destination position in n_id, CSR slot)."""
    walks = random_walks(csr, num_nodes, num_roots, walk_length, layer_key(seed, epoch, rank, batch, 0x5A1))
    n_id = np.unique(walks.reshape(-1))
    indptr, indices, eid = (np.asarray(csr[k]) for k in ('indptr', 'indices', 'eid'))
    relabel = np.full(num_nodes, -1, dtype=np.int64)
    relabel[n_id] = np.arange(len(n_id))
    src_l, dst_l, es = [], [], []
    for j, v in enumerate(n_id.tolist()):
        sl = np.arange(indptr[v], indptr[v + 1])
        keep = relabel[indices[sl]] >= 0
        src_l.append(relabel[indices[sl][keep]])
        dst_l.append(np.full(int(keep.sum()), j, dtype=np.int64))
        es.append(eid[sl][keep].astype(np.int64))
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, dtype=np.int64)  # noqa: E731
    return n_id, cat(src_l), cat(dst_l), cat(es)
