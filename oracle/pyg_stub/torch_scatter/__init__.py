"""torch-scatter primitives used by the reference's mag/utils.py (test infrastructure, see ../README.md)."""
import torch

from torch_geometric.nn import _scatter


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce='sum'):
    dim = dim if dim >= 0 else src.dim() + dim
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    return _scatter(src, index, dim, dim_size, reduce)


def segment_csr(src, indptr, out=None, reduce='sum'):
    counts = indptr[1:] - indptr[:-1]
    index = torch.repeat_interleave(torch.arange(counts.numel(), device=src.device), counts)
    return _scatter(src, index, 0, counts.numel(), reduce)


def gather_csr(src, indptr, out=None):
    counts = indptr[1:] - indptr[:-1]
    return torch.repeat_interleave(src, counts, dim=0)
