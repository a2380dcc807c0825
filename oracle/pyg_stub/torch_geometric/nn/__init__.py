"""``MessagePassing`` restated from the PyG 2.0 documentation (test infrastructure, see ../../README.md)."""
import inspect

import torch


def _scatter(src, index, dim, dim_size, reduce):
    shape = list(src.shape)
    shape[dim] = dim_size
    idx = index.view([-1 if d == dim else 1 for d in range(src.dim())]).expand_as(src)
    if reduce in ('add', 'sum', 'mean'):
        out = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add(dim, idx, src)
        if reduce == 'mean':
            cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device).scatter_add(
                0, index, torch.ones(index.numel(), dtype=src.dtype, device=src.device)).clamp(min=1)
            out = out / cnt.view([-1 if d == dim else 1 for d in range(src.dim())])
        return out
    if reduce == 'max':
        out = torch.full(shape, float('-inf'), dtype=src.dtype, device=src.device)
        out = out.scatter_reduce(dim, idx, src, 'amax', include_self=True)
        return torch.where(torch.isinf(out), torch.zeros_like(out), out)
    raise ValueError(reduce)


class MessagePassing(torch.nn.Module):
    """flow = source_to_target: ``edge_index[0]`` holds the sources j, ``edge_index[1]`` the targets i."""

    def __init__(self, aggr='add', flow='source_to_target', node_dim=-2):
        super().__init__()
        assert flow == 'source_to_target'
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        j, i = edge_index[0], edge_index[1]
        dim = self.node_dim
        dim_size = None if size is None else size[1]
        args = {}
        for name in inspect.signature(self.message).parameters:
            if name.endswith('_j') or name.endswith('_i'):
                data = kwargs[name[:-2]]
                side = 0 if name.endswith('_j') else 1
                if isinstance(data, (tuple, list)):
                    if dim_size is None and data[1] is not None:
                        dim_size = data[1].size(dim)
                    data = data[side]
                elif dim_size is None:
                    dim_size = data.size(dim)
                args[name] = data.index_select(dim, j if side == 0 else i)
            else:
                args[name] = kwargs[name]
        msg = self.message(**args)
        if dim_size is None:
            dim_size = int(i.max()) + 1 if i.numel() else 0
        out = self.aggregate(msg, i, dim_size=dim_size)
        return self.update(out)

    def message(self, x_j):
        return x_j

    def aggregate(self, inputs, index, dim_size=None):
        dim = self.node_dim if self.node_dim >= 0 else inputs.dim() + self.node_dim
        return _scatter(inputs, index, dim, dim_size, self.aggr)

    def update(self, aggr_out):
        return aggr_out


class GATv2Conv(MessagePassing):   # imported by the reference, never instantiated by the layers under test
    def __init__(self, *a, **k):
        raise NotImplementedError('stub')
