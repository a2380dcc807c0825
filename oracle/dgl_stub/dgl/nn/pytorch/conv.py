# model/GCN.py:3 and model/GAT.py:3 import these stock DGL layers at package-import time.
# They are out of scope -- the names only have to exist for ``import model`` to work.
class GraphConv:  # pragma: no cover
    def __init__(self, *a, **k):
        raise NotImplementedError('stock dgl.nn.GraphConv is outside the RE-GNN hot path')


class GATConv:  # pragma: no cover
    def __init__(self, *a, **k):
        raise NotImplementedError('stock dgl.nn.GATConv is outside the RE-GNN hot path')
