def to_undirected(*a, **k):   # imported by the reference's mag/regnn_layers.py, not called by the layers
    raise NotImplementedError('stub')
