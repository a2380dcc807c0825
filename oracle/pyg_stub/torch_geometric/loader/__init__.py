class NeighborSampler:   # import-only placeholder
    def __init__(self, *a, **k):
        raise NotImplementedError('stub')
