class PygNodePropPredDataset:   # import-only placeholders
    def __init__(self, *a, **k):
        raise NotImplementedError('stub')


class Evaluator:
    def __init__(self, *a, **k):
        raise NotImplementedError('stub')
