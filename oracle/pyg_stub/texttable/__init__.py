class Texttable:   # import-only placeholder (mag/utils.py args_print)
    def __init__(self, *a, **k):
        raise NotImplementedError('stub')
