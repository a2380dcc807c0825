# model/REMixHop.py:14 imports these names and never uses them.
CiteseerGraphDataset = CoraGraphDataset = PubmedGraphDataset = None
