from torch.nn import Identity  # noqa: F401
