from ..functional import edge_softmax  # noqa: F401
