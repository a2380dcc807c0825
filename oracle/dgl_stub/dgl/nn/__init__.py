from ..utils import expand_as_pair  # noqa: F401  (layer/REGINConv.py:5 imports it from here)
