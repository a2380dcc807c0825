"""PyG-semantics stub (test infrastructure): see ../README.md."""
__version__ = '2.0.0-stub'
