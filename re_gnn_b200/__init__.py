"""re_gnn_b200 -- B200-native (sm_100a) relation-embedded message passing for RE-GNN.

Drop-in ``torch.nn.Module`` replacements for the reference's ``layer.REGraphConv``,
``layer.REGATConv``, ``layer.REGATv2Conv`` and ``layer.REMixHopConv`` (same constructors, parameter
names and ``forward(g, feat, etype)``), a ``Graph`` stand-in for the DGLGraph argument, and the
C-ABI CUDA library underneath (``include/regnn_b200.h``).  No DGL, no Triton, no CPU fallback.
"""
from .graph import Graph, ZeroInDegreeError  # noqa: F401
from .layer import (REGraphConv, REGATConv, REGATv2Conv, REMixHopConv,  # noqa: F401
                    RESAGEConv, REGINConv)

__version__ = '0.1.0'
