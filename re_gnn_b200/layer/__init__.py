"""Drop-in replacements for the reference's ``layer`` package (layer/__init__.py:1-6): register this
package under the module name ``layer`` and the reference's model/REGCN.py, model/REGAT.py and
model/REMixHop.py import and run unchanged."""
from .conv import (REGraphConv, REGATConv, REGATv2Conv, REMixHopConv,  # noqa: F401
                   RESAGEConv, REGINConv)
