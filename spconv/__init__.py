"""`import spconv` shim: the reference imports the layer package by this name
(src/models/SPConvBlocks.py:4, config/examples/GEP.json:24-30).  Everything lives in
waveformml_b200.spconv."""
import sys as _sys

from waveformml_b200.spconv import *  # noqa: F401,F403
from waveformml_b200.spconv import (SparseConv1d, SparseConv3d, SparseConv4d, SparseConvTranspose2d,  # noqa: F401
                                    SparseConvTranspose3d, SubMConv3d, functional, is_spconv_module, ops)

_sys.modules[__name__ + ".ops"] = ops
_sys.modules[__name__ + ".functional"] = functional
