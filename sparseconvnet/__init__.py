"""Top-level alias so that `import sparseconvnet as scn` (src/models/SCNet.py:3,
config/examples/OPs3ns_SCNet.json:23) resolves to the B200 implementation."""
from waveformml_b200.sparseconvnet import *  # noqa: F401,F403
from waveformml_b200.sparseconvnet import __all__  # noqa: F401
