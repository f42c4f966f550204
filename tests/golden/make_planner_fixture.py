"""Generates tests/golden/planner_stacks.json by running the REFERENCE's own planners
(/root/reference/src/models/SPConvBlocks.py, src/utils/ModelValidation.py) against the drop-in
`spconv` package of this repo.  Run in the build container only (the reference is not on the GPU box):

    python tests/golden/make_planner_fixture.py

The fixture pins (a) that the reference's callers construct cleanly against our layer signatures
and (b) the exact layer stacks that waveformml_b200/stacks.py must reproduce.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import spconv  # noqa: E402  (this repo's shim)
from waveformml_b200.stacks import describe  # noqa: E402
from src.models.SPConvBlocks import (SparseConv2DBlock, SparseConv2DForEZ, SparseConv2DForZ,  # noqa: E402
                                     SparseConv2DPreserve, Pointwise2DForZ)
from src.utils.ModelValidation import ModelValidation  # noqa: E402

assert spconv.__file__.startswith(ROOT)

out = {}
gep = json.load(open(os.path.join(REF, "config/examples/GEP.json")))
blk = SparseConv2DBlock(300, gep["net_config"]["hparams"]["out_planes"], gep["net_config"]["hparams"]["n_conv"],
                        [14, 11, 300], True, **gep["net_config"]["hparams"]["conv_params"])
out["GEP"] = {"layers": describe(blk.func), "out_size": blk.out_size}

z = json.load(open(os.path.join(REF, "config/examples/SingleEndedZCNN.json")))
out["SingleEndedZCNN"] = {"layers": describe(SparseConv2DForZ(300, **z["net_config"]["hparams"]["conv"]).network)}

out["ForEZ_v2_k5"] = {"layers": describe(SparseConv2DForEZ(300, out_planes=2, kernel_size=5, n_conv=2, n_point=3,
                                                            conv_position=2, version=2).network)}
out["ForEZ_v0_default"] = {"layers": describe(SparseConv2DForEZ(300).network)}
out["Pointwise2DForZ"] = {"layers": describe(Pointwise2DForZ(300, 2).network)}

ioni = json.load(open(os.path.join(REF, "config/examples/IoniClassifierCNN.json")))
out["IoniClassifierCNN"] = {"layers": describe(SparseConv2DPreserve(130, 5, ioni["net_config"]["hparams"]["n_conv"],
                                                                    **ioni["net_config"]["hparams"]["conv_params"]).func)}
out["Preserve_v2_extreme"] = {"layers": describe(SparseConv2DPreserve(300, 3, 5, version=2, size_factor=3,
                                                                      filter_multiplier=1.5, n_contraction=5).func)}

# output-shape contract: the reference's own size calculator for a sweep of conv geometries
sizes = []
for k in (1, 2, 3, 5):
    for s in (1, 2, 3):
        for p in (0, 1, 2):
            for d in (1, 2):
                arg = {"DIMENSION": 2, "N_INPUT_CHANNELS": 4, "N_OUTPUT_CHANNELS": 7, "FILTER_SIZE": [k] * 4,
                       "FILTER_STRIDE": [s] * 4, "FILTER_PADDING": [p] * 4, "FILTER_DILATION": [d] * 4}
                sizes.append({"k": k, "s": s, "p": p, "d": d,
                              "out": ModelValidation.calc_output_size(arg, [14, 11, 4], "cur", "prev", 2)})
out["calc_output_size"] = sizes

with open(os.path.join(ROOT, "tests/golden/planner_stacks.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print("wrote planner_stacks.json:", {k: len(v.get("layers", v)) if isinstance(v, dict) else len(v) for k, v in out.items()})
