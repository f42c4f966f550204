"""The reference's planners (src/models/SPConvBlocks.py) were run against this repo's `spconv`
drop-in by tests/golden/make_planner_fixture.py; the fixture pins the layer stacks and the
output-shape contract (src/utils/ModelValidation.py:119-177).  No GPU needed: construction only."""
import json
import os

import pytest

from waveformml_b200 import spconv, stacks

FIX = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "planner_stacks.json")))


def test_gep_stack_matches_reference_planner():
    m = stacks.PSDClassifier()
    assert stacks.describe(m.sparseModel) == FIX["GEP"]["layers"]
    assert m.out_size == FIX["GEP"]["out_size"] and m.n_linear == 4480
    assert [(l.in_features, l.out_features) for l in m.linear] == [(4480, 116), (116, 3)]
    assert sum(p.numel() for p in m.parameters()) == 1046047
    assert tuple(m.state_dict()["sparseModel.3.weight"].shape) == (3, 3, 252, 158)


def test_z_stack_matches_reference_planner():
    assert stacks.describe(stacks.ZRegressor().model.network) == FIX["SingleEndedZCNN"]["layers"]


def test_ez_stack_matches_reference_planner():
    assert stacks.describe(stacks.EZSubM().network) == FIX["ForEZ_v2_k5"]["layers"]


def test_ioni_stack_matches_reference_planner():
    assert stacks.describe(stacks.IoniPreserve().model.func) == FIX["IoniClassifierCNN"]["layers"]


def test_output_shape_contract():
    for row in FIX["calc_output_size"]:
        k, s, p, d = row["k"], row["s"], row["p"], row["d"]
        got = spconv.ops.get_conv_output_size([14, 11], [k, k], [s, s], [p, p], [d, d])
        # the reference's calculator uses true division then int(): identical for non-negative sizes
        if min(row["out"][:2]) >= 1:
            assert got == row["out"][:2], row


def test_unused_layer_types_raise():
    for name in ("SparseConv4d", "SparseConvTranspose2d", "SparseConvTranspose3d", "SparseConv1d"):
        with pytest.raises(NotImplementedError):
            getattr(spconv, name)(4, 4, 3)


def test_3d_layer_types_construct():
    """src/utils/ModelValidation.py:24-31 lists SparseConv3d / SubMConv3d; net_type "3DConvolution"
    (src/models/SPConvNet.py:42-49) would build them on the [14, 11, n_samples] grid."""
    c = spconv.SparseConv3d(4, 6, 3, 2, 1, 1, 1, False)
    assert tuple(c.weight.shape) == (3, 3, 3, 4, 6) and c.stride == [2, 2, 2] and c.bias is None and c.ndim == 3
    s = spconv.SubMConv3d(4, 6, [3, 3, 5], indice_key="subm0")
    assert tuple(s.weight.shape) == (3, 3, 5, 4, 6) and s.subm and s.indice_key == "subm0"
    assert spconv.SparseConv3d(4, 4, 1).conv1x1


def test_import_as_spconv():
    import spconv as top
    assert top.SparseConv2d is spconv.SparseConv2d and top.ops.get_conv_output_size([14], [3], [1], [0], [1]) == [12]


def test_layer_signature_and_attributes():
    c = spconv.SparseConv2d(8, 4, 3, 1, 0, 1, 1, False)  # the 8 positionals the reference passes
    assert c.bias is None and tuple(c.weight.shape) == (3, 3, 8, 4)
    assert (c.in_channels, c.out_channels, c.kernel_size) == (8, 4, [3, 3])
    s = spconv.SubMConv2d(8, 4, 3, 2, 5, indice_key="subm0")  # stride / padding accepted and ignored
    assert s.subm and s.indice_key == "subm0" and s.bias is not None
    seq = spconv.SparseSequential(c, s)
    assert len(seq) == 2 and seq[0] is c and seq[-1] is s
    with pytest.raises(AssertionError):
        spconv.SparseConv2d(8, 4, 3, groups=2)
