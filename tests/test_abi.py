"""CPU-side checks of the C-ABI boundary: libwfsp.so loads without a GPU and exports every symbol
include/wfsp.h declares; compute entries fail loudly (no CPU fallback)."""
import os
import re

import pytest
import torch

from waveformml_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "wfsp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wfsp_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libwfsp.so does not export %s" % n
        assert n in _lib.SIGNATURES, "no ctypes signature for %s" % n
    assert sorted(_lib.SIGNATURES) == names
    assert lib.wfsp_version() == _lib.EXPECTED_VERSION


def test_host_only_helpers():
    lib = _lib.load()
    out = (_lib.ctypes.c_int * 2)()
    _lib.check(lib.wfsp_conv_out_shape(_lib.ints([14, 11]), _lib.ints([3, 3]), _lib.ints([2, 2]), _lib.ints([0, 0]),
                                       _lib.ints([1, 1]), out))
    assert list(out) == [6, 5]
    assert lib.wfsp_rulebook_workspace_bytes(185, 64, _lib.ints([12, 9]), _lib.ints([3, 3])) > 64 * 108 * 4
    assert lib.wfsp_conv_apply_workspace_bytes(9, 100, 252, 158, _lib.MATH_BF16) >= 9 * 160 * 256 * 2 + 100 * 256 * 2
    assert lib.wfsp_conv_apply_workspace_bytes(9, 100, 252, 158, _lib.MATH_FP32) == 0


def test_bad_arguments_report_errors():
    lib = _lib.load()
    rc = lib.wfsp_rulebook_conv(None, 0, None, 1, _lib.ints([14, 11]), _lib.ints([3, 3]), _lib.ints([2, 2]),
                                _lib.ints([0, 0]), _lib.ints([2, 2]), None, 0, None, None, None, None, 0, None)
    assert rc == -1 and b"stride>1 with dilation>1" in lib.wfsp_last_error()
    assert lib.wfsp_set_option(b"no_such_option", 1) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from waveformml_b200 import spconv
    x = spconv.SparseConvTensor(torch.zeros(2, 4), torch.zeros(2, 3, dtype=torch.int32), [14, 11], 1)
    for layer in (spconv.SparseConv2d(4, 4, 3), spconv.SubMConv2d(4, 4, 3), spconv.SparseConv2d(4, 4, 1)):
        with pytest.raises(RuntimeError, match="no CPU path"):
            layer(x)
    with pytest.raises(RuntimeError, match="no CPU path"):
        x.dense()


def test_sparseconvnet_facade_builds_reference_config():
    """config/examples/OPs3ns_SCNet.json:27-66 names classes by string (`sparseconvnet.Convolution`, args as a list);
    build that algorithm list the way src/utils/util.py ModuleUtility does -- no GPU needed to construct."""
    import importlib
    algorithm = ["sparseconvnet.Convolution", [2, 300, 37, 1, 1, False], "sparseconvnet.Convolution", [2, 37, 37, 3, 1, False],
                 "sparseconvnet.Convolution", [2, 37, 18, 3, 2, False], "sparseconvnet.SparseToDense", [2, 18]]
    mods = []
    for name, args in zip(algorithm[0::2], algorithm[1::2]):
        mod, cls = name.rsplit(".", 1)
        mods.append(getattr(importlib.import_module(mod), cls)(*args))
    scn = importlib.import_module("sparseconvnet")
    net = scn.Sequential(*mods)
    assert len(net) == 4 and tuple(net[1].weight.shape) == (9, 37, 37) and net[2].stride == [2, 2]
    assert scn.InputLayer(2, [14, 11], mode=0).spatial_size == [14, 11]
