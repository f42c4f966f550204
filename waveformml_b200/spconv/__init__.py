"""Drop-in for the `spconv` (v1.2.x) layer API that WaveformML's models construct
(src/models/SPConvBlocks.py, SPConvNet.py:63-64, SingleEndedZConv.py:41-45, LitBase.py:138-146),
backed by the sm_100a kernels in libwfsp.so.  Constructing modules needs no GPU; running them needs
CUDA tensors -- there is no CPU path.

Kept from upstream: constructor signatures incl. the 8 positional arguments the
reference passes (nin, nout, kernel, stride, padding, dilation, groups, bias); weight shape
[kH, kW, Cin, Cout] with kaiming_uniform_(a=sqrt(5)) init, bias [Cout]; kernel volume 1 takes the
`features @ weight` shortcut and ignores stride / padding; SubMConv2d ignores stride / padding
(pad := k//2, stride := 1); rulebooks are cached in the tensor's shared `indice_dict` under
`indice_key` (also under None); SparseInverseConv2d reuses its partner's rulebook with the roles
swapped; bias is added to active rows only; SparseSequential applies non-sparse modules to
`.features` when there is at least one row.
"""
import math
import os
from collections import OrderedDict

import numpy as np
import torch
from torch import nn

from . import functional as Fsp
from . import fused, ops
from .ops import Rulebook, get_conv_output_size, get_indice_pairs  # noqa: F401

__all__ = ["SparseConvTensor", "SparseModule", "SparseConvolution", "SparseConv2d", "SubMConv2d",
           "SparseInverseConv2d", "SparseConv3d", "SubMConv3d", "SparseInverseConv3d", "ToDense", "SparseSequential", "ops", "set_math_mode", "get_math_mode", "set_fused"]

MATH_MODES = ("bf16", "fp32", "bf16x3")
_math_mode = os.environ.get("WFSP_MATH", "bf16")
assert _math_mode in MATH_MODES


def set_math_mode(mode):
    """'bf16': tcgen05 tensor cores, bf16 operands / fp32 accumulate.  'fp32': exact fp32 CUDA-core path.
    'bf16x3': the tensor-core path at fp32-grade accuracy -- every operand split into hi + lo bf16 parts, three
    products per pair accumulated in fp32 (WFSP_MATH_BF16X3): the tight-tolerance mode without leaving tcgen05."""
    global _math_mode
    assert mode in MATH_MODES
    _math_mode = mode


def get_math_mode():
    return _math_mode


set_fused = fused.set_fused


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, grid=None, n_rows=None):
        """features [N, C]; indices int32 [N, 3] = (batch, x, y); spatial_shape e.g. [14, 11]
        (list / numpy array); batch_size int or 0-dim tensor (as the reference passes,
        src/models/SPConvNet.py:51,63).  The 3-d variant (net_type "3DConvolution", SPConvNet.py:42-49) has
        indices [N, 4] = (batch, x, y, t) and spatial_shape [14, 11, n_samples].

        n_rows (extension, graph path): int32 device scalar with the live row count; features /
        indices are then capacity-sized buffers, nothing downstream reads the count back to the host
        and the whole step can be captured in a CUDA graph."""
        self.features = features
        self.indices = indices
        if self.indices.dtype != torch.int32:
            self.indices = self.indices.int()
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict = {}
        self.grid = grid
        self.n_rows = n_rows

    @property
    def spatial_size(self):
        return int(np.prod(self.spatial_shape))

    def find_indice_pair(self, key):
        if key is None:
            return None
        if key in self.indice_dict:
            return self.indice_dict[key]
        return None

    def dense(self, channels_first=True):
        if len(self.spatial_shape) == 3:
            # [B, C, H, W, T] is [B, C, H, W*T] with the last two coordinates merged: same scatter kernel
            h, w, t = self.spatial_shape
            i = self.indices
            merged = torch.stack([i[:, 0], i[:, 1], i[:, 2] * t + i[:, 3]], dim=1).contiguous()
            out = Fsp.ToDenseFunction.apply(self.features, merged, self.batch_size, h, w * t, self.n_rows)
            out = out.view(out.shape[0], out.shape[1], h, w, t)
            return out if channels_first else out.permute(0, 2, 3, 4, 1).contiguous()
        h, w = self.spatial_shape
        out = Fsp.ToDenseFunction.apply(self.features, self.indices, self.batch_size, h, w, self.n_rows)
        if not channels_first:
            return out.permute(0, 2, 3, 1).contiguous()
        return out

    @property
    def sparity(self):
        return self.indices.shape[0] / max(1, self.spatial_size * self.batch_size)


class SparseModule(nn.Module):
    """Marker base class: modules that consume / produce a SparseConvTensor."""


class SparseConvolution(SparseModule):
    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, subm=False, output_padding=0, transposed=False, inverse=False, indice_key=None,
                 fused_bn=False, use_hash=False):
        super().__init__()
        assert groups == 1
        if ndim not in (2, 3):
            raise NotImplementedError("2-d (the 14x11 segment grid) and 3-d (14x11xsamples) sparse convolutions "
                                      "are implemented")
        if transposed:
            raise NotImplementedError("SparseConvTranspose is not used by the reference models")

        def tup(v):
            return [int(v)] * ndim if isinstance(v, (int, np.integer)) else [int(x) for x in v]

        self.ndim = ndim
        self.in_channels, self.out_channels = int(in_channels), int(out_channels)
        self.kernel_size, self.stride, self.padding = tup(kernel_size), tup(stride), tup(padding)
        self.dilation, self.output_padding = tup(dilation), tup(output_padding)
        self.conv1x1 = int(np.prod(self.kernel_size)) == 1
        self.transposed, self.inverse, self.groups, self.subm = transposed, inverse, groups, subm
        self.indice_key, self.fused_bn, self.use_hash = indice_key, fused_bn, use_hash
        self.math = None  # None -> follow the global math mode
        self.weight = nn.Parameter(torch.empty(*self.kernel_size, self.in_channels, self.out_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(self.out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / math.sqrt(fan_in)
            nn.init.uniform_(self.bias, -bound, bound)

    def extra_repr(self):
        return "{}, {}, kernel_size={}, stride={}, padding={}, dilation={}, subm={}, inverse={}, indice_key={}".format(
            self.in_channels, self.out_channels, self.kernel_size, self.stride, self.padding, self.dilation,
            self.subm, self.inverse, self.indice_key)

    def geometry(self, input, front_only=False):
        """Rulebook lookup / construction for this layer on `input` (upstream conv.py forward: reuse the
        entry cached under indice_key, else build and cache it -- also under key None).  Returns
        (rulebook or None for the 1x1 shortcut, output indices, output spatial shape, device row count
        of the output or None)."""
        if self.conv1x1:
            return None, input.indices, input.spatial_shape, input.n_rows
        datas = input.find_indice_pair(self.indice_key)
        if self.inverse:
            assert datas is not None and self.indice_key is not None
            rb = datas
            assert rb.kvol == int(np.prod(self.kernel_size)), \
                "inverse conv must have same kernel size as its couple conv"
            return rb, rb.indices, rb.spatial_shape, rb.n_in_dev
        if self.indice_key is not None and datas is not None and self._rulebook_matches(datas, input):
            rb = datas
        else:
            pad = [k // 2 for k in self.kernel_size] if self.subm else self.padding
            stride = [1] * self.ndim if self.subm else self.stride
            rb = ops.build_rulebook(input.indices, input.batch_size, input.spatial_shape, self.kernel_size, stride,
                                    pad, self.dilation, self.subm, n_rows=input.n_rows, front_only=front_only)
            input.indice_dict[self.indice_key] = rb
        out_spatial_shape = input.spatial_shape if self.subm else rb.out_spatial_shape
        return rb, rb.outids, out_spatial_shape, rb.n_out_dev

    def _rulebook_matches(self, rb, input):
        """A cached rulebook is reused only if it was built for THIS input: same index tensor, grid and kernel
        volume.  Upstream reuses whatever sits under the key; with a key shared across resolutions (submanifold ->
        strided -> submanifold with one key) that reads neighbour tables of the wrong active set and indexes past
        the feature buffer.  Here a mismatch rebuilds (and re-caches) instead."""
        idx = input.indices
        return (rb.kvol == int(np.prod(self.kernel_size)) and rb.indices.shape == idx.shape
                and rb.indices.data_ptr() == idx.data_ptr() and list(rb.spatial_shape) == list(input.spatial_shape))

    def forward(self, input):
        assert isinstance(input, SparseConvTensor)
        mode = self.math or _math_mode
        rb, outids, out_spatial_shape, out_rows = self.geometry(input)
        if rb is not None:
            rb.finish()
        out_features = Fsp.SparseConvFunction.apply(input.features, self.weight, self.bias, rb, self.inverse, mode,
                                                    input.n_rows if rb is None else None)
        out = SparseConvTensor(out_features, outids, out_spatial_shape, input.batch_size, n_rows=out_rows)
        out.indice_dict, out.grid = input.indice_dict, input.grid
        return out


class SparseConv2d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, use_hash=False):
        super().__init__(2, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         indice_key=indice_key, use_hash=use_hash)


class SubMConv2d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, use_hash=False):
        super().__init__(2, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, True,
                         indice_key=indice_key, use_hash=use_hash)


class SparseInverseConv2d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key, bias=True):
        super().__init__(2, in_channels, out_channels, kernel_size, bias=bias, inverse=True, indice_key=indice_key)


class SparseConv3d(SparseConvolution):
    """3-d counterpart (src/utils/ModelValidation.py:24-31 lists it; net_type "3DConvolution",
    src/models/SPConvNet.py:42-49).  Weight [kH, kW, kT, Cin, Cout]; same kernels, 3-d rulebook."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, use_hash=False):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         indice_key=indice_key, use_hash=use_hash)


class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, use_hash=False):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, True,
                         indice_key=indice_key, use_hash=use_hash)


class SparseInverseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key, bias=True):
        super().__init__(3, in_channels, out_channels, kernel_size, bias=bias, inverse=True, indice_key=indice_key)


def _not_implemented(name):
    class _Stub(SparseModule):
        def __init__(self, *args, **kwargs):
            raise NotImplementedError(
                "%s is listed by src/utils/ModelValidation.py:24-31 but constructed by no shipped config; "
                "the 2-d and 3-d regular / submanifold / inverse layers are implemented" % name)
    _Stub.__name__ = name
    return _Stub


SparseConv1d = _not_implemented("SparseConv1d")
SparseConv4d = _not_implemented("SparseConv4d")
SparseConvTranspose2d = _not_implemented("SparseConvTranspose2d")
SparseConvTranspose3d = _not_implemented("SparseConvTranspose3d")


class ToDense(SparseModule):
    """SparseConvTensor -> dense [B, C, H, W] (3-d: [B, C, H, W, T])."""

    def forward(self, x):
        return x.dense()


def is_spconv_module(module):
    return isinstance(module, SparseModule)


class SparseSequential(SparseModule):
    """nn.Sequential that hands the SparseConvTensor to sparse modules and `.features` to the
    rest (BatchNorm1d / ReLU / Dropout), as upstream spconv/modules.py does."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for key, module in args[0].items():
                self.add_module(key, module)
        else:
            for idx, module in enumerate(args):
                self.add_module(str(idx), module)
        for name, module in kwargs.items():
            if name in self._modules:
                raise ValueError("name exists.")
            self.add_module(name, module)
        self._sparity_dict = {}

    def __getitem__(self, idx):
        if not (-len(self) <= idx < len(self)):
            raise IndexError("index {} is out of range".format(idx))
        if idx < 0:
            idx += len(self)
        it = iter(self._modules.values())
        for _ in range(idx):
            next(it)
        return next(it)

    def __len__(self):
        return len(self._modules)

    @property
    def sparity_dict(self):
        return self._sparity_dict

    def add(self, module, name=None):
        if name is None:
            name = str(len(self._modules))
            if name in self._modules:
                raise KeyError("name exists")
        self.add_module(name, module)

    def forward(self, input):
        mods = list(self._modules.items())
        if (fused.is_enabled() and isinstance(input, SparseConvTensor) and input.features is not None
                and input.features.is_cuda and input.indices.shape[0] > 0
                and all((m.math or _math_mode) == "bf16" for _, m in mods if isinstance(m, SparseConvolution))):
            # whole stack as one autograd node with bf16-resident operands (fused.py); None = not covered
            plan = fused.compile_stack([m for _, m in mods], SparseConvolution, ToDense)
            if plan is not None:
                plan.prepared = self.__dict__.pop("_wfsp_prepared", None)
                for k, module in mods:
                    if is_spconv_module(module):
                        self._sparity_dict[k] = input.sparity
                return fused.run(plan, input)
        # per-layer path from here on; work a caller started early on side streams for the fused path is joined first
        stale = self.__dict__.pop("_wfsp_prepared", None)
        if stale is not None:
            torch.cuda.current_stream().wait_event(stale["done"])
        if isinstance(input, SparseConvTensor) and getattr(input.features, "_wfsp_ready", None) is not None:
            torch.cuda.current_stream().wait_event(input.features._wfsp_ready)
        first = next((m for _, m in mods if isinstance(m, SparseConvolution)), None)
        if (first is not None and isinstance(input, SparseConvTensor) and input.features is not None
                and fused.is_operand_format(input.features, first.in_channels)
                and input.features.shape[1] != first.in_channels):
            # per-layer path fed with the padded bf16 operand format: back to [N, C] fp32
            input.features = input.features[:, :first.in_channels].float()
        i = 0
        while i < len(mods):
            k, module = mods[i]
            i += 1
            if is_spconv_module(module):
                assert isinstance(input, SparseConvTensor)
                self._sparity_dict[k] = input.sparity
                input = module(input)
            elif isinstance(input, SparseConvTensor):
                if input.n_rows is not None:
                    # graph path: row count lives on the device, so BatchNorm1d statistics must be taken
                    # over the live rows by our own kernel (fused with a directly following ReLU);
                    # element-wise modules may run over the whole capacity-sized buffer
                    if isinstance(module, nn.BatchNorm1d):
                        own = getattr(module, "fused_relu", False)  # sparseconvnet.BatchNormReLU facade
                        fuse = (not own) and i < len(mods) and isinstance(mods[i][1], nn.ReLU)
                        input.features = Fsp.batch_norm_relu(input.features, input.n_rows, module, fuse or own)
                        i += 1 if fuse else 0
                    elif isinstance(module, (nn.ReLU, nn.Dropout, nn.Identity, nn.LeakyReLU, nn.Sigmoid, nn.Tanh)):
                        input.features = module(input.features)
                    else:
                        raise NotImplementedError("graph path: %s between sparse layers" % type(module).__name__)
                elif input.indices.shape[0] != 0:
                    input.features = module(input.features)
            else:
                input = module(input)
        return input
