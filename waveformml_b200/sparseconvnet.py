"""SparseConvNet-signature facade over the same sm_100a kernels.

The reference can build its 2-d classifier with facebookresearch/SparseConvNet instead of spconv
(src/models/SCNet.py:62-77, config/examples/OPs3ns_SCNet.json:22-66, src/utils/ModelValidation.py:16-18):

    scn.InputLayer(2, spatial_size, mode=0)          coords arrive batch-LAST: (x, y, batch) int64
                                                     (dimension 3, SCNet.py:53-55: (x, y, t, batch), size [14, 11, ns])
    scn.Convolution(dim, nIn, nOut, filter_size, filter_stride, bias)
    scn.SubmanifoldConvolution(dim, nIn, nOut, filter_size, bias)
    scn.BatchNormReLU(nPlanes) / scn.BatchNormalization(nPlanes) / scn.ReLU()
    scn.SparseToDense(dim, nPlanes)                  -> [B, nPlanes, H', W']
    scn.Sequential(*modules)

SparseConvNet itself is an absent third-party dependency (unpinned in the reference, README.md:44-46), so
-- like spconv -- parity is against its published semantics: a strided Convolution has no padding and an
output site is active iff its receptive field holds an active input (= spconv.SparseConv2d with padding 0);
weights are [filter_volume, nIn, nOut] with the filter offsets in row-major order; init N(0, 2/(nIn*volume)).
The active set and dense outputs are what the tests pin (dense convolution identities); SparseConvNet's
internal row order is not reproduced (only SparseToDense / per-site features are observable in the reference).
"""
import math

import torch
from torch import nn

from . import spconv

__all__ = ["InputLayer", "Convolution", "SubmanifoldConvolution", "BatchNormReLU", "BatchNormalization", "ReLU",
           "SparseToDense", "Sequential", "OutputLayer"]


def _tup(v, dim):
    return [int(v)] * dim if isinstance(v, int) else [int(x) for x in v]


class InputLayer(nn.Module):
    """[coords (x, y, batch) LongTensor [N, 3], features [N, C]] -> sparse tensor.  mode 0: coordinates are
    unique (what the reference passes); other modes (duplicate handling) are not implemented."""

    def __init__(self, dimension, spatial_size, mode=3):
        super().__init__()
        if dimension not in (2, 3):
            raise NotImplementedError("the 2-d 14x11 segment grid and the 3-d 14x11xsamples grid are implemented")
        if mode != 0:
            raise NotImplementedError("InputLayer mode %d (duplicate coordinates) is not used by the reference" % mode)
        self.dimension, self.mode = dimension, mode
        self.spatial_size = [int(s) for s in (spatial_size.tolist() if torch.is_tensor(spatial_size) else spatial_size)]

    def forward(self, x):
        coords, feats = x[0], x[1]
        batch_size = x[2] if len(x) > 2 else int(coords[:, -1].max()) + 1
        d = self.dimension  # batch-last (x, y[, t], batch) -> batch-first, as SCNet.forward / SPConvNet.forward permute
        idx = coords[:, [d] + list(range(d))].to(torch.int32).contiguous()
        return spconv.SparseConvTensor(feats, idx, self.spatial_size, batch_size)


class _SCNConv(spconv.SparseConvolution):
    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, subm, groups=1):
        if dimension not in (2, 3):
            raise NotImplementedError("2-d and 3-d sparse convolutions are implemented (SCNet.py:51-55)")
        assert groups == 1
        fs = _tup(filter_size, dimension)
        st = _tup(filter_stride, dimension)
        super().__init__(dimension, nIn, nOut, fs, st, 0, 1, 1, bias, subm=subm, indice_key=None)
        self.dimension, self.nIn, self.nOut = dimension, nIn, nOut
        self.filter_size, self.filter_stride = fs, st
        self.filter_volume = int(math.prod(fs))
        # SparseConvNet parameter layout and init: [filter_volume, nIn, nOut], N(0, sqrt(2 / (nIn * volume)))
        std = math.sqrt(2.0 / nIn / self.filter_volume)
        self.weight = nn.Parameter(torch.empty(self.filter_volume, nIn, nOut).normal_(0, std))
        if bias:
            self.bias = nn.Parameter(torch.zeros(nOut))

    def reset_parameters(self):  # parameters are (re)created in __init__ with SparseConvNet's init
        pass


class Convolution(_SCNConv):
    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        super().__init__(dimension, nIn, nOut, filter_size, filter_stride, bias, False, groups)


class SubmanifoldConvolution(_SCNConv):
    def __init__(self, dimension, nIn, nOut, filter_size, bias, groups=1):
        super().__init__(dimension, nIn, nOut, filter_size, 1, bias, True, groups)
        # same-size submanifold layers AT ONE RESOLUTION share a rulebook: the key is completed with the input's
        # spatial shape at run time (geometry), so the layers before and after a strided Convolution never collide
        self._key_base = "scn_subm" + "x".join(str(k) for k in self.filter_size)
        self.indice_key = self._key_base

    def geometry(self, input, front_only=False):
        self.indice_key = self._key_base + "@" + "x".join(str(int(s)) for s in input.spatial_shape)
        return super().geometry(input, front_only=front_only)


class BatchNormalization(nn.BatchNorm1d):
    """SparseConvNet's momentum is the weight of the OLD running value (default 0.9); torch's is the new one's."""

    fused_relu = False

    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, affine=True, leakiness=1):
        super().__init__(nPlanes, eps=eps, momentum=1.0 - momentum, affine=affine)
        if leakiness not in (0, 1):
            raise NotImplementedError("leaky BatchNorm activations are not used by the reference")
        self.fused_relu = leakiness == 0

    def forward(self, x):
        y = super().forward(x)
        return torch.relu(y) if self.fused_relu else y


class BatchNormReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9):
        super().__init__(nPlanes, eps, momentum, True, leakiness=0)


class ReLU(nn.ReLU):
    pass


class SparseToDense(spconv.ToDense):
    def __init__(self, dimension, nPlanes):
        super().__init__()
        self.dimension, self.nPlanes = dimension, nPlanes


class OutputLayer(spconv.SparseModule):
    """sparse tensor -> per-site feature rows (in the input's row order for submanifold stacks)."""

    def __init__(self, dimension):
        super().__init__()
        self.dimension = dimension

    def forward(self, x):
        return x.features


class Sequential(spconv.SparseSequential):
    pass
