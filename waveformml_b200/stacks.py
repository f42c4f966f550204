"""The concrete layer stacks the reference's planners construct, built on
the drop-in spconv layers.  tests/test_reference_planners.py checks them layer by layer against
what /root/reference's own planners build (golden fixture tests/golden/planner_stacks.json).

State-dict key names follow the reference modules (`sparseModel.N.weight`, `linear.N.weight`,
`model.network.N.weight`) so reference checkpoints load.
"""
import math

import torch
from torch import nn

from . import spconv

SPATIAL = [14, 11]


def _wrap(x, spatial_size):
    """x = [indices (batch, x, y) int32, features] (+ optional batch size, which skips the readback the
    reference does with `x[0][-1, -1] + 1` (SPConvNet.py:63); + optional int32 device scalar with the
    live row count = the graph path, see spconv.SparseConvTensor)."""
    indices, feats = x[0], x[1]
    batch_size = x[2] if len(x) > 2 else int(indices[-1, 0]) + 1
    n_rows = x[3] if len(x) > 3 else None
    return spconv.SparseConvTensor(feats, indices, spatial_size, batch_size, n_rows=n_rows)


def _conv_out(size, k, s, p, d):
    return [(i + 2 * p - d * (k - 1) - 1) // s + 1 for i in size]


def gep_channel_ladder(nin=300, nout=20, n=3, pointwise_factor=0.1735):
    """SparseConv2DBlock._version0 channel arithmetic (src/models/SPConvBlocks.py:460-470)."""
    frames = [nin, nin - int(math.floor((nin - nout) * pointwise_factor))]
    diff = float(nin - nout) / n
    for _ in range(n - 1):
        val = int(math.floor(frames[-1] - diff))
        frames.append(val if val > nout else nout)
    return frames


class PSDClassifier(nn.Module):
    """config/examples/GEP.json -> SPConvNet + SparseConv2DBlock._version0 (SPConvBlocks.py:450-516):
    SparseConv2d(300,252,1) BN ReLU . SparseConv2d(252,158,3) BN ReLU . SparseConv2d(158,64,3) BN ReLU .
    ToDense -> [B,64,10,7] -> view [B,4480] -> Linear(4480,116) . Linear(116,3)   (all convs bias=False)."""

    def __init__(self, n_samples=150, n_classes=3, channels=None, n_lin=2):
        super().__init__()
        nin = 2 * n_samples
        ch = channels or gep_channel_ladder(nin)
        assert ch[0] == nin
        layers, size = [], list(SPATIAL)
        for i in range(len(ch) - 1):
            k = 1 if i == 0 else 3
            layers += [spconv.SparseConv2d(ch[i], ch[i + 1], k, 1, 0, 1, 1, False), nn.BatchNorm1d(ch[i + 1]), nn.ReLU()]
            size = _conv_out(size, k, 1, 0, 1)
        layers.append(spconv.ToDense())
        self.sparseModel = spconv.SparseSequential(*layers)
        self.out_size = size + [ch[-1]]
        self.n_linear = size[0] * size[1] * ch[-1]
        # LinearBlock (src/models/ConvBlocks.py:82-102): geometric ladder, no activation in between
        factor = pow(float(n_classes) / self.n_linear, 1.0 / n_lin)
        lin = [nn.Linear(int(round(self.n_linear * pow(factor, i))), int(round(self.n_linear * pow(factor, i + 1))))
               for i in range(n_lin)]
        self.linear = nn.Sequential(*lin)
        self.spatial_size = SPATIAL

    def forward(self, x):
        d = self.sparseModel(_wrap(x, self.spatial_size))
        return self.linear(d.view(-1, self.n_linear))

    def forward_loss(self, x, target, criterion):
        """criterion(self(x), target) -- LitPSD.training_step (src/engineering/LitPSD.py:94-104) -- with the dense
        head and the loss fused into three launches when the head / loss / batch are the supported case."""
        from . import head
        d = self.sparseModel(_wrap(x, self.spatial_size)).view(-1, self.n_linear)
        if head.supported(self.linear, d, criterion):
            return head.head_cross_entropy(self.linear, d, target)
        return criterion(self.linear(d), target)


class ZRegressor(nn.Module):
    """config/examples/SingleEndedZCNN.json -> SingleEndedZConv + SparseConv2DForZ (SPConvBlocks.py:261-313):
    SparseConv2d(300,150,3,1,1) BN ReLU . SparseConv2d(150,1,1,1,0) ReLU . ToDense -> [B,1,14,11]."""

    def __init__(self, n_samples=150):
        super().__init__()
        nin = 2 * n_samples
        mid = nin - int(round(float(nin) / 2))
        net = spconv.SparseSequential(
            spconv.SparseConv2d(nin, mid, 3, 1, 1), nn.BatchNorm1d(mid), nn.ReLU(),
            spconv.SparseConv2d(mid, 1, 1, 1, 0), nn.ReLU(), spconv.ToDense())
        self.model = nn.Module()
        self.model.network = net
        self.spatial_size = SPATIAL

    def forward(self, x):
        return self.model.network(_wrap(x, self.spatial_size))


class EZSubM(nn.Module):
    """SparseConv2DForEZ(nin, out_planes, kernel_size=5, n_conv=2, n_point=3, conv_position=2, version=2)
    (SPConvBlocks.py:9-258, Appendix B.4): SubM k1 . k5 ('subm5') . k5 ('subm5', rulebook reused) . k1 . k1 . ToDense."""

    def __init__(self, nin=300, out_planes=2, kernel_size=5, pointwise_factor=0.8):
        super().__init__()
        n_layers = 5
        inc = int(round(int(round(nin * pointwise_factor - out_planes)) / float(n_layers - 1)))
        chans, out = [nin], nin
        for i in range(n_layers):
            if i == n_layers - 1:
                out = out_planes
            else:
                out -= inc
                if i == 0:
                    out = int(round(pointwise_factor * nin))
            chans.append(max(out, 1))
        ks = [1, kernel_size, kernel_size, 1, 1]
        layers = []
        for i in range(n_layers):
            k = ks[i]
            key = "subm0" if k < 4 else "subm{}".format(k)
            layers.append(spconv.SubMConv2d(chans[i], chans[i + 1], k, 1, (k - 1) // 2, indice_key=key))
            if i != n_layers - 1:
                layers.append(nn.BatchNorm1d(chans[i + 1]))
            layers.append(nn.ReLU())
        layers.append(spconv.ToDense())
        self.network = spconv.SparseSequential(*layers)
        self.spatial_size = SPATIAL

    def forward(self, x):
        return self.network(_wrap(x, self.spatial_size))


class IoniPreserve(nn.Module):
    """config/examples/IoniClassifierCNN.json -> SPConvPreserveNet + SparseConv2DPreserve._version0
    (SPConvBlocks.py:756-822, Appendix B.3): 6 x [SparseConv2d(c_i,c_{i+1},k,1,p, indice_key='ind_i') .
    SparseInverseConv2d(c_{i+1},c_{i+1},k,'ind_i') . BN . ReLU]; returns per-hit features."""

    CHANNELS = [130, 138, 146, 154, 104, 54, 5]
    KP = [(3, 1), (3, 1), (2, 0), (2, 0), (2, 0), (2, 0)]

    def __init__(self):
        super().__init__()
        layers = []
        for i, (k, p) in enumerate(self.KP):
            ci, co = self.CHANNELS[i], self.CHANNELS[i + 1]
            layers += [spconv.SparseConv2d(ci, co, k, 1, p, 1, 1, False, indice_key="ind_{}".format(i)),
                       spconv.SparseInverseConv2d(co, co, k, "ind_{}".format(i), bias=False),
                       nn.BatchNorm1d(co), nn.ReLU()]
        self.model = nn.Module()
        self.model.func = spconv.SparseSequential(*layers)
        self.spatial_size = SPATIAL

    def forward(self, x):
        return self.model.func(_wrap(x, self.spatial_size)).features


def describe(module):
    """Flat, comparable description of a layer stack (used by the planner parity test)."""
    out = []
    for m in module.modules():
        if isinstance(m, spconv.SparseConvolution):
            out.append({"type": "SubMConv2d" if m.subm else ("SparseInverseConv2d" if m.inverse else "SparseConv2d"),
                        "cin": m.in_channels, "cout": m.out_channels, "k": m.kernel_size, "s": m.stride,
                        "p": m.padding, "d": m.dilation, "bias": m.bias is not None, "key": m.indice_key})
        elif isinstance(m, spconv.ToDense):
            out.append({"type": "ToDense"})
        elif isinstance(m, nn.BatchNorm1d):
            out.append({"type": "BatchNorm1d", "c": m.num_features})
        elif isinstance(m, nn.ReLU):
            out.append({"type": "ReLU"})
        elif isinstance(m, nn.Dropout):
            out.append({"type": "Dropout", "p": m.p})
        elif isinstance(m, nn.Linear):
            out.append({"type": "Linear", "cin": m.in_features, "cout": m.out_features})
    return out
