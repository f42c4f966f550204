"""CPU ORACLE helper -- TEST INFRASTRUCTURE ONLY.  Builds the oracle (oracle/spconv_cpu.py) twin of a
layer stack made of waveformml_b200.spconv modules, sharing the weights, so a test can run the
same model through both and compare."""
import copy

import torch
from torch import nn

from oracle import spconv_cpu as osp
from waveformml_b200 import spconv as gsp


def to_oracle(module):
    if isinstance(module, gsp.SparseSequential):
        return osp.SparseSequential(*[to_oracle(m) for m in module._modules.values()])
    if isinstance(module, gsp.SparseConvolution):
        bias = module.bias is not None
        if module.inverse:
            o = osp.SparseInverseConv2d(module.in_channels, module.out_channels, module.kernel_size,
                                        module.indice_key, bias=bias)
        elif module.subm:
            o = osp.SubMConv2d(module.in_channels, module.out_channels, module.kernel_size, module.stride,
                               module.padding, module.dilation, 1, bias, indice_key=module.indice_key)
        else:
            o = osp.SparseConv2d(module.in_channels, module.out_channels, module.kernel_size, module.stride,
                                 module.padding, module.dilation, 1, bias, indice_key=module.indice_key)
        with torch.no_grad():
            o.weight.copy_(module.weight.detach().cpu())
            if bias:
                o.bias.copy_(module.bias.detach().cpu())
        return o
    if isinstance(module, gsp.ToDense):
        return osp.ToDense()
    if isinstance(module, nn.Sequential):
        return nn.Sequential(*[to_oracle(m) for m in module])
    return copy.deepcopy(module).cpu()


def run_stack(stack, indices, feats, spatial_shape, batch_size):
    return stack(osp.SparseConvTensor(feats, indices, spatial_shape, batch_size))
