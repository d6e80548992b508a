"""compressai.models.CompressionModel restated; base class of every sc2bench bottleneck
(sc2bench/models/layer.py:346,401) and the isinstance target at sc2bench/models/backbone.py:154,276."""
import math
import warnings

import torch
from torch import nn

from ..entropy_models import EntropyBottleneck, GaussianConditional
from .utils import update_registered_buffers

SCALES_MIN, SCALES_MAX, SCALES_LEVELS = 0.11, 256, 64


def get_scale_table(min=SCALES_MIN, max=SCALES_MAX, levels=SCALES_LEVELS):
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


class CompressionModel(nn.Module):
    def __init__(self, entropy_bottleneck_channels=None, init_weights=None):
        super().__init__()
        if entropy_bottleneck_channels is not None:
            warnings.warn('The entropy_bottleneck_channels parameter is deprecated.', DeprecationWarning, stacklevel=2)
            self.entropy_bottleneck = EntropyBottleneck(entropy_bottleneck_channels)

    def load_state_dict(self, state_dict, strict=True):
        for name, module in self.named_modules():
            if not any(x.startswith(name) for x in state_dict.keys()):
                continue
            if isinstance(module, EntropyBottleneck):
                update_registered_buffers(module, name, ['_quantized_cdf', '_offset', '_cdf_length'], state_dict)
            if isinstance(module, GaussianConditional):
                update_registered_buffers(module, name, ['_quantized_cdf', '_offset', '_cdf_length', 'scale_table'], state_dict)
        return nn.Module.load_state_dict(self, state_dict, strict=strict)

    def update(self, scale_table=None, force=False):
        if scale_table is None:
            scale_table = get_scale_table()
        updated = False
        for _, module in self.named_modules():
            if isinstance(module, EntropyBottleneck):
                updated |= module.update(force=force)
            if isinstance(module, GaussianConditional):
                updated |= module.update_scale_table(scale_table, force=force)
        return updated

    def aux_loss(self):
        return sum(m.loss() for m in self.modules() if isinstance(m, EntropyBottleneck))
