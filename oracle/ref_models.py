"""oracle/ref_models.py -- CPU restatement of the reference's FPBasedResNetBottleneck encode()/decode()
(sc2bench/models/layer.py:464-521) on top of oracle/shim/compressai, for places where /root/reference
does not exist (the GPU box): bench.py's cpu_baseline / --impl reference legs, __graft_entry__.smoke(),
GPU parity tests at sizes other than the golden fixtures.

TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (see oracle/README.md).  tests/test_oracle.py checks, in the
build container, that this restatement equals the reference's real class output for output.
"""
import os
import sys

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shim')


def _import_shim():
    """Imports oracle/shim/compressai under its own name unless a real compressai is already loaded."""
    import importlib.util
    mod = sys.modules.get('compressai')
    if mod is not None:
        return mod
    init = os.path.join(_SHIM, 'compressai', '__init__.py')
    spec = importlib.util.spec_from_file_location('compressai', init, submodule_search_locations=[os.path.dirname(init)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules['compressai'] = mod
    spec.loader.exec_module(mod)
    return mod


def build_fp_bottleneck(num_input_channels=3, num_bottleneck_channels=24, num_target_channels=256):
    _import_shim()
    import torch
    from torch import nn
    from compressai.layers import GDN1
    from compressai.models import CompressionModel

    class OracleFPBottleneck(CompressionModel):
        def __init__(self):
            super().__init__(entropy_bottleneck_channels=num_bottleneck_channels)
            b, t = num_bottleneck_channels, num_target_channels
            self.updated = False
            self.encoder = nn.Sequential(
                nn.Conv2d(num_input_channels, b * 4, kernel_size=5, stride=2, padding=2, bias=False), GDN1(b * 4),
                nn.Conv2d(b * 4, b * 2, kernel_size=5, stride=2, padding=2, bias=False), GDN1(b * 2),
                nn.Conv2d(b * 2, b, kernel_size=2, stride=1, padding=0, bias=False))
            self.decoder = nn.Sequential(
                nn.Conv2d(b, t * 2, kernel_size=2, stride=1, padding=1, bias=False), GDN1(t * 2, inverse=True),
                nn.Conv2d(t * 2, t, kernel_size=2, stride=1, padding=0, bias=False), GDN1(t, inverse=True),
                nn.Conv2d(t, t, kernel_size=2, stride=1, padding=1, bias=False))

        def update(self, force=False):
            self.updated = True
            return super().update(force=force)

        def encode(self, x):
            latent = self.encoder(x)
            return {'strings': [self.entropy_bottleneck.compress(latent)], 'shape': latent.size()[-2:]}

        def decode(self, strings, shape):
            return self.decoder(self.entropy_bottleneck.decompress(strings[0], shape))

        @torch.no_grad()
        def symbols(self, x):
            latent = self.encoder(x)
            med = self.entropy_bottleneck._get_medians().detach().reshape(1, -1, 1, 1)
            return latent, torch.round(latent - med).int()

    return OracleFPBottleneck()
