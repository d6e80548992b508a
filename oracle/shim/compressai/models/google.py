"""compressai.models.google restated: FactorizedPrior / ScaleHyperprior, the models behind
COMPRESSAI_DICT['bmshj2018_factorized'|'bmshj2018_hyperprior'] (sc2bench/models/registry.py:12-14,73)."""
import torch
from torch import nn

from ..entropy_models import EntropyBottleneck, GaussianConditional
from ..layers import GDN
from .base import CompressionModel, get_scale_table  # noqa: F401
from .utils import conv, deconv


class FactorizedPrior(CompressionModel):
    def __init__(self, N, M, **kwargs):
        super().__init__(**kwargs)
        self.entropy_bottleneck = EntropyBottleneck(M)
        self.g_a = nn.Sequential(conv(3, N), GDN(N), conv(N, N), GDN(N), conv(N, N), GDN(N), conv(N, M))
        self.g_s = nn.Sequential(deconv(M, N), GDN(N, inverse=True), deconv(N, N), GDN(N, inverse=True),
                                 deconv(N, N), GDN(N, inverse=True), deconv(N, 3))
        self.N, self.M = N, M

    def forward(self, x):
        y = self.g_a(x)
        y_hat, y_likelihoods = self.entropy_bottleneck(y)
        return {'x_hat': self.g_s(y_hat), 'likelihoods': {'y': y_likelihoods}}

    def compress(self, x):
        y = self.g_a(x)
        return {'strings': [self.entropy_bottleneck.compress(y)], 'shape': y.size()[-2:]}

    def decompress(self, strings, shape):
        assert isinstance(strings, list) and len(strings) == 1
        y_hat = self.entropy_bottleneck.decompress(strings[0], shape)
        return {'x_hat': self.g_s(y_hat).clamp_(0, 1)}


class ScaleHyperprior(CompressionModel):
    def __init__(self, N, M, **kwargs):
        super().__init__(**kwargs)
        self.entropy_bottleneck = EntropyBottleneck(N)
        self.g_a = nn.Sequential(conv(3, N), GDN(N), conv(N, N), GDN(N), conv(N, N), GDN(N), conv(N, M))
        self.g_s = nn.Sequential(deconv(M, N), GDN(N, inverse=True), deconv(N, N), GDN(N, inverse=True),
                                 deconv(N, N), GDN(N, inverse=True), deconv(N, 3))
        self.h_a = nn.Sequential(conv(M, N, stride=1, kernel_size=3), nn.ReLU(inplace=True), conv(N, N),
                                 nn.ReLU(inplace=True), conv(N, N))
        self.h_s = nn.Sequential(deconv(N, N), nn.ReLU(inplace=True), deconv(N, N), nn.ReLU(inplace=True),
                                 conv(N, M, stride=1, kernel_size=3), nn.ReLU(inplace=True))
        self.gaussian_conditional = GaussianConditional(None)
        self.N, self.M = int(N), int(M)

    def forward(self, x):
        y = self.g_a(x)
        z = self.h_a(torch.abs(y))
        z_hat, z_likelihoods = self.entropy_bottleneck(z)
        scales_hat = self.h_s(z_hat)
        y_hat, y_likelihoods = self.gaussian_conditional(y, scales_hat)
        return {'x_hat': self.g_s(y_hat), 'likelihoods': {'y': y_likelihoods, 'z': z_likelihoods}}

    def compress(self, x):
        y = self.g_a(x)
        z = self.h_a(torch.abs(y))
        z_strings = self.entropy_bottleneck.compress(z)
        z_hat = self.entropy_bottleneck.decompress(z_strings, z.size()[-2:])
        scales_hat = self.h_s(z_hat)
        indexes = self.gaussian_conditional.build_indexes(scales_hat)
        y_strings = self.gaussian_conditional.compress(y, indexes)
        return {'strings': [y_strings, z_strings], 'shape': z.size()[-2:]}

    def decompress(self, strings, shape):
        assert isinstance(strings, list) and len(strings) == 2
        z_hat = self.entropy_bottleneck.decompress(strings[1], shape)
        scales_hat = self.h_s(z_hat)
        indexes = self.gaussian_conditional.build_indexes(scales_hat)
        y_hat = self.gaussian_conditional.decompress(strings[0], indexes, z_hat.dtype)
        return {'x_hat': self.g_s(y_hat).clamp_(0, 1)}
