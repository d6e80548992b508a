"""GDN / GDN1 restated; instantiated at sc2bench/models/layer.py:478,481,488,491."""
import torch
import torch.nn.functional as F
from torch import nn

from ..ops import NonNegativeParametrizer


class GDN(nn.Module):
    def __init__(self, in_channels, inverse=False, beta_min=1e-6, gamma_init=0.1):
        super().__init__()
        self.inverse = bool(inverse)
        self.beta_reparam = NonNegativeParametrizer(minimum=float(beta_min))
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))
        self.gamma_reparam = NonNegativeParametrizer()
        self.gamma = nn.Parameter(self.gamma_reparam.init(float(gamma_init) * torch.eye(in_channels)))

    def _params(self, C):
        return self.gamma_reparam(self.gamma).reshape(C, C, 1, 1), self.beta_reparam(self.beta)

    def forward(self, x):
        C = x.size(1)
        gamma, beta = self._params(C)
        norm = F.conv2d(x ** 2, gamma, beta)
        norm = torch.sqrt(norm) if self.inverse else torch.rsqrt(norm)
        return x * norm


class GDN1(GDN):
    def forward(self, x):
        C = x.size(1)
        gamma, beta = self._params(C)
        norm = F.conv2d(torch.abs(x), gamma, beta)
        if not self.inverse:
            norm = 1.0 / norm
        return x * norm
