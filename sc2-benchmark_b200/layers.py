"""GDN / GDN1 with CompressAI's parameters and state-dict keys, evaluated by one fused kernel at inference.

Drop-in for `compressai.layers.GDN1` (sc2bench/models/layer.py:3, instantiated at :478,481,488,491) and
`compressai.layers.GDN` (inside the bmshj2018 zoo models, sc2bench/models/registry.py:12-14).
Keys: `beta`, `gamma`, `beta_reparam.{pedestal,lower_bound.bound}`, `gamma_reparam.{...}`.
"""
import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .entropy_models import LowerBound


class NonNegativeParametrizer(nn.Module):
    """x -> max(x, sqrt(minimum + pedestal))^2 - pedestal  (keeps beta / gamma non-negative)."""

    def __init__(self, minimum=0, reparam_offset=2 ** -18):
        super().__init__()
        self.minimum = float(minimum)
        self.reparam_offset = float(reparam_offset)
        pedestal = self.reparam_offset ** 2
        self.register_buffer('pedestal', torch.Tensor([pedestal]))
        self.lower_bound = LowerBound((self.minimum + pedestal) ** 0.5)

    def init(self, x):
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

    def forward(self, x):
        return self.lower_bound(x) ** 2 - self.pedestal


class GDN(nn.Module):
    """y_i = x_i * (beta_i + sum_j gamma_ij x_j^2)^(-1/2)   (inverse: ^(+1/2))."""
    _kind = 1

    def __init__(self, in_channels, inverse=False, beta_min=1e-6, gamma_init=0.1):
        super().__init__()
        self.inverse = bool(inverse)
        self.beta_reparam = NonNegativeParametrizer(minimum=float(beta_min))
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))
        self.gamma_reparam = NonNegativeParametrizer()
        self.gamma = nn.Parameter(self.gamma_reparam.init(float(gamma_init) * torch.eye(in_channels)))

    def effective_params(self):
        """(gamma [C, C], beta [C]) after reparametrisation."""
        return self.gamma_reparam(self.gamma), self.beta_reparam(self.beta)

    def _norm_input(self, x):
        return x * x

    def _apply_norm(self, x, norm):
        return x * (torch.sqrt(norm) if self.inverse else torch.rsqrt(norm))

    def forward(self, x):
        gamma, beta = self.effective_params()
        if not (torch.is_grad_enabled() and (x.requires_grad or self.gamma.requires_grad)):
            return ops.gdn(x, gamma, beta, kind=self._kind, inverse=self.inverse)  # inference: CUDA only, raises on CPU
        # differentiable branch (training, off the hot path): plain torch ops
        C = x.size(1)
        norm = F.conv2d(self._norm_input(x), gamma.reshape(C, C, 1, 1), beta)
        return self._apply_norm(x, norm)


class GDN1(GDN):
    """y_i = x_i / (beta_i + sum_j gamma_ij |x_j|)   (inverse: multiply)."""
    _kind = 0

    def _norm_input(self, x):
        return torch.abs(x)

    def _apply_norm(self, x, norm):
        return x * norm if self.inverse else x * (1.0 / norm)
