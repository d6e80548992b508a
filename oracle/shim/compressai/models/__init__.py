from .base import CompressionModel, get_scale_table  # noqa: F401
from .google import FactorizedPrior, ScaleHyperprior  # noqa: F401
