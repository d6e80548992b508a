"""compressai.zoo.image restated for the two zoo entries on the path (no pretrained download: no network)."""
from ..models import FactorizedPrior, ScaleHyperprior

_CFGS = {1: (128, 192), 2: (128, 192), 3: (128, 192), 4: (128, 192), 5: (128, 192), 6: (192, 320), 7: (192, 320), 8: (192, 320)}
model_architectures = {'bmshj2018-factorized': FactorizedPrior, 'bmshj2018-hyperprior': ScaleHyperprior}


def _build(arch, quality, metric, pretrained, **kwargs):
    if metric not in ('mse', 'ms-ssim'):
        raise ValueError(f'Invalid metric "{metric}"')
    if quality < 1 or quality > 8:
        raise ValueError(f'Invalid quality "{quality}", should be between (1, 8)')
    if pretrained:
        raise RuntimeError('pretrained weights need network access; the oracle restatement has none')
    return model_architectures[arch](*_CFGS[quality], **kwargs)


def bmshj2018_factorized(quality, metric='mse', pretrained=False, progress=True, **kwargs):
    return _build('bmshj2018-factorized', quality, metric, pretrained, **kwargs)


def bmshj2018_hyperprior(quality, metric='mse', pretrained=False, progress=True, **kwargs):
    return _build('bmshj2018-hyperprior', quality, metric, pretrained, **kwargs)
