"""Installs this package under the CompressAI module names the reference imports, so that the UNMODIFIED
`sc2bench` code (sc2bench/models/layer.py:2-6, backbone.py:4, registry.py:2) builds its models out of the CUDA-backed
classes.  Only the names on the bottleneck path are provided; anything else raises AttributeError.

    import sc2bench_b200.compat as compat
    compat.install_as_compressai()        # before `import sc2bench`
"""
import sys
import types

from . import entropy_models, layers, models


def _module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__dict__['__sc2bench_b200__'] = True
    sys.modules[name] = mod
    return mod


def _unavailable(what):
    def fn(*args, **kwargs):
        raise NotImplementedError('%s is off the bottleneck path and is not provided by sc2bench_b200' % what)
    return fn


def install_as_compressai(force=False):
    """Registers `compressai` and the sub-modules sc2bench touches in sys.modules.  Refuses to shadow a real
    CompressAI install unless force=True."""
    existing = sys.modules.get('compressai')
    if existing is not None and not getattr(existing, '__sc2bench_b200__', False) and not force:
        raise RuntimeError('a different `compressai` is already imported; pass force=True to shadow it')
    root = _module('compressai', __version__='1.2.x-compatible (sc2bench_b200)', __path__=[])
    root.entropy_models = _module('compressai.entropy_models', EntropyModel=entropy_models.EntropyModel,
                                  EntropyBottleneck=entropy_models.EntropyBottleneck,
                                  GaussianConditional=entropy_models.GaussianConditional)
    root.layers = _module('compressai.layers', GDN=layers.GDN, GDN1=layers.GDN1)
    root.ops = _module('compressai.ops', LowerBound=entropy_models.LowerBound,
                       NonNegativeParametrizer=layers.NonNegativeParametrizer)
    root.models = _module('compressai.models', __path__=[], CompressionModel=models.CompressionModel,
                          FactorizedPrior=models.FactorizedPrior, ScaleHyperprior=models.ScaleHyperprior,
                          get_scale_table=models.get_scale_table)
    root.models.google = _module('compressai.models.google', get_scale_table=models.get_scale_table,
                                 FactorizedPrior=models.FactorizedPrior, ScaleHyperprior=models.ScaleHyperprior,
                                 CompressionModel=models.CompressionModel)
    root.models.utils = _module('compressai.models.utils', update_registered_buffers=models.update_registered_buffers,
                                conv=models.conv, deconv=models.deconv)
    root.zoo = _module('compressai.zoo', __path__=[], bmshj2018_factorized=models.bmshj2018_factorized,
                       bmshj2018_hyperprior=models.bmshj2018_hyperprior)
    root.zoo.image = _module('compressai.zoo.image', model_architectures=models.model_architectures,
                             bmshj2018_factorized=models.bmshj2018_factorized,
                             bmshj2018_hyperprior=models.bmshj2018_hyperprior)
    # import-only names used by sc2bench/transforms/codec.py (PIL / BPG / VTM codecs: off the path)
    root.transforms = _module('compressai.transforms', __path__=[])
    root.transforms.functional = _module('compressai.transforms.functional', rgb2ycbcr=_unavailable('rgb2ycbcr'),
                                         ycbcr2rgb=_unavailable('ycbcr2rgb'))
    root.utils = _module('compressai.utils', __path__=[])
    root.utils.bench = _module('compressai.utils.bench', __path__=[])
    root.utils.bench.codecs = _module('compressai.utils.bench.codecs', run_command=_unavailable('run_command'))
    return root
