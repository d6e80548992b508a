"""sc2bench_b200 -- B200-native (sm_100a) implementation of sc2bench's supervised-compression bottleneck path.

The directory is called `sc2-benchmark_b200/` (repo layout contract); it is imported as `sc2bench_b200` through the
thin alias package next to it.  Importing the package loads (and, if needed, builds) libsc2b200.so: the CUDA
library is the product, there is no CPU fallback for the hot path.
"""
import os as _os

# Throughput mode keeps 10-20 CUDA streams busy (pipeline.py).  With the default of 8 hardware queues, streams share queues and a
# coder kernel that waits for its batch's g_a blocks the transforms queued behind it (21-41 k images/s from run to run; 43 k
# every run with 32).  The driver reads this when it creates the context; a value the user has set is left alone.
_os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')

from . import _native  # noqa: E402

_native.load()  # fail loudly right here if the native library is missing and cannot be built

from . import ops  # noqa: E402
from .backbone import (AnalyzableModule, FeatureExtractionBackbone, FileSizeAnalyzer, SplittableResNet,  # noqa: E402,F401
                       UpdatableBackbone, check_if_updatable, get_backbone, splittable_resnet)
from .bottleneck import (LAYER_CLASS_DICT, BaseBottleneck, EntropyBottleneckLayer, FPBasedResNetBottleneck,  # noqa: E402,F401
                         MSHPBasedResNetBottleneck,
                         SHPBasedResNetBottleneck, get_layer, register_layer_class, register_layer_func)
from .entropy_models import EntropyBottleneck, EntropyModel, GaussianConditional  # noqa: E402,F401
from .layers import GDN, GDN1  # noqa: E402,F401
from .models import (CompressionModel, FactorizedPrior, ScaleHyperprior, bmshj2018_factorized,  # noqa: E402,F401
                     bmshj2018_hyperprior, get_scale_table, update_registered_buffers)

from . import backbone, pipeline, wrapper  # noqa: E402,F401
from .pipeline import CodecPipeline  # noqa: E402,F401
from .wrapper import (COMPRESSAI_DICT, WRAPPER_CLASS_DICT, AdaptivePad, EntropicClassifier,  # noqa: E402,F401
                      NeuralInputCompressionClassifier, get_compression_model, redesign_model, register_compressai_model)

__version__ = '0.1.0'
