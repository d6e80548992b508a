"""Input-compression wrapper and its model registry, mirroring `sc2bench.models.wrapper` / `sc2bench.models.registry`.

  - NeuralInputCompressionClassifier.forward  <- sc2bench/models/wrapper.py:80-135
      pre_transform -> compression_model.compress(x) -> analyze -> decompress(**obj)['x_hat'] -> post_transform -> classifier
  - COMPRESSAI_DICT, register_compressai_model, get_compressai_model, get_compression_model
                                              <- sc2bench/models/registry.py:12-29, 58-105
  - AdaptivePad                               <- sc2bench/transforms/misc.py:106-154 (defines the padded codec input, e.g. 224 -> 256)
`compress()` / `decompress()` of the registered models run on libsc2b200.so (models.py).
"""
import torch
import torch.nn.functional as F
from torch import nn

from .backbone import AnalyzableModule
from .models import bmshj2018_factorized, bmshj2018_hyperprior

COMPRESSAI_DICT = {'bmshj2018_factorized': bmshj2018_factorized, 'bmshj2018_hyperprior': bmshj2018_hyperprior}
WRAPPER_CLASS_DICT = dict()


def register_compressai_model(cls_or_func):
    COMPRESSAI_DICT[cls_or_func.__name__] = cls_or_func
    return cls_or_func


def register_wrapper_class(cls):
    WRAPPER_CLASS_DICT[cls.__name__] = cls
    return cls


def get_compressai_model(compression_model_name, ckpt_file_path=None, updates=False, **compression_model_kwargs):
    model = COMPRESSAI_DICT[compression_model_name](**compression_model_kwargs)
    if ckpt_file_path is not None:
        ckpt = torch.load(ckpt_file_path, map_location='cpu')
        model.load_state_dict(ckpt['model'] if 'model' in ckpt else ckpt)
    if updates:
        model.update()
    return model


def get_compression_model(compression_model_config, device):
    """{'key', 'kwargs', 'update' (default True), 'src_ckpt'} -> updated model on `device`; None passes through."""
    if compression_model_config is None:
        return None
    name = compression_model_config['key']
    if name not in COMPRESSAI_DICT:
        raise ValueError('compression_model_name `{}` is not expected'.format(name))
    model = get_compressai_model(name, compression_model_config.get('src_ckpt', None),
                                 compression_model_config.get('update', True), **compression_model_config['kwargs'])
    return model.to(device)


class AdaptivePad(nn.Module):
    """Pads H and W up to the next multiple of `factor`: right/bottom, or split over both sides for 'equal_side'."""

    def __init__(self, fill=0, padding_position='hw', padding_mode='constant', factor=128, returns_org_patch_size=False):
        super().__init__()
        self.fill, self.padding_position, self.padding_mode = fill, padding_position, padding_mode
        self.factor, self.returns_org_patch_size = factor, returns_org_patch_size

    def forward(self, x):
        height, width = x.shape[-2:]
        pad_h, pad_w = (-height) % self.factor, (-width) % self.factor
        if self.padding_position == 'equal_side':
            padding = (pad_w // 2, pad_w // 2, pad_h // 2, pad_h // 2)  # torchvision's 2-value form pads both sides equally
        else:
            padding = (0, pad_w, 0, pad_h)
        mode = self.padding_mode
        x = F.pad(x, padding, mode=mode, value=self.fill) if mode == 'constant' else F.pad(x, padding, mode=mode)
        return (x, (height, width)) if self.returns_org_patch_size else x


@register_wrapper_class
class NeuralInputCompressionClassifier(AnalyzableModule):
    """Neural image codec in front of a classifier (the input-compression baseline)."""

    def __init__(self, classification_model, pre_transform=None, compression_model=None, uses_cpu4compression_model=False,
                 post_transform=None, analysis_config=None, **kwargs):
        analysis_config = analysis_config or dict()
        super().__init__(analysis_config.get('analyzer_configs', list()))
        self.analyzes_after_pre_transform = analysis_config.get('analyzes_after_pre_transform', False)
        self.analyzes_after_compress = analysis_config.get('analyzes_after_compress', False)
        self.pre_transform = pre_transform
        self.compression_model = compression_model
        self.uses_cpu4compression_model = uses_cpu4compression_model
        self.classification_model = classification_model
        self.post_transform = post_transform

    def use_cpu4compression(self):
        if self.uses_cpu4compression_model and self.compression_model is not None:
            raise RuntimeError('sc2bench_b200 codecs run on CUDA only (no CPU fallback); keep uses_cpu4compression_model False')

    def forward(self, x):
        if self.pre_transform is not None:
            x = self.pre_transform(x)
            if not self.training and self.analyzes_after_pre_transform:
                self.analyze(x)
        if self.compression_model is not None:
            compressed = self.compression_model.compress(x)
            if not self.training and self.analyzes_after_compress:
                self.analyze(compressed)
            x = self.compression_model.decompress(**compressed)
            if isinstance(x, dict):
                x = x['x_hat']
        if self.post_transform is not None:
            x = self.post_transform(x)
        return self.classification_model(x)
