"""Input-compression wrapper and its model registry, mirroring `sc2bench.models.wrapper` / `sc2bench.models.registry`.

  - NeuralInputCompressionClassifier.forward  <- sc2bench/models/wrapper.py:80-135
      pre_transform -> compression_model.compress(x) -> analyze -> decompress(**obj)['x_hat'] -> post_transform -> classifier
  - COMPRESSAI_DICT, register_compressai_model, get_compressai_model, get_compression_model
                                              <- sc2bench/models/registry.py:12-29, 58-105
  - AdaptivePad                               <- sc2bench/transforms/misc.py:106-154 (defines the padded codec input, e.g. 224 -> 256)
  - EntropicClassifier                        <- sc2bench/models/wrapper.py:196-264 (EntropyBottleneckLayer after a stage of the
      classifier; the 38 fine-tuning configs), with `redesign_model` restating the slice of torchdistill it needs
`compress()` / `decompress()` of the registered models run on libsc2b200.so (models.py).
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F
from torch import nn

from .backbone import AnalyzableModule, UpdatableBackbone
from .bottleneck import EntropyBottleneckLayer
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


def redesign_model(org_model, model_config, model_label='', model_type='original'):
    """The slice of torchdistill.models.util.redesign_model the wrappers use (sc2bench/models/wrapper.py:172-177,231-235):
    `frozen_modules` paths get requires_grad False; a non-empty `sequential` list of (dotted) module paths of `org_model`
    becomes an nn.Sequential of those modules in that order; otherwise the original model is returned."""
    def get_module(path):
        module = org_model
        for name in path.split('.'):
            module = getattr(module, name)
        return module

    for path in model_config.get('frozen_modules', list()):
        for param in get_module(path).parameters():
            param.requires_grad = False
    module_paths = model_config.get('sequential', list())
    if not isinstance(module_paths, list) or len(module_paths) == 0:
        return org_model
    if any(path.startswith('+') for path in module_paths):
        raise NotImplementedError('adaptation modules (`+name`) are not part of the bottleneck path')
    return nn.Sequential(OrderedDict((path.replace('.', '__'), get_module(path)) for path in module_paths))


@register_wrapper_class
class EntropicClassifier(UpdatableBackbone):
    """Classifier with an EntropyBottleneckLayer dropped after `encoder` (a prefix of the classifier's own modules).
    Once updated and in eval mode the features go through compress -> analyze -> decompress, i.e. the coder kernels
    on a latent with one CDF row per feature channel (e.g. 256 x 56 x 56 after layer1)."""

    def __init__(self, classification_model, encoder_config, compression_model_kwargs, decoder_config, classifier_config,
                 analysis_config=None, **kwargs):
        analysis_config = analysis_config or dict()
        super().__init__(analysis_config.get('analyzer_configs', list()))
        self.analyzes_after_compress = analysis_config.get('analyzes_after_compress', False)
        self.entropy_bottleneck = EntropyBottleneckLayer(**compression_model_kwargs)
        self.encoder = nn.Identity() if encoder_config.get('ignored', False) \
            else redesign_model(classification_model, encoder_config, model_label='encoder')
        self.decoder = nn.Identity() if decoder_config.get('ignored', False) \
            else redesign_model(classification_model, decoder_config, model_label='decoder')
        self.classifier = redesign_model(classification_model, classifier_config, model_label='classification')

    def forward(self, x):
        x = self.encoder(x)
        if self.bottleneck_updated and not self.training:
            x = self.entropy_bottleneck.compress(x)
            if self.analyzes_after_compress:
                self.analyze(x)
            x = self.entropy_bottleneck.decompress(**x)
        else:
            x, _ = self.entropy_bottleneck(x)
        x = self.decoder(x)
        x = torch.flatten(x, 1)
        return self.classifier(x)

    def update(self):
        self.entropy_bottleneck.update()
        self.bottleneck_updated = True

    def load_state_dict(self, state_dict, **kwargs):
        eb_state_dict = OrderedDict()
        for key in list(state_dict.keys()):
            if key.startswith('entropy_bottleneck.'):
                eb_state_dict[key.replace('entropy_bottleneck.', '', 1)] = state_dict.pop(key)
        super().load_state_dict(state_dict, strict=False)
        self.entropy_bottleneck.load_state_dict(eb_state_dict)

    def get_aux_module(self, **kwargs):
        return self.entropy_bottleneck
