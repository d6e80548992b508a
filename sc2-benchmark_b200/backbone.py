"""Splittable backbone and analysis hooks, mirroring `sc2bench.models.backbone` / `sc2bench.analysis`.

  - AnalyzableModule, FileSizeAnalyzer, get_analyzer   <- sc2bench/analysis.py:24-148
  - UpdatableBackbone, check_if_updatable              <- sc2bench/models/backbone.py:47-87
  - SplittableResNet, splittable_resnet, get_backbone  <- sc2bench/models/backbone.py:175-276, 658-698, 894-909
`SplittableResNet.forward` is the orchestrator of the hot path: encode -> analyze -> decode -> layer2..fc.
The ResNet tail is downstream of the path and stays torchvision / cuDNN (SURVEY.md 2.3 last table row).
"""
import logging
import pickle
import sys
from collections import OrderedDict

import numpy as np
import torch
from torch import nn
from torchvision import models
from torchvision.ops import misc as misc_nn_ops

from .bottleneck import get_layer
from .models import CompressionModel

logger = logging.getLogger('sc2bench_b200')
ANALYZER_CLASS_DICT = dict()
BACKBONE_CLASS_DICT = dict()
BACKBONE_FUNC_DICT = dict()


def register_analysis_class(cls):
    ANALYZER_CLASS_DICT[cls.__name__] = cls
    return cls


def register_backbone_class(cls):
    BACKBONE_CLASS_DICT[cls.__name__] = cls
    return cls


def register_backbone_func(func):
    BACKBONE_FUNC_DICT[func.__name__] = func
    return func


def get_binary_object_size(x, unit_size=1024):
    """torchdistill.common.file_util.get_binary_object_size: size of the pickled object."""
    return sys.getsizeof(pickle.dumps(x)) / unit_size


class BaseAnalyzer(object):
    def analyze(self, *args, **kwargs):
        raise NotImplementedError()

    def summarize(self):
        raise NotImplementedError()

    def clear(self):
        raise NotImplementedError()


@register_analysis_class
class FileSizeAnalyzer(BaseAnalyzer):
    """Records the pickled size of each compressed object in B / KB / MB."""
    UNIT_DICT = {'B': 1, 'KB': 1024, 'MB': 1024 * 1024}

    def __init__(self, unit='KB', **kwargs):
        self.unit = unit
        self.unit_size = self.UNIT_DICT[unit]
        self.kwargs = kwargs
        self.file_size_list = list()

    def analyze(self, compressed_obj):
        self.file_size_list.append(get_binary_object_size(compressed_obj, unit_size=self.unit_size))

    def summarize(self):
        sizes = np.array(self.file_size_list)
        logger.info('Bottleneck size [{}]: mean {} std {} for {} samples'.format(self.unit, sizes.mean(), sizes.std(), len(sizes)))

    def clear(self):
        self.file_size_list.clear()


def get_analyzer(cls_name, **kwargs):
    if cls_name not in ANALYZER_CLASS_DICT:
        return None
    return ANALYZER_CLASS_DICT[cls_name](**kwargs)


class AnalyzableModule(nn.Module):
    """A module whose compressed intermediate representation can be inspected by analyzers."""

    def __init__(self, analyzer_configs=None):
        super().__init__()
        self.analyzers = [get_analyzer(cfg['key'], **cfg['kwargs']) for cfg in (analyzer_configs or list())]
        self.activated_analysis = False

    def forward(self, *args, **kwargs):
        raise NotImplementedError()

    def activate_analysis(self):
        self.activated_analysis = True

    def deactivate_analysis(self):
        self.activated_analysis = False

    def analyze(self, compressed_obj):
        if not self.activated_analysis:
            return
        for analyzer in self.analyzers:
            analyzer.analyze(compressed_obj)

    def summarize(self):
        for analyzer in self.analyzers:
            analyzer.summarize()

    def clear_analysis(self):
        for analyzer in self.analyzers:
            analyzer.clear()


class UpdatableBackbone(AnalyzableModule):
    def __init__(self, analyzer_configs=None):
        super().__init__(analyzer_configs)
        self.bottleneck_updated = False

    def update(self, **kwargs):
        raise NotImplementedError()

    def get_aux_module(self, **kwargs):
        raise NotImplementedError()


def check_if_updatable(model):
    return isinstance(model, UpdatableBackbone)


@register_backbone_class
class FeatureExtractionBackbone(UpdatableBackbone):
    """Runs the children of `model` in order up to the last requested layer and returns {out_name: feature}
    (detection / segmentation bodies; sc2bench/models/backbone.py:90-172).  The child named `analyzable_layer_key` is the
    bottleneck: once updated and in eval mode it goes through encode -> analyze -> decode, i.e. the hot path."""

    def __init__(self, model, return_layer_dict, analyzer_configs, analyzes_after_compress=False, analyzable_layer_key=None):
        children = OrderedDict(model.named_children())
        if not set(return_layer_dict).issubset(children):
            raise ValueError('return_layer_dict are not present in model')
        super().__init__(analyzer_configs)
        wanted = {str(k) for k in return_layer_dict}
        for name, module in children.items():  # layers after the last requested one are pruned
            self.add_module(name, module)
            wanted.discard(name)
            if not wanted:
                break
        self.return_layer_dict = return_layer_dict
        self.analyzable_layer_key = analyzable_layer_key
        self.analyzes_after_compress = analyzes_after_compress

    def forward(self, x):
        out = OrderedDict()
        for key, module in self.named_children():
            if key == self.analyzable_layer_key and self.bottleneck_updated and not self.training:
                compressed = module.encode(x)
                if self.analyzes_after_compress:
                    self.analyze(compressed)
                x = module.decode(**compressed)
            else:
                x = module(x)
            if key in self.return_layer_dict:
                out[self.return_layer_dict[key]] = x
        return out

    def check_if_updatable(self):
        key = self.analyzable_layer_key
        return key is not None and key in self._modules and isinstance(self._modules[key], CompressionModel)

    def update(self):
        if self.analyzable_layer_key is None:
            return
        if not self.check_if_updatable():
            raise KeyError(f'`analyzable_layer_key` ({self.analyzable_layer_key}) does not name an updatable bottleneck in {type(self).__name__}')
        self._modules[self.analyzable_layer_key].update()
        self.bottleneck_updated = True

    def get_aux_module(self, **kwargs):
        return self._modules[self.analyzable_layer_key] if self.check_if_updatable() else None


@register_backbone_class
class SplittableResNet(UpdatableBackbone):
    """ResNet whose stem + layer1 are replaced by a bottleneck layer (encoder | entropy bottleneck | decoder)."""

    def __init__(self, bottleneck_layer, resnet_model, inplanes=None, skips_avgpool=True, skips_fc=True,
                 pre_transform=None, analysis_config=None, short_module_names=None):
        analysis_config = analysis_config or dict()
        if short_module_names is None:
            kept = {'layer2', 'layer3', 'layer4'}
        else:
            kept = set(short_module_names)
        super().__init__(analysis_config.get('analyzer_configs', list()))
        self.pre_transform = pre_transform
        self.analyzes_after_compress = analysis_config.get('analyzes_after_compress', False)
        self.bottleneck_layer = bottleneck_layer
        self.layer2 = resnet_model.layer2 if 'layer2' in kept else None
        self.layer3 = resnet_model.layer3 if 'layer3' in kept else None
        self.layer4 = resnet_model.layer4 if 'layer4' in kept else None
        pool = resnet_model.global_pool if hasattr(resnet_model, 'global_pool') else resnet_model.avgpool
        self.avgpool = None if skips_avgpool else pool
        self.fc = None if skips_fc else resnet_model.fc
        self.inplanes = resnet_model.inplanes if inplanes is None else inplanes

    def forward(self, x):
        if self.pre_transform is not None:
            x = self.pre_transform(x)
        if self.bottleneck_updated and not self.training:
            compressed = self.bottleneck_layer.encode(x)
            if self.analyzes_after_compress:
                self.analyze(compressed)
            x = self.bottleneck_layer.decode(**compressed)
        else:
            x = self.bottleneck_layer(x)
        for stage in (self.layer2, self.layer3, self.layer4):
            if stage is not None:
                x = stage(x)
        if self.avgpool is None:
            return x
        x = self.avgpool(x)
        if self.fc is None:
            return x
        return self.fc(torch.flatten(x, 1))

    def update(self):
        self.bottleneck_layer.update()
        self.bottleneck_updated = True

    def load_state_dict(self, state_dict, **kwargs):
        """Tail loaded non-strictly, bottleneck through CompressionModel.load_state_dict (resizes the CDF buffers)."""
        prefix = 'bottleneck_layer.'
        bottleneck_state = OrderedDict((k[len(prefix):], state_dict.pop(k)) for k in list(state_dict.keys()) if k.startswith(prefix))
        super().load_state_dict(state_dict, strict=False)
        self.bottleneck_layer.load_state_dict(bottleneck_state)

    def get_aux_module(self, **kwargs):
        return self.bottleneck_layer if isinstance(self.bottleneck_layer, CompressionModel) else None


@register_backbone_func
def splittable_resnet(bottleneck_config, resnet_name='resnet50', inplanes=None, skips_avgpool=True, skips_fc=True,
                      pre_transform=None, analysis_config=None, org_model_ckpt_file_path_or_url=None,
                      org_ckpt_strict=True, short_module_names=None, **resnet_kwargs):
    bottleneck_layer = get_layer(bottleneck_config['key'], **bottleneck_config['kwargs'])
    if resnet_kwargs.pop('norm_layer', '') == 'FrozenBatchNorm2d':
        resnet_kwargs['norm_layer'] = misc_nn_ops.FrozenBatchNorm2d
    resnet_model = models.__dict__[resnet_name](**resnet_kwargs)
    if org_model_ckpt_file_path_or_url is not None:
        ckpt = torch.load(org_model_ckpt_file_path_or_url, map_location='cpu')
        resnet_model.load_state_dict(ckpt['model'] if 'model' in ckpt else ckpt, strict=org_ckpt_strict)
    return SplittableResNet(bottleneck_layer, resnet_model, inplanes, skips_avgpool, skips_fc, pre_transform,
                            analysis_config, short_module_names=short_module_names)


def get_backbone(cls_or_func_name, **kwargs):
    if cls_or_func_name in BACKBONE_CLASS_DICT:
        return BACKBONE_CLASS_DICT[cls_or_func_name](**kwargs)
    if cls_or_func_name in BACKBONE_FUNC_DICT:
        return BACKBONE_FUNC_DICT[cls_or_func_name](**kwargs)
    return None
