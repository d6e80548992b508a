"""CPU, build container only (needs /root/reference): the UNMODIFIED reference model code runs on top of this package
installed under the CompressAI names (INTEGRATION.md route 2).  torchdistill / timm come from the import stubs in
oracle/shim (they are absent from this image).  No GPU compute here: construction, update(), tables, state-dict layout."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = '/root/reference'

SCRIPT = r'''
import sys, warnings
warnings.simplefilter('ignore')
sys.path.insert(0, %(root)r)
import sc2bench_b200 as s2
import sc2bench_b200.compat as compat
compat.install_as_compressai()
sys.path.insert(0, %(ref)r)
sys.path.append(%(shim)r)          # only torchdistill / timm stubs are taken from here: compressai is already installed
import compressai
assert getattr(compressai, '__sc2bench_b200__', False)
import torch
from sc2bench.models.backbone import splittable_resnet, SplittableResNet
from sc2bench.models.layer import FPBasedResNetBottleneck, get_layer
from sc2bench.models.registry import COMPRESSAI_DICT
torch.manual_seed(0)
model = splittable_resnet(bottleneck_config={'key': 'FPBasedResNetBottleneck', 'kwargs': {'num_bottleneck_channels': 24, 'num_target_channels': 256}},
                          resnet_name='resnet50', skips_avgpool=False, skips_fc=False, weights=None)
bl = model.bottleneck_layer
assert type(bl) is FPBasedResNetBottleneck and type(bl).__module__ == 'sc2bench.models.layer'      # the reference's class ...
assert isinstance(bl, s2.CompressionModel) and type(bl.entropy_bottleneck) is s2.EntropyBottleneck  # ... built from ours
assert type(bl.encoder[1]) is s2.GDN1
model.eval(); model.update()
assert model.bottleneck_updated and bl.updated
assert model.get_aux_module() is bl
import numpy as np
g = np.load(%(gold)r)
assert (bl.entropy_bottleneck._quantized_cdf.numpy() == g['eb24_cdf']).all()
# the reference's eval branch reaches our hot path, which refuses CPU tensors loudly
try:
    model(torch.randn(1, 3, 224, 224))
    raise SystemExit('expected a RuntimeError: no CPU fallback')
except RuntimeError as e:
    assert 'CUDA' in str(e)
# the product's own mirror has the identical state-dict layout
mine = s2.splittable_resnet(bottleneck_config={'key': 'FPBasedResNetBottleneck', 'kwargs': {'num_bottleneck_channels': 24, 'num_target_channels': 256}},
                            resnet_name='resnet50', skips_avgpool=False, skips_fc=False, weights=None)
mine.update()
assert list(mine.state_dict().keys()) == list(model.state_dict().keys())
mine.load_state_dict(model.state_dict())
assert 'bmshj2018_factorized' in COMPRESSAI_DICT and COMPRESSAI_DICT['bmshj2018_factorized'] is s2.bmshj2018_factorized
print('DROPIN-OK')
'''


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason='/root/reference is only present in the build container')
def test_unmodified_reference_models_build_on_this_package():
    code = SCRIPT % {'root': ROOT, 'ref': REFERENCE, 'shim': os.path.join(ROOT, 'oracle', 'shim'),
                     'gold': os.path.join(ROOT, 'tests', 'golden', 'rans_cases.npz')}
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and 'DROPIN-OK' in out.stdout, out.stdout + out.stderr
