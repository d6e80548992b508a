"""Runs the bottleneck's transforms (g_a and g_s) a few times at batch 256 for ncu captures.
    ncu --set full --import-source on -k regex:<pattern> -s <skip> -c <n> -o gpurun_out/prof python scripts/prof_transforms.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import sc2bench_b200 as s2

B = int(os.environ.get('PROF_BATCH', '256'))
dev = torch.device('cuda:0')
torch.manual_seed(0)
layer = s2.get_layer('FPBasedResNetBottleneck', num_bottleneck_channels=24, num_target_channels=256).eval()
layer.update()
layer.to(dev)
torch.manual_seed(1)
x = torch.randn(B, 3, 224, 224, device=dev)
with torch.inference_mode():
    for it in range(int(os.environ.get('PROF_ITERS', '3'))):
        sym = layer.analyze_to_symbols(x)
        lat = sym.float()
        out = layer.synthesize(lat)
    torch.cuda.synchronize()
print('ok', tuple(sym.shape), tuple(out.shape))
