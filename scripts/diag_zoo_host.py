"""Where the host time of a zoo codec step goes (cProfile of compress_packed + decompress, hyperprior q8, batch 32)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import sc2bench_b200 as s2

dev = torch.device('cuda:0')
torch.manual_seed(0)
arch = sys.argv[1] if len(sys.argv) > 1 else 'bmshj2018_hyperprior'
m = s2.get_compression_model({'key': arch, 'kwargs': {'quality': 8, 'pretrained': False}}, 'cpu').eval()
m.update()
m.to(dev)
x = torch.rand(32, 3, 256, 256, device=dev)


def step():
    strs, shp = m.compress_packed(x)
    return m.decompress(list(strs) if isinstance(strs, tuple) else [strs], shp)['x_hat']


with torch.inference_mode():
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print('host issue %.2f ms/step, with drain %.2f ms/step, launches/step %d' % ((t1 - t0) * 50, (t2 - t0) * 50, 0))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        step()
    pr.disable()
    torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('tottime').print_stats(22)
