import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, sc2bench_b200 as s2
dev = torch.device('cuda:0')
torch.manual_seed(0)
layer = s2.get_layer('FPBasedResNetBottleneck').eval(); layer.update(); layer.to(dev)
x_host = torch.randn(256, 3, 224, 224).pin_memory()
def t(fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3, r
with torch.inference_mode():
    ms, x = t(lambda: x_host.to(dev, non_blocking=True)); print('H2D 154MB: %.2f ms (%.1f GB/s)' % (ms, 154.1 / ms))
    ms, (streams, shape) = t(lambda: layer.encode_packed(x)); print('encode_packed: %.2f ms' % ms)
    ms, strings = t(lambda: layer.encode_packed(x)[0].tolist()); print('encode_packed + tolist: %.2f ms' % ms)
    ms, ps = t(lambda: s2.ops.PackedStreams.from_list(strings, dev)); print('from_list (H2D bytes): %.2f ms' % ms)
    ms, feat = t(lambda: layer.decode_packed(ps, shape)); print('decode_packed: %.2f ms' % ms)
    ms, feat = t(lambda: layer.decode([strings], shape)); print('decode(list[bytes]): %.2f ms' % ms)
    ms, r = t(lambda: feat.mean(dim=(1, 2, 3)).cpu()); print('mean + D2H: %.2f ms' % ms)
    ms, r = t(lambda: layer.decode(**layer.encode(x_host.to(dev, non_blocking=True))).mean(dim=(1, 2, 3)).cpu()); print('full e2e step: %.2f ms' % ms)
