"""Coder-only timing at the bench shape (256 streams x 72,600 symbols): python scripts/diag_coder.py [reps] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import sc2bench_b200 as s2  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device('cuda:0')
torch.manual_seed(0)
layer = s2.get_layer('FPBasedResNetBottleneck').eval()
layer.update()
layer.to(dev)
eb = layer.entropy_bottleneck
with torch.inference_mode():
    x = torch.randn(batch, 3, 224, 224, device=dev)
    sym = layer.analyze_to_symbols(x)
    shape = tuple(sym.shape[-2:])
    n = sym[0].numel()

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, r

    ms_e, streams = timed(lambda: eb.compress_symbols(sym, spatial=sym[0, 0].numel()))
    ms_d, lat = timed(lambda: eb.decompress_packed(streams, shape))
    med = eb._get_medians().detach().reshape(1, -1, 1, 1)
    assert torch.equal(lat, sym.float() + med), 'round trip'
    clk = 1.9e9
    print('batch %d  n %d  bytes/stream %.0f' % (batch, n, streams.total_bytes() / batch))
    print('encode %.3f ms (%.0f cycles/symbol/stream)   decode %.3f ms (%.0f cycles/symbol/stream)'
          % (ms_e, ms_e * 1e-3 * clk / n, ms_d, ms_d * 1e-3 * clk / n))
