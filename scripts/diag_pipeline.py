"""Explicit software pipeline: all transforms on ONE stream in a fixed order (g_a(i + depth) before g_s(i)), each batch's coder
on a side stream.  python scripts/diag_pipeline.py [depth] [steps] [side_streams]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import sc2bench_b200 as s2  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 48
n_side = int(sys.argv[3]) if len(sys.argv) > 3 else depth + 1
dev = torch.device('cuda:0')
torch.manual_seed(0)
layer = s2.get_layer('FPBasedResNetBottleneck').eval()
layer.update()
layer.to(dev)
eb = layer.entropy_bottleneck
xs = [torch.randn(256, 3, 224, 224, device=dev) for _ in range(2)]
with torch.inference_mode():
    conv = torch.cuda.Stream(device=dev)
    side = [torch.cuda.Stream(device=dev) for _ in range(n_side)]

    def run(n):
        main = torch.cuda.current_stream()
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record(main)
        conv.wait_event(e0)
        for s in side:
            s.wait_event(e0)
        pending = {}

        def front(i):
            with torch.cuda.stream(conv):
                sym = layer.analyze_to_symbols(xs[i & 1])
                ev = torch.cuda.Event()
                ev.record(conv)
            s = side[i % n_side]
            s.wait_event(ev)
            with torch.cuda.stream(s):
                sym.record_stream(s)
                st = eb.compress_symbols(sym, spatial=sym[0, 0].numel())
                lat = eb.decompress_packed(st, tuple(sym.shape[-2:]))
                done = torch.cuda.Event()
                done.record(s)
            pending[i] = (lat, done, s)

        def back(i):
            lat, done, s = pending.pop(i)
            conv.wait_event(done)
            with torch.cuda.stream(conv):
                lat.record_stream(conv)
                return layer.synthesize(lat)

        for i in range(n + depth):
            if i < n:
                front(i)
            if i >= depth:
                out = back(i - depth)
        main.wait_stream(conv)
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record(main)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    run(8)
    for d in (depth,):
        print('depth %d, %d side streams, %d steps: %.3f ms/step' % (depth, n_side, steps, run(steps)))
