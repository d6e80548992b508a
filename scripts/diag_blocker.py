"""Do resident coder blocks slow the (persistent) convolution kernels down?  Times transform-only steps on 4 streams while
background streams keep decode kernels resident.  python scripts/diag_blocker.py [n_background_streams]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import sc2bench_b200 as s2  # noqa: E402

nbg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device('cuda:0')
torch.manual_seed(0)
layer = s2.get_layer('FPBasedResNetBottleneck').eval()
layer.update()
layer.to(dev)
eb = layer.entropy_bottleneck
xs = [torch.randn(256, 3, 224, 224, device=dev) for _ in range(2)]
with torch.inference_mode():
    sym0 = layer.analyze_to_symbols(xs[0])
    streams0 = eb.compress_symbols(sym0, spatial=sym0[0, 0].numel())
    shape = tuple(sym0.shape[-2:])
    lat0 = eb.decompress_packed(streams0, shape)
    torch.cuda.synchronize()
    fg = [torch.cuda.Stream(device=dev) for _ in range(4)]
    bg = [torch.cuda.Stream(device=dev) for _ in range(nbg)]

    def transforms(i):
        layer.analyze_to_symbols(xs[i & 1])
        return layer.synthesize(lat0)

    def run(n, with_bg, bg_fn):
        main = torch.cuda.current_stream()
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record(main)
        if with_bg:
            for r in range(3):
                for b in bg:
                    b.wait_event(e0)
                    with torch.cuda.stream(b):
                        bg_fn()
        for i in range(n):
            w = fg[i % len(fg)]
            w.wait_event(e0)
            with torch.cuda.stream(w):
                transforms(i)
        e1s = []
        for w in fg:
            e = torch.cuda.Event(enable_timing=True)
            e.record(w)
            e1s.append(e)
        torch.cuda.synchronize()
        return max(e0.elapsed_time(e) for e in e1s) / n

    dec = lambda: eb.decompress_packed(streams0, shape)
    enc = lambda: eb.compress_symbols(sym0, spatial=sym0[0, 0].numel())
    run(8, True, dec)
    print('transform-only steps, 4 streams, 16 steps: %.3f ms/step' % run(16, False, dec))
    print('  + %d background streams of decode kernels: %.3f ms/step' % (nbg, run(16, True, dec)))
    print('  + %d background streams of encode kernels: %.3f ms/step' % (nbg, run(16, True, enc)))
