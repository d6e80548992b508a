"""Does a resident coder block keep a convolution CTA off its SM?  Background: ONE decode (or encode) launch with G blocks
(G * 32 streams, ~10 ms); foreground: the g_s kernels on another stream while it runs.  If a coder block blocks its SM the
foreground slows down like 148 / (148 - G); if they co-reside it does not.  python scripts/diag_coresident.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import sc2bench_b200 as s2  # noqa: E402

dev = torch.device('cuda:0')
torch.manual_seed(0)
layer = s2.get_layer('FPBasedResNetBottleneck').eval()
layer.update()
layer.to(dev)
eb = layer.entropy_bottleneck
with torch.inference_mode():
    x = torch.randn(256, 3, 224, 224, device=dev)
    sym = layer.analyze_to_symbols(x)
    shape = tuple(sym.shape[-2:])
    lat = eb.decompress_packed(eb.compress_symbols(sym, spatial=sym[0, 0].numel()), shape)
    fg, bg = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    which = sys.argv[1] if len(sys.argv) > 1 else 'gs'
    fg_fn = (lambda: layer.synthesize(lat)) if which == 'gs' else (lambda: layer.analyze_to_symbols(x))
    for G in (0, 8, 37, 74, 111, 148):
        for kind in ('decode', 'encode'):
            if G == 0 and kind == 'encode':
                continue
            if G:
                big = sym[:1].expand(G * 32, -1, -1, -1).contiguous()
                st = eb.compress_symbols(big, spatial=big[0, 0].numel())
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            if G:
                with torch.cuda.stream(bg):
                    b0 = torch.cuda.Event(enable_timing=True); b0.record()
                    if kind == 'decode':
                        eb.decompress_packed(st, shape)
                    else:
                        eb.compress_symbols(big, spatial=big[0, 0].numel())
                    b1 = torch.cuda.Event(enable_timing=True); b1.record()
            with torch.cuda.stream(fg):
                fg_fn()  # gives the background a head start (its blocks are resident first)
                e0.record()
                for _ in range(2):
                    fg_fn()
                e1.record()
            torch.cuda.synchronize()
            print('G=%3d %s background (%.1f ms): foreground %s %.3f ms' % (
                G, kind if G else 'no', b0.elapsed_time(b1) if G else 0.0, which, e0.elapsed_time(e1) / 2))
            if G:
                del big, st
