"""Where does the pipelined step go?  (run on the GPU box: python scripts/diag_overlap.py [inflight] [steps])

Times, with CUDA events, N steps of (a) the whole path, (b) the transforms only (g_a -> symbols, g_s on a fixed latent),
(c) the coder only (rANS encode + pack + decode of fixed symbols), each with 1 and with `inflight` batches in flight, and
prints the per-kernel mean duration under overlap next to the serial one (event time on the launching stream, so under
overlap it includes waiting for SMs)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import sc2bench_b200 as s2  # noqa: E402

inflight = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
dev = torch.device('cuda:0')
torch.manual_seed(0)
layer = s2.get_layer('FPBasedResNetBottleneck').eval()
layer.update()
layer.to(dev)
xs = [torch.randn(256, 3, 224, 224, device=dev) for _ in range(2)]
eb = layer.entropy_bottleneck

with torch.inference_mode():
    sym0 = layer.analyze_to_symbols(xs[0])
    streams0 = eb.compress_symbols(sym0, spatial=sym0[0, 0].numel())
    shape = tuple(sym0.shape[-2:])
    lat0 = eb.decompress_packed(streams0, shape)

    def full(i):
        st, sh = layer.encode_packed(xs[i & 1])
        return layer.decode_packed(st, sh)

    def transforms(i):
        layer.analyze_to_symbols(xs[i & 1])
        return layer.synthesize(lat0)

    def coder(i):
        st = eb.compress_symbols(sym0, spatial=sym0[0, 0].numel())
        return eb.decompress_packed(st, shape)

    def encode_only(i):
        return eb.compress_symbols(sym0, spatial=sym0[0, 0].numel())

    def decode_only(i):
        return eb.decompress_packed(streams0, shape)

    host_ms = [0.0]

    def run(fn, n_streams, n):
        workers = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
        main = torch.cuda.current_stream()

        def go(k):
            import time
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record(main)
            t0 = time.perf_counter()
            for i in range(k):
                w = workers[i % n_streams]
                w.wait_event(e0)
                with torch.cuda.stream(w):
                    fn(i)
            host_ms[0] = (time.perf_counter() - t0) * 1e3 / k
            for w in workers:
                main.wait_stream(w)
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record(main)
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / k
        go(max(n_streams, 3))
        return go(n)

    for name, fn in (('full', full), ('transforms', transforms), ('coder', coder), ('encode', encode_only), ('decode', decode_only)):
        a = run(fn, 1, 6)
        b = run(fn, inflight, steps)
        hb = host_ms[0]
        c = run(fn, 2 * inflight, 2 * steps)
        print('%-11s serial %7.3f ms/step   %2d in flight %7.3f (host issue %.3f)   %2d in flight %7.3f (host issue %.3f)'
              % (name, a, inflight, b, hb, 2 * inflight, c, host_ms[0]))

    # per-kernel times: serial vs overlapped
    res = {}
    for label, n_streams in (('serial', 1), ('overlap', inflight)):
        s2.ops.profile_kernels('all')
        run(full, n_streams, steps)
        r = s2.ops.profile_results()
        s2.ops.profile_kernels(None)
        res[label] = {k: sum(v[len(v) // 2:]) / len(v[len(v) // 2:]) for k, v in r.items()}
    print('%-44s %9s %9s' % ('kernel', 'serial', 'overlap'))
    tot = [0.0, 0.0]
    for k in res['serial']:
        print('%-44s %9.3f %9.3f' % (k, res['serial'][k], res['overlap'].get(k, float('nan'))))
        tot[0] += res['serial'][k]
        tot[1] += res['overlap'].get(k, 0.0)
    print('%-44s %9.3f %9.3f' % ('sum', tot[0], tot[1]))
