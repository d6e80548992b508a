"""Per-CTA trace of batches in flight (sc2_trace_start): which SM runs what, when.
python scripts/diag_trace.py [streams] [steps] [streams|pipeline]
  streams   one CUDA stream per batch, everything of a batch on its stream (lane-per-stream coder)
  pipeline  CodecPipeline: transforms on one stream in a fixed order, coders on per-batch streams"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import sc2bench_b200 as s2  # noqa: E402

n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 48
mode = sys.argv[3] if len(sys.argv) > 3 else 'streams'
dev = torch.device('cuda:0')
torch.manual_seed(0)
layer = s2.get_layer('FPBasedResNetBottleneck').eval()
layer.update()
layer.to(dev)
xs = [torch.randn(256, 3, 224, 224, device=dev) for _ in range(2)]
lib = s2._native.load()
with torch.inference_mode():
    def full(i):
        st, sh = layer.encode_packed(xs[i & 1])
        return layer.decode_packed(st, sh)

    workers = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
    layer.entropy_bottleneck.coder_layout = 'lanes'
    pipe = s2.CodecPipeline(layer, depth=n_streams) if mode == 'pipeline' else None

    def go_pipeline(k):
        main = torch.cuda.current_stream()
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for i in range(k):
            pipe.submit(xs[i & 1])
        pipe.drain()
        main.wait_stream(pipe.transform_stream)
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record(main)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k

    def go(k):
        if pipe is not None:
            return go_pipeline(k)
        main = torch.cuda.current_stream()
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for i in range(k):
            w = workers[i % n_streams]
            w.wait_event(e0)
            with torch.cuda.stream(w):
                full(i)
        for w in workers:
            main.wait_stream(w)
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record(main)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k

    go(n_streams)
    cap = 1 << 20
    buf = torch.zeros(16 + 32 * cap, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    s2._native.check(lib.sc2_trace_start(ctypes.c_void_p(buf.data_ptr()), buf.numel()), 'trace_start')
    ms = go(steps)
    lib.sc2_trace_stop()
    raw = buf.cpu().numpy()
n = int(raw[:4].view(np.uint32)[0])
rec = raw[16:16 + 32 * min(n, cap)].view(np.dtype([('t0', '<u8'), ('t1', '<u8'), ('kind', '<i4'), ('sm', '<i4'), ('aux', '<i4'), ('bid', '<i4')]))
print('mode %s, %d batches in flight: %.3f ms/step, %d trace records' % (mode, n_streams, ms, n))
t_begin, t_end = rec['t0'].min(), rec['t1'].max()
lo = t_begin + (t_end - t_begin) // 4
hi = t_end - (t_end - t_begin) // 4
res = 2000  # ns per bin
nb = int((hi - lo) // res)
conv = np.zeros((148, nb), dtype=bool)
coder = np.zeros((148, nb), dtype=np.int16)
tiles_rate = np.zeros(148)
for r in rec:
    a, b = (int(r['t0']) - int(lo)) // res, (int(r['t1']) - int(lo)) // res
    a, b = max(a, 0), min(b, nb - 1)
    if b < a:
        continue
    if r['kind'] <= 3:
        conv[r['sm'], a:b + 1] = True
    else:
        coder[r['sm'], a:b + 1] += 1
has_coder = coder > 0
print('steady window %.1f ms; SM-time with a coder block resident: %.1f%% (mean %.1f blocks resident, max %d on one SM)'
      % ((hi - lo) / 1e6, 100 * has_coder.mean(), coder.sum(0).mean(), coder.max()))
print('conv CTA resident | coder on the SM : %.1f%% of that SM-time' % (100 * conv[has_coder].mean() if has_coder.any() else -1))
print('conv CTA resident | no coder        : %.1f%%' % (100 * conv[~has_coder].mean()))
k = rec[(rec['kind'] <= 3) & (rec['t0'] >= lo) & (rec['t1'] <= hi)]
dur = (k['t1'] - k['t0']).astype(np.float64) / 1e3
print('conv CTAs in window: %d, zero-tile CTAs: %.1f%%, mean CTA lifetime %.1f us' % (len(k), 100 * (k['aux'] == 0).mean(), dur.mean()))
# tiles per SM-time for SMs with / without a coder block at the CTA's start
for kind, name in ((1, 'conv_tc'), (2, 'conv_split'), (3, 'conv_first')):
    kk = k[k['kind'] == kind]
    if not len(kk):
        continue
    a = ((kk['t0'] - lo) // res).astype(np.int64).clip(0, nb - 1)
    with_c = has_coder[kk['sm'], a]
    d = (kk['t1'] - kk['t0']).astype(np.float64) / 1e3
    def rate(m):
        return kk['aux'][m].sum() / max(d[m].sum(), 1e-9)
    print('%-10s CTAs starting next to a coder block: %5.1f%%; tiles/us with coder %.4f, without %.4f'
          % (name, 100 * with_c.mean(), rate(with_c), rate(~with_c)))
c = rec[(rec['kind'] >= 4)]
for kind, name in ((4, 'encode'), (5, 'decode')):
    cc = c[c['kind'] == kind]
    if len(cc):
        print('%s blocks: %d, mean duration %.2f ms' % (name, len(cc), ((cc['t1'] - cc['t0']).astype(np.float64) / 1e6).mean()))
# per kernel kind: CTA lifetimes (a persistent CTA lives as long as its kernel)
for kind, name in ((1, 'conv_tc'), (2, 'conv_split'), (3, 'conv_first')):
    kk = k[k['kind'] == kind]
    if len(kk):
        d = (kk['t1'] - kk['t0']).astype(np.float64) / 1e3
        live = d[kk['aux'] > 0]
        print('%-10s CTAs %6d  mean lifetime of working CTAs %7.1f us (p10 %7.1f, p90 %7.1f)  tiles %8d  sum of lifetimes %8.1f ms'
              % (name, len(kk), live.mean(), np.percentile(live, 10), np.percentile(live, 90), kk['aux'].sum(), d.sum() / 1e3))
# coarse timeline: per 0.5 ms, mean number of SMs with a conv CTA and number of resident coder blocks
step = 250  # bins of `res` ns -> 0.5 ms
print('t(ms)  conv-SMs  coder-blocks  enc dec   (first 40 rows of the steady window)')
enc_cnt = np.zeros(nb, dtype=np.int16)
dec_cnt = np.zeros(nb, dtype=np.int16)
for r in rec[rec['kind'] >= 4]:
    a, b = max((int(r['t0']) - int(lo)) // res, 0), min((int(r['t1']) - int(lo)) // res, nb - 1)
    if b >= a:
        (enc_cnt if r['kind'] == 4 else dec_cnt)[a:b + 1] += 1
for i in range(0, min(nb, 60 * step), step):
    print('%5.1f  %7.1f  %7.1f   %5.1f %5.1f' % (i * res / 1e6, conv[:, i:i + step].sum(0).mean(), coder[:, i:i + step].sum(0).mean(),
                                             enc_cnt[i:i + step].mean(), dec_cnt[i:i + step].mean()))
