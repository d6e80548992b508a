"""Host-side phase times of the threaded e2e step: python scripts/diag_e2e_threads.py [threads] [steps] [transform_stream 0/1]"""
import concurrent.futures
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import sc2bench_b200 as s2  # noqa: E402

n_thr = int(sys.argv[1]) if len(sys.argv) > 1 else 6
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 36
use_ts = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device('cuda:0')
torch.manual_seed(0)
layer = s2.get_layer('FPBasedResNetBottleneck').eval()
layer.update()
layer.to(dev)
host_inputs = [torch.randn(256, 3, 224, 224).pin_memory() for _ in range(2)]
streams = [torch.cuda.Stream(device=dev) for _ in range(n_thr)]
if use_ts:
    layer.use_transform_stream(True, host_wait=True)
phases = {}
import numpy as np
sub = {'get_h2d': [], 'join': [], 'to': [], 'offs.cpu': [], 'd2h copy': [], 'split': []}
_ops = s2.ops
_orig_get = _ops._PINNED.get_h2d
def _get(nbytes):
    t = time.perf_counter(); r = _orig_get(nbytes); sub['get_h2d'].append(time.perf_counter() - t); return r
_ops._PINNED.get_h2d = _get
_hb = s2._native.hostbytes()
class _HB:
    @staticmethod
    def join(*a):
        t = time.perf_counter(); r = _hb.join(*a); sub['join'].append(time.perf_counter() - t); return r
    @staticmethod
    def split(*a):
        t = time.perf_counter(); r = _hb.split(*a); sub['split'].append(time.perf_counter() - t); return r
s2._native._hostbytes = _HB
_orig_to_host = _ops.PackedStreams._to_host
def _to_host(self):
    t = time.perf_counter(); r = _orig_to_host(self); sub['offs.cpu'].append(time.perf_counter() - t); return r
_ops.PackedStreams._to_host = _to_host


def step(i):
    t = [time.perf_counter()]
    with torch.inference_mode(), torch.cuda.stream(streams[i % n_thr]):
        x = host_inputs[i & 1].to(dev, non_blocking=True)
        t.append(time.perf_counter())
        st, shape = layer.encode_packed(x)
        t.append(time.perf_counter())
        strings = st.tolist()
        t.append(time.perf_counter())
        ps = s2.ops.PackedStreams.from_list(strings, dev)
        t.append(time.perf_counter())
        feat = layer.decode_packed(ps, shape, check_status=True)
        t.append(time.perf_counter())
        res = feat.mean(dim=(1, 2, 3)); torch.cuda.current_stream().synchronize(); res = res.cpu()
        t.append(time.perf_counter())
    return t


with concurrent.futures.ThreadPoolExecutor(max_workers=n_thr) as pool:
    list(pool.map(step, range(2 * n_thr)))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ts = list(pool.map(step, range(steps)))
    torch.cuda.synchronize()
    total = (time.perf_counter() - t0) * 1e3
names = ['h2d issue', 'encode_packed (issue' + (' + wait for copy' if use_ts else '') + ')', 'tolist (wait coder + D2H + bytes)', 'from_list',
         'decode_packed (issue + status/transform wait)', 'mean + cpu (wait g_s)']
print('%d threads, %d steps, transform stream %d: %.2f ms/step (%.0f images/s)' % (n_thr, steps, use_ts, total / steps, 256 * steps / total * 1e3))
for k, nm in enumerate(names):
    d = [(t[k + 1] - t[k]) * 1e3 for t in ts[n_thr:]]
    print('  %-48s %7.2f ms' % (nm, sum(d) / len(d)))
for k, v in sub.items():
    if v:
        print('    sub %-20s %7.2f ms (n=%d)' % (k, 1e3 * sum(v) / len(v), len(v)))
print('  %-48s %7.2f ms' % ('step latency', sum((t[-1] - t[0]) * 1e3 for t in ts[n_thr:]) / len(ts[n_thr:])))
