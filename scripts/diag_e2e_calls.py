import concurrent.futures, os, sys, time, threading, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, sc2bench_b200 as s2
n_thr = int(sys.argv[1]) if len(sys.argv) > 1 else 6
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
dev = torch.device('cuda:0'); torch.manual_seed(0)
layer = s2.get_layer('FPBasedResNetBottleneck').eval(); layer.update(); layer.to(dev)
host_inputs = [torch.randn(256, 3, 224, 224).pin_memory() for _ in range(2)]
streams = [torch.cuda.Stream(device=dev) for _ in range(n_thr)]
layer.use_transform_stream(True, host_wait=True)
T = collections.defaultdict(list)
def timed(name, fn):
    def w(*a, **k):
        t = time.perf_counter(); r = fn(*a, **k); T[name].append(time.perf_counter() - t); return r
    return w
torch.cuda.Stream.wait_stream = timed('Stream.wait_stream', torch.cuda.Stream.wait_stream)
torch.cuda.Stream.synchronize = timed('Stream.synchronize', torch.cuda.Stream.synchronize)
torch.cuda.Event.record = timed('Event.record', torch.cuda.Event.record)
_to = torch.Tensor.to
torch.Tensor.to = timed('Tensor.to', _to)
torch.Tensor.cpu = timed('Tensor.cpu', torch.Tensor.cpu)
torch.Tensor.item = timed('Tensor.item', torch.Tensor.item)
torch.Tensor.copy_ = timed('Tensor.copy_', torch.Tensor.copy_)
torch.empty = timed('torch.empty', torch.empty)
torch.zeros = timed('torch.zeros', torch.zeros)
lib = s2._native.load()
for name in ('sc2_tc_conv_nhwc', 'sc2_tc_split_conv', 'sc2_tc_first_layer', 'sc2_rans_encode_batch', 'sc2_rans_decode_batch', 'sc2_rans_pack'):
    setattr(lib, name, timed(name, getattr(lib, name)))
def step(i):
    with torch.inference_mode(), torch.cuda.stream(streams[i % n_thr]):
        x = host_inputs[i & 1].to(dev, non_blocking=True)
        obj = layer.encode(x); feat = layer.decode(**obj)
        r = feat.mean(dim=(1, 2, 3)); torch.cuda.current_stream().synchronize(); return r.cpu()
with concurrent.futures.ThreadPoolExecutor(max_workers=n_thr) as pool:
    list(pool.map(step, range(2 * n_thr))); torch.cuda.synchronize()
    for v in T.values(): v.clear()
    ms0 = torch.cuda.memory_stats()
    t0 = time.perf_counter(); list(pool.map(step, range(steps))); torch.cuda.synchronize()
    total = (time.perf_counter() - t0) * 1e3
print('%d threads: %.2f ms/step' % (n_thr, total / steps))
ms1 = torch.cuda.memory_stats()
for k in ('num_device_alloc', 'num_device_free', 'num_alloc_retries', 'reserved_bytes.all.current', 'reserved_bytes.all.peak', 'allocated_bytes.all.peak', 'num_sync_all_streams'):
    print('  %-32s %s -> %s' % (k, ms0.get(k), ms1.get(k)))
for k, v in sorted(T.items(), key=lambda kv: -sum(kv[1])):
    print('%-24s calls/step %5.1f  mean %8.3f ms  max %8.3f ms  total/step %7.2f ms' % (k, len(v) / steps, 1e3 * sum(v) / len(v), 1e3 * max(v), 1e3 * sum(v) / steps))
