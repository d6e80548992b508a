"""Poor man's sampling profiler for the threaded e2e loop: where do the host threads spend wall time?"""
import collections
import concurrent.futures
import os
import sys
import threading
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import sc2bench_b200 as s2  # noqa: E402

n_thr = int(sys.argv[1]) if len(sys.argv) > 1 else 6
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
dev = torch.device('cuda:0')
torch.manual_seed(0)
layer = s2.get_layer('FPBasedResNetBottleneck').eval()
layer.update()
layer.to(dev)
host_inputs = [torch.randn(256, 3, 224, 224).pin_memory() for _ in range(2)]
streams = [torch.cuda.Stream(device=dev) for _ in range(n_thr)]
layer.use_transform_stream(True, host_wait=True)


def step(i):
    with torch.inference_mode(), torch.cuda.stream(streams[i % n_thr]):
        x = host_inputs[i & 1].to(dev, non_blocking=True)
        obj = layer.encode(x)
        feat = layer.decode(**obj)
        return feat.mean(dim=(1, 2, 3)).cpu()


hist = collections.Counter()
stop = False


def sampler():
    me = threading.get_ident()
    while not stop:
        for tid, frame in sys._current_frames().items():
            if tid == me:
                continue
            stack = traceback.extract_stack(frame, limit=3)
            key = ' <- '.join('%s:%d %s' % (os.path.basename(f.filename), f.lineno, f.name) for f in reversed(stack))
            hist[key] += 1
        time.sleep(0.0005)


with concurrent.futures.ThreadPoolExecutor(max_workers=n_thr) as pool:
    list(pool.map(step, range(2 * n_thr)))
    torch.cuda.synchronize()
    th = threading.Thread(target=sampler)
    th.start()
    t0 = time.perf_counter()
    list(pool.map(step, range(steps)))
    torch.cuda.synchronize()
    total = (time.perf_counter() - t0) * 1e3
    stop = True
    th.join()
print('%d threads: %.2f ms/step' % (n_thr, total / steps))
tot = sum(hist.values())
for k, v in hist.most_common(22):
    print('%5.1f%%  %s' % (100 * v / tot, k))
