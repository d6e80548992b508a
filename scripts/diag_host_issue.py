"""Where does the HOST time of a pipelined step go?  Repeats the timed loop of bench.py (config 2) several times in one process
and prints, per repetition, loop time, back-pressure wait and the top host functions (cProfile) -- the slow runs of the
pipelined bench (1 in 4) are host-bound: 3-8 ms per step in the issue loop instead of 0.95."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
import torch
import sc2bench_b200 as s2
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
steps = 40
dev = torch.device('cuda:0')
torch.manual_seed(0)
layer = s2.get_layer('FPBasedResNetBottleneck', num_bottleneck_channels=24, num_target_channels=256).eval()
layer.update(); layer.to(dev)
xs = [torch.randn(256, 3, 224, 224, device=dev) for _ in range(2)]
with torch.inference_mode():
    pipe = s2.CodecPipeline(layer, depth=8, max_ahead=4)
    def run(n):
        last = None
        for i in range(n):
            last = pipe.submit(xs[i & 1]) or last
        return last
    run(22); torch.cuda.synchronize()
    for rep in range(reps):
        pr = cProfile.Profile()
        w0 = pipe.wait_s
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0 = torch.cuda.memory_stats(dev).get('num_device_alloc', 0)
        e0.record()
        t0 = time.perf_counter()
        pr.enable()
        run(steps)
        pr.disable()
        t1 = time.perf_counter()
        for r in pipe.drain():
            pass
        torch.cuda.current_stream().wait_stream(pipe.transform_stream)
        e1.record(); torch.cuda.synchronize()
        wait = pipe.wait_s - w0
        print('rep %d: gpu %.2f ms/step, host loop %.2f ms/step of which back-pressure wait %.2f, allocs %d' % (
            rep, e0.elapsed_time(e1) / steps, (t1 - t0) * 1e3 / steps, wait * 1e3 / steps,
            torch.cuda.memory_stats(dev).get('num_device_alloc', 0) - a0), flush=True)
        if (t1 - t0 - wait) * 1e3 / steps > 2.0 or rep == 0:
            s = io.StringIO()
            pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(12)
            print('\n'.join(l for l in s.getvalue().splitlines()[4:24]), flush=True)
        run(10); torch.cuda.synchronize()
