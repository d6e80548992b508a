#!/usr/bin/env python
"""bench.py -- images/s of the supervised-compression bottleneck path (encode + rANS + decode) on B200.

    python bench.py --gpus 1 --steps 10 --warmup 3            # this repo's CUDA path
    python bench.py --impl reference --steps 3 --warmup 1     # the reference's CPU path (restated oracle) on host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                # one rank per GPU, weak scaling

Workload = BASELINE.json configs[1]: Entropic Student ResNet-50 bottleneck (FPBasedResNetBottleneck, 24 bottleneck /
256 target channels, factorized-prior EntropyBottleneck), random init (seed 0), synthetic 3x224x224 images, batch 256
PER GPU (weak scaling: images are independent units, no data-path collective; the only collective is one counter
all-reduce after the timed region).  A step = bottleneck_layer.encode + bottleneck_layer.decode over one batch.

One JSON line on stdout (rank 0):
  value     images/s, inputs resident in HBM, device-timed (CUDA events), max over ranks
  e2e       images/s through the public plugin API with HOST buffers: pinned images -> H2D -> encode() -> list[bytes]
            on the host (D2H) -> decode(strings) (H2D) -> per-image feature means read back (D2H)
  roofline  dominant kernel: algorithmic FLOPs per launch / its CUDA-event time inside the timed region, vs MEASURED_PEAKS.json
  cpu_baseline  the restated reference (oracle/, "port": CompressAI is not installable here) on the host cores, bounded sample
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# keep stdout clean for the single JSON line: NCCL's version / debug banner goes to a file
os.environ.setdefault('NCCL_DEBUG_FILE', '/tmp/sc2b200_nccl.%h.%p.log')
# One hardware queue per CUDA stream for the 10-20 streams of the batch pipeline: with the default 8, streams share queues and
# a coder kernel waiting for its batch's g_a blocks the transforms queued behind it (measured: 21-41 k images/s from run to
# run with 8 connections, 43 k every run with 32).  Read by the driver when the context is created, i.e. before torch starts.
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')

METRIC = 'images/s encode+rANS+decode @224^2 (FPBasedResNetBottleneck, Entropic Student ResNet-50)'
UNIT = 'images/s'
IMG = (3, 224, 224)
LATENT = (24, 55, 55)
# algorithmic FLOPs (2*MAC) per image
FLOPS = {'conv2d_f32[3->96,k5,s2]': 180.6e6, 'gdn_f32[96]': 231.2e6, 'conv2d_f32[96->48,k5,s2]': 722.5e6,
         'gdn_f32[48]': 14.5e6, 'conv2d_f32[48->24,k2,s1]': 27.9e6, 'conv2d_f32[24->512,k2,s1]': 308.3e6,
         'gdn_f32[512,inv]': 1644.2e6, 'conv2d_f32[512->256,k2,s1]': 3172.0e6, 'gdn_f32[256,inv]': 396.5e6,
         'conv2d_f32[256->256,k2,s1]': 1644.2e6,
         # tensor-core g_s (tags carry the PADDED input channels; FLOPs are the algorithmic ones)
         'tc_conv[64->512,k2,m0]': 308.3e6, 'tc_conv[512->512,k1,m2]': 1644.2e6, 'tc_conv[512->256,k2,m0]': 3172.0e6,
         'tc_conv[256->256,k1,m2]': 396.5e6, 'tc_conv[256->256,k2,m1]': 1644.2e6}
# algorithmic HBM bytes per image of the HBM-bound kernels (every kernel reads its input once and writes its
# output once, 4 bytes per activation -- the split fp16 (hi, lo) pairs of g_a are 4 bytes per value as well)
BYTES = {'rans_encode': 290400 + 49240, 'rans_decode': 49240 + 290400, 'rans_pack': 2 * 49240,
         'nchw_to_nhwc_f16': 290400 + 55 * 55 * 64 * 2,
         'tc_first[3->96,k5,s2]': 602112 + 4816896,              # K1 with fused im2col: image in, x1 out
         'patchify_split': 602112 + 4 * 12544 * 80,            # (unfused route) image in, im2col patches out
         'tc_split[80->96,k1,s1,m0]': 4 * 12544 * 80 + 4816896,  # K1: patches in, x1 out
         'tc_split[96->96,k1,s1,m1]': 2 * 4816896,               # GDN1(96): x1 in, y1 out
         'tc_split[96->48,k5,s2,m0]': 4816896 + 602112,          # K3
         'tc_split[48->48,k1,s1,m1]': 2 * 602112,                # GDN1(48)
         'tc_split[48->24,k2,s1,m2]': 602112 + 290400}           # K5 + quantise: y2 in, int32 symbols out
PATH_FLOPS_PER_IMAGE = 8.342e9
PATH_BYTES_PER_IMAGE = 34.85e6


def load_traffic():
    """DRAM bytes per launch per kernel from the newest ncu --set full capture summarised under profiles/ (batch 256)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, 'profiles', '*_traffic.json')))
    if not files:
        return {}, None
    with open(files[-1]) as f:
        return json.load(f), os.path.basename(files[-1])


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {'hbm_gbs': p['hbm_gbs'], 'tflops_burst': p['bf16_tflops'], 'tflops_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']),
                'source': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'tflops_burst': 1590.0, 'tflops_sustained': 1400.0, 'source': 'fallback (B200_PROFILING.md)'}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu_index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, smmax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smmax.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(smmax) if smmax else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def build_product_layer(device):
    import torch
    import sc2bench_b200 as s2
    torch.manual_seed(0)
    layer = s2.get_layer('FPBasedResNetBottleneck', num_bottleneck_channels=24, num_target_channels=256)
    layer.eval()
    layer.update()
    return layer.to(device)


def cpu_reference_run(n_images, steps, warmup, state_dict=None):
    """Times the restated reference path (oracle/) on the host cores: torch CPU convs (all threads) + the per-sample
    CompressAI coder loop through Python lists (as EntropyModel.compress/decompress does).  Returns images/s etc."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import torch
    import ref_models
    torch.manual_seed(0)
    layer = ref_models.build_fp_bottleneck(3, 24, 256)
    if state_dict is not None:
        layer.load_state_dict(state_dict)
    layer.eval()
    layer.update()
    torch.manual_seed(1)
    x = torch.randn(n_images, *IMG)
    times, nbytes = [], 0
    with torch.inference_mode():
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            obj = layer.encode(x)
            out = layer.decode(**obj)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
            nbytes = sum(len(s) for s in obj['strings'][0])
    total = sum(times)
    return {'images_per_s': n_images * len(times) / total, 'ms_per_step': 1e3 * total / len(times), 'threads': torch.get_num_threads(),
            'cores': os.cpu_count(), 'bytes_per_image': nbytes / n_images, 'out_shape': list(out.shape)}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    n = args.cpu_images
    r = cpu_reference_run(n, args.steps, args.warmup)
    sample = '%d images of 3x224x224 per step (the B200 arm runs 256 per GPU per step); restated reference ' \
             '(oracle/shim compressai restatement + C rANS, torch CPU convs); compressai itself is not installable here' % n
    line = {'impl': 'reference', 'metric': METRIC, 'value': r['images_per_s'], 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'entropic-student-resnet50 FPBasedResNetBottleneck encode+decode, 3x224x224, random init',
                       'images_per_step': n, 'device': 'cpu'},
            'cpu_baseline': {'value': r['images_per_s'], 'unit': UNIT, 'cores': r['threads'], 'kind': 'port', 'sample': sample},
            'e2e': {'value': r['images_per_s'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=256, help='images per GPU per step')
    ap.add_argument('--cpu-images', type=int, default=8, help='images per step of the CPU baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--e2e-threads', type=int, default=0, help='host threads (one CUDA stream each) driving the e2e steps')
    ap.add_argument('--max-ahead', type=int, default=4, help='batches the host may run ahead of the GPU beyond the pipeline depth')
    ap.add_argument('--inflight', type=int, default=8, help='depth of the batch pipeline: the serial rANS chains of up to this many batches overlap the convolutions of the others')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'b200':
        args.warmup = 3  # timing rule: at least 3 warm-up steps

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        return run_reference_arm(args, rank, world)

    import torch
    import torch.distributed as dist
    import sc2bench_b200 as s2
    from sc2bench_b200 import parallel

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)')
    rank, world, local_rank = parallel.init_distributed('nccl')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    peaks = load_peaks()
    layer = build_product_layer(device)
    B = args.batch
    n_sym = LATENT[0] * LATENT[1] * LATENT[2]

    # two distinct input batches (154 MB each, larger than the 126 MB L2) alternate between steps
    gen = torch.Generator(device='cpu').manual_seed(1 + rank)
    host_inputs = [torch.randn(B, *IMG, generator=gen).pin_memory() for _ in range(2)]
    dev_inputs = [h.to(device) for h in host_inputs]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(i):
        streams, shape = layer.encode_packed(dev_inputs[i & 1])
        return streams, layer.decode_packed(streams, shape)

    # Batches are independent and a batch's coder is one serial rANS chain per image (milliseconds on a handful of warps),
    # so steps are software-pipelined (sc2bench_b200/pipeline.py): every g_a / g_s on ONE transform stream in a fixed order,
    # g_a(i + depth) ahead of g_s(i), and the coder of each batch on its own stream in the layout that occupies one SM.
    # Every step does all of its work inside the timed region; the region ends when the pipeline has drained.
    pipe = s2.pipeline.CodecPipeline(layer, depth=max(1, args.inflight), max_ahead=args.max_ahead)

    def run_steps(n, first=0):
        main = torch.cuda.current_stream()
        start = torch.cuda.Event(enable_timing=True)
        start.record(main)
        last = None
        for i in range(first, first + n):
            last = pipe.submit(dev_inputs[i & 1]) or last
        for r in pipe.drain():
            last = r
        main.wait_stream(pipe.transform_stream)  # the last g_s, hence every batch, has finished
        last.wait(main)
        stop = torch.cuda.Event(enable_timing=True)
        stop.record(main)
        return start, stop, (last.streams, last.features)

    # ---- device-resident throughput ("value") -------------------------------------------------
    with torch.inference_mode():
        # warm-up: at least W steps and at least one step per batch stream (each stream has its own allocator pool)
        n_warm = max(args.warmup, 2 * len(pipe.batch_streams) + args.max_ahead)
        run_steps(n_warm)
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        launches0 = s2.ops.STATS['launches']
        allocs0 = torch.cuda.memory_stats(device).get('num_device_alloc', 0)
        t_issue = time.perf_counter()
        e0, e1, (streams, out) = run_steps(args.steps, first=n_warm)
        t_issue = (time.perf_counter() - t_issue) * 1e3
        barrier()
        launches = s2.ops.STATS['launches'] - launches0
        allocs = torch.cuda.memory_stats(device).get('num_device_alloc', 0) - allocs0
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop()
        total_bytes = streams.total_bytes()
        pipe.close()  # back to one batch at a time
        # per-kernel accounting: a second, SERIAL timed pass over the SAME kernels (one batch in flight, CUDA events around
        # every launch on the launching stream) -- with batches overlapping, a kernel's event time would include waiting for
        # SMs held by others.  The coder keeps the layout of the timed region (lane per stream).
        layer.entropy_bottleneck.coder_layout = 'lanes'
        n_prof = max(1, min(args.steps, 5))
        device_step(0)
        torch.cuda.synchronize()
        s2.ops.profile_kernels('all')
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for i in range(n_prof):
            device_step(i)
        p1.record()
        torch.cuda.synchronize()
        serial_ms = p0.elapsed_time(p1) / n_prof
        prof = s2.ops.profile_results()
        s2.ops.profile_kernels(None)
        # ... and the latency of ONE batch with the low-latency coder layout (a warp per stream), for reference
        layer.entropy_bottleneck.coder_layout = None
        device_step(0)
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record()
        for i in range(n_prof):
            device_step(i)
        l1.record()
        torch.cuda.synchronize()
        latency_ms = l0.elapsed_time(l1) / n_prof

    t = torch.tensor([ms], dtype=torch.float64, device=device)
    counters = parallel.EvalCounters(device)
    counters.add(images=B * args.steps, bytes=total_bytes * args.steps, symbols=B * n_sym * args.steps)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    counters.all_reduce()  # the path's only collective: one small counter vector per evaluation
    ms = float(t.item())
    c = counters.as_dict()
    value = c['images'] / (ms / 1e3)

    # ---- end-to-end through the public API with host buffers ("e2e") ---------------------------
    # Each step is the reference-facing call sequence on HOST data: pinned images -> H2D -> layer.encode(x) (returns the
    # contract object with real `bytes` on the host) -> layer.decode(**obj) (bytes -> H2D -> features) -> a per-image
    # result read back.  `e2e_threads` host threads each drive their own CUDA stream (like the reference's
    # nn.DataParallel replicas are host threads), so the PCIe copies and the host-side bytes handling of one batch overlap
    # the GPU work of another; every step still runs entirely inside the timed region.
    e2e = None
    if not args.no_e2e:
        import concurrent.futures

        def e2e_step(i):
            with torch.inference_mode(), torch.cuda.stream(e2e_streams[i % len(e2e_streams)]):
                x = host_inputs[i & 1].to(device, non_blocking=True)
                obj = layer.encode(x)                       # {'strings': [list[bytes]], 'shape'}: bitstreams land on the host
                feat = layer.decode(**obj)                  # list[bytes] -> device -> features
                res = feat.mean(dim=(1, 2, 3))
                torch.cuda.current_stream().synchronize()   # (a blocking .cpu() would hold a driver lock while it waits)
                res = res.cpu()                             # per-image result read back
            return obj, res

        # default: 12 host threads per GPU (they mostly wait on the GPU with the GIL released), fewer when ranks share few cores
        n_thr = args.e2e_threads if args.e2e_threads > 0 else min(12, max(6, 2 * (os.cpu_count() or 8) // max(world, 1)))
        e2e_streams = [torch.cuda.Stream(device=device) for _ in range(n_thr)]
        # the threads' transforms share one stream (in arrival order); a thread queues a transform only once its own stream
        # has produced the input, so a batch that is still copying or coding never holds up the others
        layer.use_transform_stream(True, host_wait=True)
        with concurrent.futures.ThreadPoolExecutor(max_workers=n_thr) as pool:
            list(pool.map(e2e_step, range(max(args.warmup, 2 * n_thr))))
            barrier()
            e2e_allocs0 = torch.cuda.memory_stats(device).get('num_device_alloc', 0)
            t0 = time.perf_counter()
            results = list(pool.map(e2e_step, range(2 * n_thr, 2 * n_thr + args.steps)))
            torch.cuda.synchronize()
            e2e_ms = (time.perf_counter() - t0) * 1e3  # host wall clock: host work is part of this contract
            barrier()
        layer.use_transform_stream(None)
        obj, res = results[-1]
        te = torch.tensor([e2e_ms], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        stream_bytes = sum(len(s) for s in obj['strings'][0])
        e2e = {'value': world * B * args.steps / (float(te.item()) / 1e3), 'unit': UNIT,
               'h2d_bytes_per_step': B * IMG[0] * IMG[1] * IMG[2] * 4 + stream_bytes + 8 * (B + 1),
               'd2h_bytes_per_step': stream_bytes + 8 * (B + 1) + 4 + B * 4, 'host_threads': n_thr,
               'cudaMalloc_calls_in_timed_region': torch.cuda.memory_stats(device).get('num_device_alloc', 0) - e2e_allocs0,
               'ms_per_step': float(te.item()) / args.steps}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel accounting and the roofline of the dominant kernel ---------------------------
    step_ms = serial_ms  # shares are relative to the serial step the kernel times were taken in
    traffic, traffic_src = load_traffic()
    kernels = []
    for tag, times in prof.items():
        avg = sum(times) / len(times)
        per_step = sum(times) / n_prof
        k = {'kernel': tag, 'launches_per_step': len(times) / n_prof, 'avg_launch_ms': avg, 'share_of_step': per_step / step_ms}
        if tag in FLOPS:
            k.update(bound='tensor', achieved=FLOPS[tag] * B / (avg / 1e3) / 1e12, unit='TFLOP/s', peak=peaks['tflops_sustained'])
        elif tag in BYTES:
            k.update(bound='hbm', achieved=BYTES[tag] * B / (avg / 1e3) / 1e9, unit='GB/s', peak=peaks['hbm_gbs'])
        if 'achieved' in k:
            k['frac'] = k['achieved'] / k['peak']
        if tag in traffic:  # measured DRAM bytes of one launch (ncu), scaled to this batch size
            k['traffic'] = traffic[tag]['dram_bytes_per_launch'] * B / traffic[tag].get('batch', 256)
        k['stream'] = 'batch' if tag.startswith('rans_') else 'transform'
        if tag in ('rans_encode', 'rans_decode'):
            k['symbols_per_s_per_stream'] = n_sym / (avg / 1e3)
            k['symbols_per_s_aggregate'] = n_sym * B / (avg / 1e3)
            k['note'] = ('serial rANS state chain per stream (SURVEY.md H1): latency-bound, neither roofline applies.  Lane-per-stream '
                         'layout: the %d streams of a batch are ONE block of %d warps on one SM, running next to the transforms of '
                         'the other batches in flight -- off the critical path of the pipelined step' % (B, (B + 31) // 32))
        else:
            k['share_of_pipelined_step'] = per_step / (ms / args.steps)
        kernels.append(k)
    kernels.sort(key=lambda k: -k['share_of_step'])
    roofline = None
    if kernels:
        top = kernels[0]
        roofline = {'kernel': top['kernel'], 'bound': top.get('bound'), 'achieved': top.get('achieved'), 'peak': top.get('peak'),
                    'unit': top.get('unit'), 'frac': top.get('frac'), 'traffic': top.get('traffic'), 'traffic_source': traffic_src,
                    'peak_source': peaks['source'] + (', bf16 sustained' if top.get('bound') == 'tensor' else ', copy bandwidth'),
                    'avg_launch_ms': top['avg_launch_ms'], 'share_of_step': top['share_of_step'], 'note': top.get('note')}
        keys = ('kernel', 'bound', 'achieved', 'unit', 'peak', 'frac', 'traffic', 'avg_launch_ms', 'share_of_step', 'share_of_pipelined_step')
        crit = [k for k in kernels if k['stream'] == 'transform']
        if crit and crit[0] is not top:  # the kernel that bounds the pipelined step: largest on the transform stream
            roofline['critical_path_kernel'] = {kk: crit[0].get(kk) for kk in keys}
        gemm = [k for k in kernels if k.get('bound') == 'tensor']
        if gemm and gemm[0] is not top:
            roofline['dominant_gemm'] = {kk: gemm[0].get(kk) for kk in keys}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        sd = {k: v.detach().cpu() for k, v in layer.state_dict().items()}
        r = cpu_reference_run(args.cpu_images, steps=2, warmup=1, state_dict=sd)
        cpu_baseline = {'value': r['images_per_s'], 'unit': UNIT, 'cores': r['threads'], 'kind': 'port',
                        'sample': '%d images per step x 2 steps of the same workload; restated reference (oracle/: CompressAI '
                                  'restatement, torch CPU convs on %d threads, per-sample Python-list coder loop + C rANS)'
                                  % (args.cpu_images, r['threads'])}

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': n_warm,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32-grade (g_a: split-f16 operands, 3 tensor-core passes, f32 accumulate) / f16 operands with f32 accumulate (g_s) / u64 (coder)',
            'data': 'synthetic',
            'config': {'workload': 'configs[1]: entropic-student-resnet50 FPBasedResNetBottleneck(24,256) encode+rANS+decode, '
                                   '3x224x224, random init, batch %d per GPU' % B,
                       'images_per_gpu_per_step': B, 'global_images_per_step': B * world, 'symbols_per_image': n_sym,
                       'l2_policy': 'two alternating 154 MB input batches (> 126 MB L2); activations are GBs per step',
                       'parallelism': 'dp%d (batch sharded, one counter all-reduce per evaluation)' % world,
                       'batches_in_flight': args.inflight,
                       'schedule': 'software pipeline: transforms on one stream, g_a(i + depth) ahead of g_s(i); coders on per-batch streams'},
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu_baseline,
            'serial_ms_per_step': serial_ms, 'one_batch_latency_ms': latency_ms,
            'host_issue_ms_per_step': t_issue / args.steps, 'cudaMalloc_calls_in_timed_region': allocs,
            'kernel_accounting': 'serial pass of %d steps after the timed region, same kernels (CUDA events per launch); shares are of '
                                 'that serial step, as in the ncu launch list; one_batch_latency_ms uses the warp-per-stream coder' % n_prof,
            'kernels': kernels,
            'bytes_per_image': c['bytes_per_image'], 'bits_per_symbol': c['bits_per_symbol'],
            'path_tflops': value * PATH_FLOPS_PER_IMAGE / 1e12, 'path_hbm_gbs_algorithmic': value * PATH_BYTES_PER_IMAGE / 1e9}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
