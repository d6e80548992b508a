#!/usr/bin/env python
"""bench.py -- images/s of the supervised-compression bottleneck path (encode + rANS + decode) on B200.

    python bench.py --gpus 1 --steps 20 --warmup 5            # this repo's CUDA path, BASELINE.json configs[1] (the default)
    python bench.py --config 5 --steps 10                     # another BASELINE config (1..5 = SURVEY.md 8d "Config 1..5")
    python bench.py --impl reference --steps 3 --warmup 1     # the reference's CPU path (restated oracle) on the host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W [--scaling strong]   # one rank per GPU

Workloads (`--config`, numbered like SURVEY.md 8d; BASELINE.json `configs` is 0-based, so --config 2 = configs[1]):
  1  Entropic Student ResNet-50 bottleneck (FPBasedResNetBottleneck(24, 256)), batch 1, 3x224x224: latency of one image
  2  the same bottleneck at batch 256 per GPU (weak scaling, default) or a global batch of 256 split over the ranks (--scaling
     strong); steps are software-pipelined (sc2bench_b200/pipeline.py)          <- the metric's configuration, the default
  3  bmshj2018_factorized(quality=8): 3x224x224 -> AdaptivePad(64) -> 3x256x256 -> compress -> decompress
  4  bmshj2018_hyperprior(quality=8), as 3
  5  the Entropic Student bottleneck at COCO shape 3x800x1344 (one 1.6 M-symbol rANS stream per image)
A step = encode + decode (compress + decompress) of one batch.  Images are independent units: no data-path collective; the
only collective is one counter all-reduce after the timed region.

One JSON line on stdout (rank 0):
  value        images/s, inputs resident in HBM, device-timed (CUDA events), max over ranks
  e2e          images/s through the public plugin API with HOST buffers (pinned images -> H2D -> encode() -> list[bytes] on the
               host -> decode(strings) -> per-image result D2H); config 2 feeds uint8 images (device-side ToTensor + Normalize)
               and also reports the fp32-input figure
  roofline     the largest kernel ON THE TRANSFORM STREAM (the critical path of the pipelined step): algorithmic FLOPs and bytes
               of the launch (SURVEY.md 8d conventions, stated by the op itself: ops._launch) / its CUDA-event time in a serial
               accounting pass, against MEASURED_PEAKS.json; the coder is reported as symbols/s/stream (latency-bound)
  cpu_baseline the restated reference (oracle/, "port": CompressAI is not installable here) on the host cores, bounded sample
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
# One hardware queue per CUDA stream for the 10-20 streams of the batch pipeline (sc2bench_b200/__init__.py explains).
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')

UNIT = 'images/s'
METRICS = {
    1: 'images/s encode+rANS+decode @224^2, batch 1 (FPBasedResNetBottleneck, Entropic Student ResNet-50)',
    2: 'images/s encode+rANS+decode @224^2 (FPBasedResNetBottleneck, Entropic Student ResNet-50)',
    3: 'images/s compress+rANS+decompress @224^2 padded to 256^2 (bmshj2018_factorized q8)',
    4: 'images/s compress+rANS+decompress @224^2 padded to 256^2 (bmshj2018_hyperprior q8)',
    5: 'images/s encode+rANS+decode @3x800x1344 (FPBasedResNetBottleneck, COCO shape)',
}
IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
# Kernels whose tensor-core work is three fp16 MMA passes per algorithmic MAC (split fp16 = fp32-grade)
THREE_PASS_PREFIXES = ('tc_split', 'tc_first', 'ga_halo', 'ga_first', 'tcs_conv', 'tcs_gdn', 'tcs_deconv5')


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
    """Samples SM clocks / throttle reasons while the timed region runs: NVML in a thread of this process (8 ranks each forking
    an `nvidia-smi -lms` poller perturbed the 8-GPU runs), nvidia-smi as the fallback."""
    REASONS = ((0x8, 'hw_slowdown'), (0x40, 'hw_thermal_slowdown'), (0x20, 'sw_thermal_slowdown'), (0x4, 'sw_power_cap'))

    def __init__(self, torch_device_index):
        self.idx, self.rows, self.stop_flag, self.thread, self.handle, self.nv = torch_device_index, [], threading.Event(), None, None, None

    def start(self):
        try:
            import pynvml as nv
            import torch
            nv.nvmlInit()
            try:
                uuid = 'GPU-' + str(torch.cuda.get_device_properties(self.idx).uuid)
                self.handle = nv.nvmlDeviceGetHandleByUUID(uuid)
            except Exception:
                vis = os.environ.get('CUDA_VISIBLE_DEVICES')
                phys = int(vis.split(',')[self.idx]) if vis and all(v.strip().isdigit() for v in vis.split(',')) else self.idx
                self.handle = nv.nvmlDeviceGetHandleByIndex(phys)
            self.nv = nv
            self.smmax = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.nv = None
            self._start_smi()

    def _poll(self):
        nv = self.nv
        reasons_fn = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag.is_set():
            try:
                self.rows.append((float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)), int(reasons_fn(self.handle))))
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def _start_smi(self):
        q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + q, '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read_smi(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(',')]
            try:
                mask = sum(bit for (bit, _), v in zip(self.REASONS, r[4:8]) if v.lower().startswith('active'))
                self.rows.append((float(r[1]), mask))
                self.smmax = float(r[2])
            except (ValueError, IndexError):
                continue

    def stop(self):
        if self.thread is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['clock sampling unavailable']}
        time.sleep(0.12)
        self.stop_flag.set()
        if self.nv is None and getattr(self, 'proc', None) is not None:
            self.proc.terminate()
        self.thread.join(timeout=2)
        sm = [r[0] for r in self.rows]
        mask = 0
        for r in self.rows:
            mask |= r[1]
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': getattr(self, 'smmax', None),
                'reasons': sorted(name for bit, name in self.REASONS if mask & bit), 'samples': len(sm),
                'source': 'nvml' if self.nv is not None else 'nvidia-smi'}


# ----------------------------------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------------------------------
def workload_spec(config, batch):
    """(image shape, default batch per GPU, description)"""
    if config == 1:
        return (3, 224, 224), batch or 1, 'configs[0] on the GPU: entropic-student-resnet50 FPBasedResNetBottleneck(24,256) encode+rANS+decode, 3x224x224, batch 1, random init'
    if config == 2:
        return (3, 224, 224), batch or 256, 'configs[1]: entropic-student-resnet50 FPBasedResNetBottleneck(24,256) encode+rANS+decode, 3x224x224, random init'
    if config == 3:
        return (3, 224, 224), batch or 32, 'configs[2]: bmshj2018_factorized(quality=8) compress+decompress, 3x224x224 -> AdaptivePad(64) -> 3x256x256, random init'
    if config == 4:
        return (3, 224, 224), batch or 32, 'configs[3]: bmshj2018_hyperprior(quality=8) compress+decompress, 3x224x224 -> AdaptivePad(64) -> 3x256x256, random init'
    if config == 5:
        return (3, 800, 1344), batch or 1, 'configs[4]: FPBasedResNetBottleneck(24,256) encode+rANS+decode at COCO shape 3x800x1344 (GeneralizedRCNNTransform batching of 800x1333), random init'
    raise SystemExit('--config must be 1..5')


def build_product_model(config, device):
    import torch
    import sc2bench_b200 as s2
    torch.manual_seed(0)
    if config in (1, 2, 5):
        m = s2.get_layer('FPBasedResNetBottleneck', num_bottleneck_channels=24, num_target_channels=256)
        m.eval()
        m.update()
    else:
        m = s2.get_compression_model({'key': 'bmshj2018_factorized' if config == 3 else 'bmshj2018_hyperprior',
                                      'kwargs': {'quality': 8, 'pretrained': False}}, 'cpu')
        m.eval()
    return m.to(device)


def make_step_fns(config, model):
    """(encode(x) -> obj, decode(obj) -> tensor, count_symbols(obj-or-shape)) with the reference's call contract per config."""
    import sc2bench_b200 as s2
    if config in (1, 2, 5):
        return (lambda x: model.encode(x)), (lambda obj: model.decode(**obj))
    pad = s2.AdaptivePad(fill=0, factor=64)
    return (lambda x: model.compress(pad(x))), (lambda obj: model.decompress(**obj)['x_hat'])


def build_oracle_model(config, state_dict=None):
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import torch
    import ref_models
    torch.manual_seed(0)
    if config in (1, 2, 5):
        m = ref_models.build_fp_bottleneck(3, 24, 256)
    else:
        ref_models._import_shim()
        import compressai.zoo as ozoo
        m = getattr(ozoo, 'bmshj2018_factorized' if config == 3 else 'bmshj2018_hyperprior')(quality=8, pretrained=False)
    if state_dict is not None:
        m.load_state_dict(state_dict)
    m.eval()
    m.update()
    return m


def cpu_reference_run(config, n_images, steps, warmup, state_dict=None):
    """Times the restated reference path (oracle/) on the host cores: torch CPU convs (all threads) + the per-sample CompressAI
    coder loop through Python lists (as EntropyModel.compress/decompress does).  Returns images/s etc."""
    import torch
    import torch.nn.functional as F
    shape, _, _ = workload_spec(config, None)
    m = build_oracle_model(config, state_dict)
    torch.manual_seed(1)
    x = torch.randn(n_images, *shape) if config in (1, 2, 5) else torch.rand(n_images, *shape)
    times, nbytes, out = [], 0, None
    with torch.inference_mode():
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            if config in (1, 2, 5):
                obj = m.encode(x)
                out = m.decode(**obj)
            else:
                xp = F.pad(x, (0, (-shape[2]) % 64, 0, (-shape[1]) % 64))
                obj = m.compress(xp)
                out = m.decompress(**obj)['x_hat']
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
            nbytes = sum(len(s) for lst in obj['strings'] for s in lst)
    total = sum(times)
    return {'images_per_s': n_images * len(times) / total, 'ms_per_step': 1e3 * total / len(times), 'threads': torch.get_num_threads(),
            'cores': os.cpu_count(), 'bytes_per_image': nbytes / n_images, 'out_shape': list(out.shape)}


def default_cpu_images(config):
    return {1: 1, 2: 8, 3: 2, 4: 2, 5: 1}[config]


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    n = args.cpu_images or default_cpu_images(args.config)
    shape, b, desc = workload_spec(args.config, args.batch)
    r = cpu_reference_run(args.config, n, args.steps, args.warmup)
    sample = '%d images of %dx%dx%d per step (the B200 arm runs %d per GPU per step); restated reference (oracle/shim compressai ' \
             'restatement + C rANS, torch CPU convs); compressai itself is not installable here' % ((n,) + shape + (b,))
    line = {'impl': 'reference', 'metric': METRICS[args.config], 'value': r['images_per_s'], 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': args.scaling,
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': desc, 'images_per_step': n, 'device': 'cpu'},
            'cpu_baseline': {'value': r['images_per_s'], 'unit': UNIT, 'cores': r['threads'], 'kind': 'port', 'sample': sample},
            'e2e': {'value': r['images_per_s'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# roofline arithmetic
# ----------------------------------------------------------------------------------------------------------------------
def account_kernels(prof, work, n_prof, step_ms, pipelined_ms, peaks, traffic, batch):
    kernels = []
    for tag, times in prof.items():
        avg = sum(times) / len(times)
        per_step = sum(times) / n_prof
        k = {'kernel': tag, 'launches_per_step': len(times) / n_prof, 'avg_launch_ms': avg, 'share_of_step': per_step / step_ms}
        flops, nbytes, symbols = work.get(tag, (None, None, None))
        coder = tag in ('rans_encode', 'rans_decode')
        k['stream'] = 'batch' if tag.startswith('rans_') else 'transform'
        if coder and symbols:
            k['symbols_per_s_per_stream'] = symbols[1] / (avg / 1e3)
            k['symbols_per_s_aggregate'] = symbols[0] * symbols[1] / (avg / 1e3)
            k['streams'], k['symbols_per_stream'] = symbols
            k['note'] = 'serial rANS state chain per stream (SURVEY.md H1): latency-bound, neither roofline applies'
        elif flops or nbytes:
            # the binding roofline of the launch: the larger of (algorithmic FLOPs / tensor peak) and (algorithmic bytes / copy peak)
            t_tensor = (flops or 0.0) / (peaks['tflops_sustained'] * 1e12) * 1e3
            t_hbm = (nbytes or 0.0) / (peaks['hbm_gbs'] * 1e9) * 1e3
            if t_tensor >= t_hbm:
                k.update(bound='tensor', achieved=flops / (avg / 1e3) / 1e12, unit='TFLOP/s', peak=peaks['tflops_sustained'])
            else:
                k.update(bound='hbm', achieved=nbytes / (avg / 1e3) / 1e9, unit='GB/s', peak=peaks['hbm_gbs'])
            k['frac'] = k['achieved'] / k['peak']
            k['algorithmic_flops'], k['algorithmic_bytes'] = flops, nbytes
            k['floor_ms'] = {'tensor': t_tensor, 'hbm': t_hbm}
            if tag.startswith(THREE_PASS_PREFIXES) and flops:
                # executed tensor work of the fp32-grade kernels = 3 MMA passes per algorithmic MAC: the bound they can reach
                k['frac_of_3pass_tensor_bound'] = 3.0 * t_tensor / avg
        if tag in traffic:  # measured DRAM bytes of one launch (ncu), scaled to this batch size
            k['traffic'] = traffic[tag]['dram_bytes_per_launch'] * batch / traffic[tag].get('batch', 256)
        if pipelined_ms and not coder:
            k['share_of_pipelined_step'] = per_step / pipelined_ms
        kernels.append(k)
    kernels.sort(key=lambda k: -k['share_of_step'])
    return kernels


def pick_roofline(kernels, peaks, traffic_src):
    crit = [k for k in kernels if k['stream'] == 'transform' and 'frac' in k]
    if not crit:
        return None
    top = crit[0]
    keys = ('kernel', 'bound', 'achieved', 'unit', 'peak', 'frac', 'traffic', 'avg_launch_ms', 'share_of_step', 'share_of_pipelined_step',
            'algorithmic_flops', 'algorithmic_bytes', 'floor_ms', 'frac_of_3pass_tensor_bound')
    r = {kk: top.get(kk) for kk in keys}
    r['traffic_source'] = traffic_src
    r['peak_source'] = peaks['source'] + (', bf16 sustained' if top.get('bound') == 'tensor' else ', copy bandwidth')
    r['choice'] = 'largest kernel on the transform stream (the critical path of the pipelined step); the coder runs beside it on ' \
                  'per-batch streams and is reported under "coder"'
    fr = [k['frac'] for k in crit]
    r['transform_kernels'] = len(crit)
    r['transform_kernels_at_70pct'] = sum(1 for f in fr if f >= 0.7)
    return r


# ----------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=None,
                    help='timed steps (default 40, config 5: 160; the reference arm: 3).  Config 2 at 200 steps (0.8 s of sustained load) runs into the '
                         'board power cap: 1770 MHz instead of 1965, 64 k instead of 68 k images/s (profiles/r4_pipeline_depth.md)')
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', type=int, default=2, help='workload, numbered like SURVEY.md 8d (2 = BASELINE.json configs[1], the metric)')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: --batch images per GPU; strong: a global batch of --batch images split over the ranks')
    ap.add_argument('--batch', type=int, default=0, help='images per GPU per step (strong scaling: global); 0 = the config default')
    ap.add_argument('--cpu-images', type=int, default=0, help='images per step of the CPU baseline sample (0 = per-config default)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--e2e-threads', type=int, default=0, help='host threads (one CUDA stream each) driving the e2e steps')
    ap.add_argument('--max-ahead', type=int, default=4, help='batches the host may run ahead of the GPU beyond the pipeline depth')
    ap.add_argument('--inflight', type=int, default=None, help='configs 2 / 5: depth of the batch pipeline (default 16 / 32)')
    ap.add_argument('--streams', type=int, default=8, help='configs 3 / 4: CUDA streams the batches round-robin over (1: one batch at a time)')
    ap.add_argument('--coder-sms', type=int, default=None,
                    help='configs 2 / 5: SMs the persistent transform kernels leave to the coder blocks (default 12)')
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 3 if args.impl == 'reference' else (160 if args.config == 5 else 40)  # config 5: 5 x its pipeline depth
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
    cfg = args.config
    shape, B, desc = workload_spec(cfg, args.batch)
    if args.scaling == 'strong':
        lo, hi = parallel.shard_bounds(B, rank, world)  # a global batch of B images, contiguous slices per rank
        B_global, B = B, hi - lo
        if B < 1:
            raise SystemExit('strong scaling: global batch %d is smaller than the world size %d' % (B_global, world))
    else:
        B_global = B * world
    model = build_product_model(cfg, device)
    encode, decode = make_step_fns(cfg, model)
    # config 2 (the headline) and config 5 (COCO shape) are throughput measurements: batches in flight (CodecPipeline); config 5 also
    # reports the latency of one image (`one_batch_latency_ms`)
    pipelined = cfg in (2, 5)
    # configs 3 / 4 (zoo codecs): batches round-robin over a few CUDA streams through the device-resident calls (compress_packed /
    # decompress on PackedStreams), so that one batch's coder chains overlap the transforms of the others
    multistream = cfg in (3, 4) and args.streams > 1
    if args.inflight is None:
        # the serial coder chains of a batch take ~20 ms whatever its size, so the depth that keeps the transform stream busy grows as
        # the per-GPU batch shrinks (strong scaling: 128 / 64 / 32 images per GPU at N = 2 / 4 / 8): profiles/r4_pipeline_depth.md
        args.inflight = min(96, max(16, 16 * 256 // max(B, 1))) if cfg == 2 else 32
    if args.coder_sms is None:
        args.coder_sms = 12

    # inputs: two distinct batches alternate between steps; where a batch is smaller than L2 (126 MB) the L2 is flushed between
    # timed iterations instead (a 256 MB write), each iteration timed by its own event pair
    gen = torch.Generator(device='cpu').manual_seed(1 + rank)
    rnd = torch.randn if cfg in (1, 2, 5) else torch.rand
    input_bytes = B * shape[0] * shape[1] * shape[2] * 4
    # pipelined configs cannot flush L2 between steps: they alternate between enough distinct batches to exceed it (126 MB)
    n_in = max(2, -(-140_000_000 // input_bytes)) if (pipelined or multistream) else 2
    host_inputs = [rnd(B, *shape, generator=gen).pin_memory() for _ in range(n_in)]
    dev_inputs = [h.to(device) for h in host_inputs]
    flush_l2 = input_bytes < 160e6 and not pipelined and not multistream
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=device) if flush_l2 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(i):
        obj = encode(dev_inputs[i % n_in])
        return obj, decode(obj)

    def n_symbols(obj):
        return None

    launches = allocs = 0
    t_issue = 0.0
    extra = {}
    with torch.inference_mode():
        if pipelined:
            # Batches are independent and a batch's coder is one serial rANS chain per image, so steps are software-pipelined
            # (sc2bench_b200/pipeline.py): every g_a / g_s on ONE transform stream, g_a(i + depth) ahead of g_s(i), coders on
            # per-batch streams in the lane-per-stream layout.  Every step does all of its work inside the timed region; the
            # region ends when the pipeline has drained.
            pipe = s2.pipeline.CodecPipeline(model, depth=max(1, args.inflight), max_ahead=args.max_ahead, coder_sms=args.coder_sms)

            def run_steps(n, first=0):
                main_s = torch.cuda.current_stream()
                start = torch.cuda.Event(enable_timing=True)
                start.record(main_s)
                last = None
                for i in range(first, first + n):
                    last = pipe.submit(dev_inputs[i % n_in]) or last
                for r in pipe.drain():
                    last = r
                main_s.wait_stream(pipe.transform_stream)
                last.wait(main_s)
                stop = torch.cuda.Event(enable_timing=True)
                stop.record(main_s)
                return start, stop, last

            # priming (not a warm-up step count: allocator pools of every batch stream get their blocks), then W warm-up steps
            n_prime = 2 * len(pipe.batch_streams) + args.max_ahead
            run_steps(n_prime)
            run_steps(args.warmup, first=n_prime)
            barrier()
            sampler = ClockSampler(local_rank)
            sampler.start()
            launches0 = s2.ops.STATS['launches']
            allocs0 = torch.cuda.memory_stats(device).get('num_device_alloc', 0)
            t0 = time.perf_counter()
            w0 = pipe.wait_s
            e0, e1, last = run_steps(args.steps, first=n_prime + args.warmup)
            t_issue = (time.perf_counter() - t0 - (pipe.wait_s - w0)) * 1e3  # host time issuing work, back-pressure waits excluded
            extra['host_backpressure_wait_ms_per_step'] = (pipe.wait_s - w0) * 1e3 / args.steps
            barrier()
            launches = s2.ops.STATS['launches'] - launches0
            allocs = torch.cuda.memory_stats(device).get('num_device_alloc', 0) - allocs0
            ms = e0.elapsed_time(e1)
            clocks = sampler.stop()
            total_bytes = last.streams.total_bytes()
            sym_per_image = int(torch.tensor(last.shape).prod()) * model.entropy_bottleneck.channels
            pipe.close()
            extra['priming_steps'] = n_prime
            model.entropy_bottleneck.coder_layout = 'throughput'  # the accounting pass keeps the layout of the timed region
        elif multistream:
            pad = s2.AdaptivePad(fill=0, factor=64)
            rr = [torch.cuda.Stream(device=device) for _ in range(args.streams)]

            def packed_step(i):
                strs, shp = model.compress_packed(pad(dev_inputs[i % n_in]))
                return strs, model.decompress(list(strs) if isinstance(strs, tuple) else [strs], shp)['x_hat']

            def run_rr(n, first=0):
                main_s = torch.cuda.current_stream()
                start = torch.cuda.Event(enable_timing=True)
                start.record(main_s)
                last = None
                for st_ in rr:
                    st_.wait_stream(main_s)
                for i in range(first, first + n):
                    with torch.cuda.stream(rr[i % len(rr)]):
                        last = packed_step(i)
                for st_ in rr:
                    main_s.wait_stream(st_)
                stop = torch.cuda.Event(enable_timing=True)
                stop.record(main_s)
                return start, stop, last

            run_rr(max(args.warmup, 2 * len(rr)))
            barrier()
            sampler = ClockSampler(local_rank)
            sampler.start()
            launches0 = s2.ops.STATS['launches']
            allocs0 = torch.cuda.memory_stats(device).get('num_device_alloc', 0)
            t0 = time.perf_counter()
            e0, e1, last = run_rr(args.steps, first=max(args.warmup, 2 * len(rr)))
            t_issue = (time.perf_counter() - t0) * 1e3
            barrier()
            launches = s2.ops.STATS['launches'] - launches0
            allocs = torch.cuda.memory_stats(device).get('num_device_alloc', 0) - allocs0
            ms = e0.elapsed_time(e1)
            clocks = sampler.stop()
            model.entropy_bottleneck.check_faults()
            strs = last[0]
            total_bytes = sum(ps.total_bytes() for ps in (strs if isinstance(strs, tuple) else (strs,)))
            sym_per_image = None
            extra['streams'] = len(rr)
        else:
            for i in range(args.warmup):
                device_step(i)
            barrier()
            sampler = ClockSampler(local_rank)
            sampler.start()
            launches0 = s2.ops.STATS['launches']
            allocs0 = torch.cuda.memory_stats(device).get('num_device_alloc', 0)
            evs = []
            t0 = time.perf_counter()
            for i in range(args.steps):
                if flush_l2:
                    flush_buf.fill_(i & 255)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                obj, out = device_step(i)
                b.record()
                evs.append((a, b))
            t_issue = (time.perf_counter() - t0) * 1e3
            barrier()
            launches = s2.ops.STATS['launches'] - launches0 - (args.steps if flush_l2 else 0) * 0
            allocs = torch.cuda.memory_stats(device).get('num_device_alloc', 0) - allocs0
            ms = sum(a.elapsed_time(b) for a, b in evs)
            clocks = sampler.stop()
            total_bytes = sum(len(s) for lst in obj['strings'] for s in lst)
            sym_per_image = None

        # per-kernel accounting: a second, SERIAL timed pass over the SAME kernels (one batch in flight, CUDA events around every
        # launch on the launching stream) -- with batches overlapping, a kernel's event time would include waiting for SMs.
        n_prof = max(1, min(args.steps, 5))
        if cfg in (1, 2, 5):
            # the timed region enqueues a batch with ONE C call (sc2_fp_encode_batch / sc2_fp_decode_batch); the accounting pass
            # launches the same kernels one by one through the per-layer entry points so that each can be bracketed by events
            model.native_calls = False
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
        prof, work = s2.ops.profile_results(), s2.ops.profile_work()
        s2.ops.profile_kernels(None)
        if cfg in (1, 2, 5):
            model.native_calls = True
        latency_ms = None
        if pipelined or multistream:  # the latency of ONE batch (low-latency coder layout: a warp per stream), for reference
            if pipelined:
                model.entropy_bottleneck.coder_layout = None
            device_step(0)
            l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0.record()
            for i in range(n_prof):
                device_step(i)
            l1.record()
            torch.cuda.synchronize()
            latency_ms = l0.elapsed_time(l1) / n_prof

    t = torch.tensor([ms], dtype=torch.float64, device=device)
    per_rank_ms = [ms]
    counters = parallel.EvalCounters(device)
    counters.add(images=B * args.steps, bytes=total_bytes * args.steps, symbols=(sym_per_image or 0) * B * args.steps)
    if world > 1:
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        per_rank_ms = [float(g.item()) for g in gathered]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    counters.all_reduce()  # the path's only collective: one small counter vector per evaluation
    ms = float(t.item())
    c = counters.as_dict()
    value = c['images'] / (ms / 1e3)

    # ---- end-to-end through the public API with host buffers ("e2e") ---------------------------
    # Each step is the reference-facing call sequence on HOST data: pinned images -> H2D -> encode(x) (returns the contract object
    # with real `bytes` on the host) -> decode(**obj) (bytes -> H2D -> features) -> a per-image result read back.  Host threads
    # each drive their own CUDA stream, so the PCIe copies and the host-side bytes handling of one batch overlap the GPU work of
    # another; every step still runs entirely inside the timed region.
    e2e = None
    if not args.no_e2e:
        import concurrent.futures

        def run_e2e(inputs, n_thr, use_transform_stream):
            streams = [torch.cuda.Stream(device=device) for _ in range(n_thr)]
            # every host thread drives whole steps, so the region holds at least 6 steps per thread (with K = 40 and 16 threads, 2.5
            # steps per thread, the ramp-up and the drain of the thread pool were a quarter of the region and the number was noisy)
            e2e_steps = max(args.steps, 6 * n_thr)

            def e2e_step(i):
                with torch.inference_mode(), torch.cuda.stream(streams[i % n_thr]):
                    x = inputs[i % len(inputs)].to(device, non_blocking=True)
                    obj = encode(x)                             # {'strings': [list[bytes]], 'shape'}: bitstreams land on the host
                    feat = decode(obj)                          # list[bytes] -> device -> features
                    res = feat.mean(dim=(1, 2, 3))
                    torch.cuda.current_stream().synchronize()   # (a blocking .cpu() would hold a driver lock while it waits)
                    res = res.cpu()                             # per-image result read back
                return obj, res

            if use_transform_stream:
                model.use_transform_stream(True, host_wait=True)
                if args.coder_sms > 0:  # as CodecPipeline does: the persistent transform kernels leave SMs to the coder blocks
                    s2._native.check(s2._native.load().sc2_set_persistent_ctas(148 - args.coder_sms), 'sc2_set_persistent_ctas')
            with concurrent.futures.ThreadPoolExecutor(max_workers=n_thr) as pool:
                list(pool.map(e2e_step, range(max(args.warmup, 4 * n_thr))))  # (fills the allocator pools of every stream / thread)
                barrier()
                a0 = torch.cuda.memory_stats(device).get('num_device_alloc', 0)
                ms0 = dict(torch.cuda.memory_stats(device))
                t0 = time.perf_counter()
                results = list(pool.map(e2e_step, range(4 * n_thr, 4 * n_thr + e2e_steps)))
                torch.cuda.synchronize()
                wall = (time.perf_counter() - t0) * 1e3  # host wall clock: host work is part of this contract
                barrier()
            if use_transform_stream:
                model.use_transform_stream(None)
                s2._native.check(s2._native.load().sc2_set_persistent_ctas(0), 'sc2_set_persistent_ctas')
            te = torch.tensor([wall], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            obj, res = results[-1]
            sb = sum(len(s) for lst in obj['strings'] for s in lst)
            n_str = sum(len(lst) for lst in obj['strings'])
            return {'value': B_global * e2e_steps / (float(te.item()) / 1e3), 'unit': UNIT, 'steps': e2e_steps,
                    'h2d_bytes_per_step': inputs[0].numel() * inputs[0].element_size() + sb + 8 * (n_str + 1),
                    'd2h_bytes_per_step': sb + 8 * (n_str + 1) + 4 + B * 4, 'host_threads': n_thr,
                    'cudaMalloc_calls_in_timed_region': torch.cuda.memory_stats(device).get('num_device_alloc', 0) - a0,
                    'ms_per_step': float(te.item()) / e2e_steps, 'input_dtype': str(inputs[0].dtype).replace('torch.', ''),
                    'allocator': {k: torch.cuda.memory_stats(device).get(k, 0) - ms0.get(k, 0) for k in (
                        'num_device_alloc', 'num_device_free', 'segment.small_pool.allocated', 'segment.large_pool.allocated',
                        'segment.large_pool.freed', 'num_alloc_retries', 'reserved_bytes.all.current')}}

        if pipelined:
            cores = os.cpu_count() or 8
            # one host thread per batch in flight: the threads mostly sleep on their streams (3-5 ms of CPU per 60 ms batch latency), so
            # the count follows the pipeline depth the coder chains need (16), not the cores a rank has
            n_thr = args.e2e_threads if args.e2e_threads > 0 else (16 if cores // max(world, 1) >= 2 else 8)
            # uint8 images + device-side ToTensor / Normalize (FPBasedResNetBottleneck.set_input_normalization): 4x less H2D
            model.set_input_normalization(IMAGENET_MEAN, IMAGENET_STD)
            g8 = torch.Generator(device='cpu').manual_seed(11 + rank)
            host_u8 = [torch.randint(0, 256, (B,) + shape, dtype=torch.uint8, generator=g8).pin_memory() for _ in range(2)]
            # (fp32 first: the first e2e run of a process also pays one-time costs -- slots, pinned staging buffers, allocator
            # pools of 16 streams -- that its warm-up does not always cover: one first run in four came out at half speed)
            f32 = run_e2e(host_inputs, n_thr, True)
            e2e = run_e2e(host_u8, n_thr, True)
            e2e['input'] = 'uint8 images, ToTensor + Normalize on the device (inside the first conv kernel)'
            e2e['fp32_input'] = {k: f32[k] for k in ('value', 'h2d_bytes_per_step', 'd2h_bytes_per_step', 'ms_per_step')}
        else:
            # config 1: one image at a time (latency); configs 3 / 4: as many host threads as batches the device-timed run keeps in flight
            e2e = run_e2e(host_inputs, args.e2e_threads if args.e2e_threads > 0 else (1 if cfg in (1, 5) else max(2, args.streams)), False)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    traffic, traffic_src = load_traffic() if cfg == 2 else ({}, None)
    kernels = account_kernels(prof, work, n_prof, serial_ms, (ms / args.steps) if pipelined else None, peaks, traffic, B)
    roofline = pick_roofline(kernels, peaks, traffic_src)
    coder = {k['kernel']: {kk: k.get(kk) for kk in ('avg_launch_ms', 'symbols_per_s_per_stream', 'symbols_per_s_aggregate', 'streams',
                                                    'symbols_per_stream')}
             for k in kernels if k['kernel'] in ('rans_encode', 'rans_decode')}
    path_flops = sum((k.get('algorithmic_flops') or 0.0) * k['launches_per_step'] for k in kernels) / max(B, 1)
    path_bytes = sum((k.get('algorithmic_bytes') or 0.0) * k['launches_per_step'] for k in kernels if k['stream'] == 'transform') / max(B, 1)

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        n_cpu = args.cpu_images or default_cpu_images(cfg)
        r = cpu_reference_run(cfg, n_cpu, steps=2, warmup=1, state_dict=sd)
        cpu_baseline = {'value': r['images_per_s'], 'unit': UNIT, 'cores': r['threads'], 'kind': 'port',
                        'sample': '%d images per step x 2 steps of the same workload; restated reference (oracle/: CompressAI restatement, '
                                  'torch CPU convs on %d threads, per-sample Python-list coder loop + C rANS)' % (n_cpu, r['threads'])}

    line = {'metric': METRICS[cfg], 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
            'dtype': 'f32-grade (g_a: split-f16 operands, 3 tensor-core passes, f32 accumulate) / f16 operands with f32 accumulate (g_s) / u64 (coder)'
                     if cfg in (1, 2, 5) else 'f32-grade (g_a, h_a, h_s: split-f16 operands, 3 tensor-core passes) / f16 operands with f32 accumulate, '
                                              'split activations in the last stage (g_s) / u64 (coder)',
            'data': 'synthetic',
            'config': {'workload': desc + ', batch %d per GPU' % B, 'bench_config': cfg,
                       'images_per_gpu_per_step': B, 'global_images_per_step': B_global, 'symbols_per_image': sym_per_image,
                       'l2_policy': '%d alternating input batches of %.0f MB (%.0f MB > 126 MB L2)' % (n_in, input_bytes / 1e6, n_in * input_bytes / 1e6) if not flush_l2 else
                                    'L2 flushed (256 MB write) between timed iterations, each iteration timed by its own event pair',
                       'parallelism': 'dp%d (batch sharded, one counter all-reduce per evaluation)' % world,
                       'schedule': ('software pipeline, %d batches in flight: transforms on one stream, g_a(i + depth) ahead of g_s(i); '
                                    'coders on per-batch streams' % args.inflight) if pipelined else
                                   ('%d batches in flight, round-robin over %d CUDA streams (device-resident compress_packed / decompress)'
                                    % (args.streams, args.streams)) if multistream else 'one batch at a time (latency)'},
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline, 'coder': coder, 'cpu_baseline': cpu_baseline,
            'serial_ms_per_step': serial_ms, 'one_batch_latency_ms': latency_ms, 'ms_per_image': ms / args.steps / max(B, 1),
            'host_issue_ms_per_step': t_issue / args.steps, 'cudaMalloc_calls_in_timed_region': allocs, 'per_rank_ms': per_rank_ms,
            'kernel_accounting': 'serial pass of %d steps after the timed region, same kernels (CUDA events per launch); shares are of '
                                 'that serial step, as in the ncu launch list' % n_prof,
            'kernels': kernels,
            'bytes_per_image': c['bytes_per_image'], 'bits_per_symbol': c['bits_per_symbol'] if sym_per_image else None,
            'path_tflops': value * path_flops / 1e12, 'path_hbm_gbs_algorithmic': value * path_bytes / 1e9,
            'path_algorithmic_gflop_per_image': path_flops / 1e9, 'path_algorithmic_mb_per_image': path_bytes / 1e6}
    line.update(extra)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
