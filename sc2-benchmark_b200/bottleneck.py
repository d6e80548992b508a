"""Bottleneck layers of the supervised-compression path, mirroring `sc2bench.models.layer`.

Same registry (`LAYER_CLASS_DICT`, `register_layer_class`, `get_layer`), class names, constructor arguments,
child-module names (`encoder`, `decoder`, `entropy_bottleneck`) and `encode` / `decode` / `forward` / `update`
contract as the reference, so YAML `bottleneck_config: {key, kwargs}` entries resolve unchanged:
  - FPBasedResNetBottleneck  <- sc2bench/models/layer.py:444-550
  - BaseBottleneck           <- sc2bench/models/layer.py:401-441
  - EntropyBottleneckLayer   <- sc2bench/models/layer.py:346-398
  - get_layer                <- sc2bench/models/layer.py:820-835
The eval-time branch (`updated and not training`) is the hot path: g_a, quantisation, rANS encode / decode and
g_s all run in libsc2b200.so.  The two training-time branches stay differentiable torch.
"""
import logging
import os
import threading

import torch
from torch import nn

from . import _native, ops
from .entropy_models import GaussianConditional
from .layers import GDN1
from .models import CompressionModel, get_scale_table, run_analysis, run_transform, update_registered_buffers



_log = logging.getLogger('sc2bench_b200')
_warned = set()
_PLAN_LOCK = threading.RLock()  # re-entrant: building a native codec (under the lock) prepares the per-layer plans (under the lock)


def warn_fallback(what, why):
    """One log line per (layer, reason) whenever a transform leaves the tensor-core kernels for the exact-fp32 CUDA-core ones
    (conv2d_f32_kernel: 10-20 TFLOP/s), so that a slow model is visible as such."""
    key = (what, why)
    if key not in _warned:
        _warned.add(key)
        _log.warning('%s runs on the fp32 CUDA-core kernels (conv2d_f32_kernel), not on the tensor cores: %s', what, why)


def _param_key(seq):
    return tuple((q.data_ptr(), q._version, q.device) for q in seq.parameters())


class TensorCoreTransform:
    """Execution plan that runs a stride-1 `Conv2d / GDN1` transform (the bottleneck's synthesis transform g_s) on the
    tcgen05 kernels: activations NHWC fp16 between layers, fp32 accumulation in TMEM, GDN1 fused with its 1x1 gamma
    GEMM, last layer written as fp32.  Weights are repacked once and re-packed when the parameters change."""

    def __init__(self, seq):
        self.seq = seq
        self._plan = None  # (key, steps, c_in_pad): built into locals and published with ONE assignment (host threads share it)
        self.split_signs = True  # conv -> IGDN1 pairs exchange (|x|, sign words); False: round-1 route (x, |.| pass in the kernel)

    @staticmethod
    def why_not(seq):
        """None when the plan covers `seq`, else the reason (logged by the caller)."""
        mods = list(seq)
        if not mods or not isinstance(mods[-1], nn.Conv2d):
            return 'the transform does not end in a Conv2d'
        for m in mods:
            if isinstance(m, nn.Conv2d):
                k = m.kernel_size
                if m.bias is not None or m.groups != 1 or tuple(m.dilation) != (1, 1):
                    return 'bias / groups / dilation'
                if tuple(m.stride) != (1, 1):
                    return 'stride %s' % (tuple(m.stride),)
                if k[0] != k[1] or isinstance(m.padding, str) or m.padding[0] != m.padding[1]:
                    return 'non-square kernel or padding'
                if m.out_channels % 64:
                    return 'c_out %d is not a multiple of 64' % m.out_channels
            elif type(m) is GDN1:
                if m.beta.numel() % 64 or m.beta.numel() > 512:  # (the kernel keeps beta in shared memory: <= 512 channels)
                    return 'GDN1 over %d channels' % m.beta.numel()
            else:
                return 'layer type %s' % type(m).__name__
        return None

    @classmethod
    def supports(cls, seq):
        return cls.why_not(seq) is None

    def _prepare(self):
        key = _param_key(self.seq)
        plan = self._plan
        if plan is not None and plan[0] == key:
            return plan
        with _PLAN_LOCK:
            plan = self._plan
            if plan is not None and plan[0] == key:
                return plan
            steps, mods = [], list(self.seq)
            for i, m in enumerate(mods):
                if isinstance(m, nn.Conv2d):
                    c_in_pad = (m.in_channels + 63) // 64 * 64
                    mode = _native.TC_STORE_F32 if i == len(mods) - 1 else _native.TC_STORE_F16
                    # a conv in front of an IGDN1 stores |x| + packed signs: the IGDN1's gamma GEMM then reads |x| straight from
                    # its TMA tiles (conv_tc.cu, modes 4 / 5)
                    nxt = mods[i + 1] if i + 1 < len(mods) else None
                    if self.split_signs and type(nxt) is GDN1 and nxt.inverse and m.out_channels % 32 == 0:
                        mode = _native.TC_STORE_ABS_F16
                    steps.append(('conv', ops.pack_conv_weight_f16(m.weight, c_in_pad), m.kernel_size[0], m.padding[0], mode, None,
                                  m.in_channels))
                else:
                    gamma, beta = m.effective_params()
                    C = beta.numel()
                    mode = _native.TC_IGDN1_F16 if m.inverse else _native.TC_GDN1_F16
                    if steps and steps[-1][4] == _native.TC_STORE_ABS_F16:
                        mode = _native.TC_IGDN1_ABS_F16
                    steps.append(('gdn', gamma.detach().reshape(1, C, C).half().contiguous(), 1, 0, mode, beta.detach().float().contiguous(), C))
            plan = (key, steps, (mods[0].in_channels + 63) // 64 * 64)
            self._plan = plan
        return plan

    @torch.no_grad()
    def __call__(self, x_nchw):
        """fp32 NCHW in -> fp32 output, logically NCHW (physically channels-last: what cuDNN prefers for the tail)."""
        _, steps, c_in_pad = self._prepare()
        x = ops.nchw_to_nhwc_f16(x_nchw, c_in_pad)
        signs = None
        for kind, w, k, pad, mode, beta, c_in in steps:
            x = ops.tc_conv(x, w, k, k, pad, mode=mode, beta=beta, gdn_x=x if kind == 'gdn' else None, c_in=c_in, signs=signs)
            signs = None
            if mode == _native.TC_STORE_ABS_F16:
                x, signs = x
        return x.permute(0, 3, 1, 2)


_NO_QUANT = object()  # TensorCoreAnalysis(x, _NO_QUANT): return the latent y (fp32 NCHW) instead of symbols


class TensorCoreAnalysis:
    """Execution plan for the bottleneck's analysis transform g_a (Conv s2 - GDN1 - Conv s2 - GDN1 - Conv s1) on the
    fp32-grade "split fp16" tcgen05 kernels, ending in the fused quantise-to-symbols epilogue.  Three launches:

        image (fp32 NCHW, or uint8 + normalisation table)
          --[conv 5x5 s2 + GDN1, conv_ga_first.cu]--> y1 (parity planes)
          --[conv 5x5 s2 + GDN1, conv_ga_halo.cu]---> y2
          --[conv 2x2 + round(y - median), conv_tc_split.cu]--> int32 symbols (NCHW = coder order)

    Every intermediate is a (hi, lo) pair of NHWC fp16 planes.  Weights / gammas are split, stacked and packed once.
    Shapes the fused kernels do not cover take the round-1 route (separate conv and GDN1 launches on the split kernels)."""

    def __init__(self, seq):
        self.seq = seq
        self._plan = None
        self.fuse_first_layer = True
        self.fused = True        # conv + GDN1 in one kernel where the shapes allow (False: round-1 route, for A/B measurements)
        self._unfused = set()    # (stage, shape) combinations the fused kernels refused
        self._env_unfused = set(os.environ.get('SC2_GA_UNFUSED', '').split(','))  # experiments: 'first', 'mid' -> two-kernel route

    @staticmethod
    def why_not(seq, x_shape):
        mods = list(seq)
        if len(mods) != 5 or not all(isinstance(mods[i], nn.Conv2d) for i in (0, 2, 4)) or not all(type(mods[i]) is GDN1 for i in (1, 3)):
            return 'not Conv - GDN1 - Conv - GDN1 - Conv'
        c1, c2, c3 = mods[0], mods[2], mods[4]
        for c in (c1, c2, c3):
            if (c.bias is not None or c.groups != 1 or tuple(c.dilation) != (1, 1) or c.kernel_size[0] != c.kernel_size[1]
                    or isinstance(c.padding, str) or c.padding[0] != c.padding[1] or c.stride[0] != c.stride[1]):
                return 'bias / groups / dilation / non-square kernel'
        if c1.stride[0] != 2 or c2.stride[0] != 2 or c3.stride[0] != 1 or mods[1].inverse or mods[3].inverse:
            return 'strides are not (2, 2, 1) or a GDN1 is inverse'
        if c1.in_channels * c1.kernel_size[0] ** 2 > 128 or c2.kernel_size[0] ** 2 > 25 or c3.kernel_size[0] ** 2 > 25:
            return 'kernel too large'
        if c1.out_channels % 16 or c2.out_channels % 16 or max(c1.out_channels, c2.out_channels, c3.out_channels) > 128:
            return 'channel counts %d / %d / %d' % (c1.out_channels, c2.out_channels, c3.out_channels)
        H, W = x_shape[-2:]
        k, p = c1.kernel_size[0], c1.padding[0]
        h1, w1 = (H + 2 * p - k) // 2 + 1, (W + 2 * p - k) // 2 + 1
        if not (h1 >= 2 and w1 >= 2 and h1 % 2 == 0 and w1 % 2 == 0):
            return 'first-layer output %d x %d is not even' % (h1, w1)
        return None

    @classmethod
    def supports(cls, seq, x_shape):
        return cls.why_not(seq, x_shape) is None

    def _prepare(self):
        key = _param_key(self.seq)
        plan = self._plan
        if plan is not None and plan['key'] == key:
            return plan
        with _PLAN_LOCK:
            plan = self._plan
            if plan is not None and plan['key'] == key:
                return plan
            c1, g1, c2, g2, c3 = list(self.seq)
            plan = {'key': key}
            plan['k1_pad'] = (c1.in_channels * c1.kernel_size[0] ** 2 + 15) // 16 * 16
            plan['w1'] = ops.pack_conv_weight_split(c1.weight, c_in_pad=plan['k1_pad'], as_patches=True)
            plan['w2'] = ops.pack_conv_weight_split(c2.weight)
            plan['w3'] = ops.pack_conv_weight_split(c3.weight)
            plan['gdn'] = []
            for g in (g1, g2):
                gamma, beta = g.effective_params()
                C = beta.numel()
                plan['gdn'].append((ops.pack_conv_weight_split(gamma.detach().reshape(C, C, 1, 1)), beta.detach().float().contiguous()))
            # stacked (hi; lo) packs of the fused kernels
            lib = _native.load()
            n1, n2 = lib.sc2_ga_halo_n(c1.out_channels), lib.sc2_ga_halo_n(c2.out_channels)
            first_ok = (c1.in_channels, c1.kernel_size[0], c1.stride[0], c1.padding[0]) == (3, 5, 2, 2) and n1 > 0
            plan['first'] = None
            if first_ok:
                gamma, _ = g1.effective_params()
                C = c1.out_channels
                plan['first'] = (ops.pack_first_layer_stacked(c1.weight),
                                 ops.pack_conv_weight_stacked(gamma.detach().reshape(C, C, 1, 1), n=n1, c_in_pad=n1)[0].contiguous())
            plan['mid'] = None
            if n2 > 0 and c2.stride[0] == 2 and c2.in_channels % 16 == 0:
                gamma, _ = g2.effective_params()
                C = c2.out_channels
                plan['mid'] = (ops.pack_conv_weight_stacked(c2.weight),
                               ops.pack_conv_weight_stacked(gamma.detach().reshape(C, C, 1, 1), n=n2, c_in_pad=n2)[0].contiguous())
            self._plan = plan
        return plan

    def _try_fused(self, stage, shape_key, fn):
        """Runs a fused kernel; a shape it refuses (SC2_ERR_UNSUPPORTED) is remembered and takes the two-kernel route from then on."""
        if not self.fused or stage in self._env_unfused or (stage, shape_key) in self._unfused:
            return None
        try:
            return fn()
        except _native.NativeError as e:
            if 'unsupported' not in str(e):
                raise
            self._unfused.add((stage, shape_key))
            _log.info('g_a %s stage: fused conv + GDN1 kernel does not cover %s; using separate conv and GDN1 launches', stage, shape_key)
            return None

    @torch.no_grad()
    def __call__(self, x, medians, lut=None):
        plan = self._prepare()
        c1, _, c2, _, c3 = list(self.seq)
        T = _native
        out = None
        if plan['first'] is not None:
            ws, gs = plan['first']
            out = self._try_fused('first', (tuple(x.shape[1:]), x.dtype), lambda: ops.ga_first_conv_gdn(x, ws, gs, plan['gdn'][0][1], c1.out_channels, lut=lut))
        if out is None:
            if x.dtype == torch.uint8:
                x = ops.normalize_u8(x, lut)
            if c1.out_channels <= 96 and self.fuse_first_layer:
                # im2col fused into the kernel: no patch tensor in HBM
                h, l = ops.tc_first_layer(x, plan['w1'][0], plan['w1'][1], c1.out_channels, c1.kernel_size[0], c1.kernel_size[0], c1.padding[0])
            else:
                ph, pl = ops.patchify_split(x, c1.kernel_size[0], c1.kernel_size[0], 2, c1.padding[0], plan['k1_pad'])
                h, l = ops.tc_split_conv(ph, pl, plan['w1'][0], plan['w1'][1], c1.out_channels, 1, 1, 1, 0, T.TCS_STORE)
            (gh, gl), beta = plan['gdn'][0]
            out = ops.tc_split_conv(h, l, gh, gl, c1.out_channels, 1, 1, 1, 0, T.TCS_GDN1, beta=beta, gdn=True)
        h, l = out
        out = None
        if plan['mid'] is not None:
            ws, gs = plan['mid']
            out = self._try_fused('mid', tuple(h.shape[1:]), lambda: ops.ga_halo_conv_gdn(h, l, ws, gs, plan['gdn'][1][1], c2.out_channels,
                                                                                         c2.kernel_size[0], c2.kernel_size[0], c2.padding[0]))
        if out is None:
            h, l = ops.tc_split_conv(h, l, plan['w2'][0], plan['w2'][1], c2.out_channels, c2.kernel_size[0], c2.kernel_size[0], 2,
                                     c2.padding[0], T.TCS_STORE)
            (gh, gl), beta = plan['gdn'][1]
            out = ops.tc_split_conv(h, l, gh, gl, c2.out_channels, 1, 1, 1, 0, T.TCS_GDN1, beta=beta, gdn=True)
        h, l = out
        if medians is _NO_QUANT:
            # the latent itself (scale-hyperprior bottlenecks need y for h_a and for the Gaussian-conditional coder): split planes
            # -> fp32 NCHW (value = hi + lo / 2048)
            oh, ol = ops.tc_split_conv(h, l, plan['w3'][0], plan['w3'][1], c3.out_channels, c3.kernel_size[0], c3.kernel_size[0], 1,
                                       c3.padding[0], T.TCS_STORE)
            y = torch.add(oh[..., :c3.out_channels].float(), ol[..., :c3.out_channels].float(), alpha=1.0 / ops.LO_SCALE)
            return y.permute(0, 3, 1, 2).contiguous()
        return ops.tc_split_conv(h, l, plan['w3'][0], plan['w3'][1], c3.out_channels, c3.kernel_size[0], c3.kernel_size[0], 1,
                                 c3.padding[0], T.TCS_QUANT, medians=medians)


def _run_synthesis(layer, seq, latent_hat):
    """g_s on the tensor-core plan when `layer.decoder_precision` and the shapes allow, else on the fp32 kernels (logged)."""
    why = 'decoder_precision = %r' % layer.decoder_precision if layer.decoder_precision != 'fp16-tc' else TensorCoreTransform.why_not(seq)
    if why is None:
        if layer._tc_decoder is None:
            with _PLAN_LOCK:
                if layer._tc_decoder is None:
                    layer._tc_decoder = TensorCoreTransform(seq)
        return layer._tc_decoder(latent_hat)
    warn_fallback('%s synthesis transform (g_s)' % type(layer).__name__, why)
    return run_transform(seq, latent_hat)


LAYER_CLASS_DICT = dict()
LAYER_FUNC_DICT = dict()


def register_layer_class(cls):
    LAYER_CLASS_DICT[cls.__name__] = cls
    return cls


def register_layer_func(func):
    LAYER_FUNC_DICT[func.__name__] = func
    return func


def get_layer(cls_or_func_name, **kwargs):
    """Builds a registered layer; `None` for an unknown key, like the reference."""
    if cls_or_func_name in LAYER_CLASS_DICT:
        return LAYER_CLASS_DICT[cls_or_func_name](**kwargs)
    if cls_or_func_name in LAYER_FUNC_DICT:
        return LAYER_FUNC_DICT[cls_or_func_name](**kwargs)
    return None


class EntropyBottleneckLayer(CompressionModel):
    """A bare EntropyBottleneck as a CompressionModel (dropped after a backbone stage by EntropicClassifier)."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.updated = False

    def forward(self, x):
        return self.entropy_bottleneck(x)

    def compress(self, x):
        return {'strings': [self.entropy_bottleneck.compress(x)], 'shape': x.size()[-2:]}

    def decompress(self, strings, shape):
        assert isinstance(strings, list) and len(strings) == 1
        return self.entropy_bottleneck.decompress(strings[0], shape)

    def update(self, force=False):
        self.updated = True
        return super().update(force=force)


class BaseBottleneck(CompressionModel):
    """Entropy-bottleneck based encoder / decoder pair; subclasses provide encode() / decode() / forward()."""

    def __init__(self, entropy_bottleneck_channels):
        super().__init__(entropy_bottleneck_channels=entropy_bottleneck_channels)
        self.updated = False

    def encode(self, *args, **kwargs):
        raise NotImplementedError()

    def decode(self, *args, **kwargs):
        raise NotImplementedError()

    def forward(self, *args):
        raise NotImplementedError()

    def update(self, force=False):
        self.updated = True
        return super().update(force=force)


@register_layer_class
class FPBasedResNetBottleneck(BaseBottleneck):
    """Factorized-prior bottleneck for ResNet-style students (Entropic Student, Matsubara et al. WACV 2022).

    encoder: Conv(5x5, s2) - GDN1 - Conv(5x5, s2) - GDN1 - Conv(2x2)            3 -> 4b -> 2b -> b channels
    decoder: Conv(2x2, p1) - IGDN1 - Conv(2x2) - IGDN1 - Conv(2x2, p1)          b -> 2t -> t -> t channels
    with b = num_bottleneck_channels, t = num_target_channels; no conv has a bias.
    """

    def __init__(self, num_input_channels=3, num_bottleneck_channels=24, num_target_channels=256,
                 encoder_channel_sizes=None, decoder_channel_sizes=None):
        if encoder_channel_sizes is None:
            b = num_bottleneck_channels
            encoder_channel_sizes = [num_input_channels, b * 4, b * 2, b]
        if decoder_channel_sizes is None:
            t = num_target_channels
            decoder_channel_sizes = [encoder_channel_sizes[-1], t * 2, t, t]
        super().__init__(entropy_bottleneck_channels=num_bottleneck_channels)
        e, d = encoder_channel_sizes, decoder_channel_sizes
        self.encoder = nn.Sequential(
            nn.Conv2d(e[0], e[1], kernel_size=5, stride=2, padding=2, bias=False), GDN1(e[1]),
            nn.Conv2d(e[1], e[2], kernel_size=5, stride=2, padding=2, bias=False), GDN1(e[2]),
            nn.Conv2d(e[2], e[3], kernel_size=2, stride=1, padding=0, bias=False))
        self.decoder = nn.Sequential(
            nn.Conv2d(d[0], d[1], kernel_size=2, stride=1, padding=1, bias=False), GDN1(d[1], inverse=True),
            nn.Conv2d(d[1], d[2], kernel_size=2, stride=1, padding=0, bias=False), GDN1(d[2], inverse=True),
            nn.Conv2d(d[2], d[3], kernel_size=2, stride=1, padding=1, bias=False))
        # 'fp16-tc': tcgen05 tensor-core kernels (fp16 operands, fp32 accumulate; the 1e-3 feature tolerance);
        # 'fp32': exact-fp32 CUDA-core kernels.  Shapes the tensor-core kernels do not cover use 'fp32'.
        self.decoder_precision = 'fp16-tc'
        self._tc_decoder = None
        # 'split-tc': fp32-grade split-fp16 tensor-core kernels (three MMA passes); 'fp32': exact-fp32 CUDA-core kernels.
        self.encoder_precision = 'split-tc'
        self._tc_encoder = None
        self._transform_stream = None

    # ---- hot path ---------------------------------------------------------------------------------
    def use_transform_stream(self, stream=True, host_wait=False):
        """Throughput mode for callers that keep several batches in flight (one CUDA stream, or one host thread, per batch).

        The transforms are persistent tensor-core kernels that fill the GPU; the coder of a batch is a handful of warps that
        run for milliseconds.  Issued on per-batch streams, the batches drift into lock-step -- all in their transforms, then
        all in their coders with the tensor cores idle (measured with the per-CTA trace, scripts/diag_trace.py).  With a
        transform stream every g_a / g_s runs on that ONE stream in call order (event-ordered against the caller's current
        stream), and only the coders stay on the callers' streams: calling encode for batch i + d before decode for batch i
        is then a software pipeline of depth d.  `stream`: a torch.cuda.Stream, True (create one) or None / False (off).
        host_wait: for one-host-thread-per-batch callers -- the calling thread waits (GIL released) until its own stream has
        produced the transform's input before it queues the transform, so that a batch whose copy or coder is still running
        does not hold up the transforms of the other threads' batches."""
        self._transform_host_wait = bool(host_wait)
        if stream is True:
            # High priority: when SMs free up, the transforms' CTAs go first and coder blocks (normal priority, one SM each,
            # milliseconds long) take what is left in the kernel tails.  At equal priority about one run in five settled at
            # 7.7-10 ms per step instead of 5.9 (with the coder streams at HIGH priority, seven in ten); 10 of 10 runs with this.
            stream = torch.cuda.Stream(device=self.entropy_bottleneck._quantized_cdf.device, priority=-1)
        self._transform_stream = stream or None
        # batches in flight: the coder layout that leaves the SMs to the transforms (sc2_rans_encode_batch, `layout`): a lane per stream
        # for batches of at least one warp of streams, else a warp per stream (_native.rans_layout)
        self.entropy_bottleneck.coder_layout = 'throughput' if self._transform_stream is not None else None
        return self._transform_stream

    def _on_transform_stream(self, fn, *tensors):
        """Runs fn() on the transform stream, ordered after the current stream's work on `tensors` and before whatever the
        current stream does next with the result."""
        ts = self._transform_stream
        if ts is None:
            return fn()
        cur = torch.cuda.current_stream()
        if ts == cur:
            return fn()
        if getattr(self, '_transform_host_wait', False):
            cur.synchronize()
        else:
            ts.wait_stream(cur)
        for t in tensors:
            t.record_stream(ts)
        with torch.cuda.stream(ts):
            out = fn()
        out.record_stream(cur)
        cur.wait_stream(ts)
        return out

    @torch.no_grad()
    def encode_packed(self, x, _slot_buffers=False):
        """g_a + quantise + rANS, bitstreams left on the device: (PackedStreams, latent (H, W))."""
        nat = self._native_state(x)
        if nat is not None:
            codec, slot, ws = nat
            dev = x.device
            cur = torch.cuda.current_stream(dev)
            ts = self._transform_stream or cur
            if _slot_buffers:
                # (encode(): the bytes are copied to the host before the call returns, so the calling thread's ring buffers do --
                # no worst-case-sized device allocation per call)
                packed = offsets = status = None
            else:
                packed = torch.empty(codec.batch * codec.slot_bytes, dtype=torch.uint8, device=dev)
                offsets = torch.empty(codec.batch + 1, dtype=torch.int64, device=dev)
                status = torch.empty(1, dtype=torch.int32, device=dev)  # (zeroed by the call, on the coder stream)
            ev_in = None
            if ts != cur:
                if getattr(self, '_transform_host_wait', False):
                    cur.synchronize()
                else:
                    slot.ev_in.record(cur)
                    ev_in = slot.ev_in
                x.record_stream(ts)
            streams = codec.encode(x.contiguous(), slot, ts, cur, ev_in=ev_in, packed=packed, offsets=offsets, status=status,
                                   ws=ws[0] if ws else None)
            return streams, torch.Size(codec.latent_hw)
        eb = self.entropy_bottleneck
        symbols = self._on_transform_stream(lambda: self.analyze_to_symbols(x), x)
        return eb.compress_symbols(symbols, spatial=symbols[0, 0].numel()), symbols.size()[-2:]

    def _native_state(self, x, batch_shape=None):
        """(codec, slot, private workspaces or None) of the calling thread and its current stream for batches shaped like x, or
        None when the one-call-per-batch route (native_codec.FpNativeCodec) does not apply."""
        if not getattr(self, 'native_calls', True) or not isinstance(x, torch.Tensor) or not x.is_cuda or x.dim() != 4:
            return None
        from .native_codec import FpNativeCodec
        key = (x.device, tuple(x.shape) if batch_shape is None else tuple(batch_shape), getattr(self, '_input_norm', None) is not None)
        codecs = self.__dict__.setdefault('_native_codecs', {})
        codec = codecs.get(key)
        if codec is None or (codec is not False and codec.stale()):
            with _PLAN_LOCK:
                codec = codecs.get(key)
                if codec is None or (codec is not False and codec.stale()):
                    try:
                        with torch.inference_mode(False):
                            codec = FpNativeCodec(self, key[1][0], key[1][2], key[1][3], x.device,
                                                  coder_layout=self.entropy_bottleneck.coder_layout if getattr(self.entropy_bottleneck, 'coder_layout', None) else 'auto')
                    except (ValueError, _native.NativeError) as e:
                        _log.info('%s: per-kernel route for batches of shape %s (%s)', type(self).__name__, key[1], e)
                        codec = False
                    codecs[key] = codec
        if codec is False:
            return None
        wanted = _native.rans_layout(getattr(self.entropy_bottleneck, 'coder_layout', None), codec.batch)
        codec.coder_layout = wanted
        # one slot per CUDA stream (all work on a slot's buffers is ordered by its stream, whichever host thread issues it)
        slots = codec.__dict__.setdefault('_stream_slots', {})
        shared_ws = self._transform_stream is not None  # transforms of all callers run on ONE stream: one workspace pair serves all
        skey = (torch.cuda.current_stream(x.device).cuda_stream, shared_ws)
        st = slots.get(skey)
        if st is None:
            with _PLAN_LOCK:
                st = slots.get(skey)
                if st is None:
                    st = (codec.new_slot(), None if shared_ws else codec.make_workspaces())
                    slots[skey] = st
        return codec, st[0], st[1]

    def set_input_normalization(self, mean, std):
        """Device-side ToTensor + Normalize: after this call `encode` also accepts uint8 NCHW images and
        applies (v / 255 - mean[c]) / std[c] on the fly inside the first layer's im2col (a 3 x 256 table built with the data
        loader's own torch ops, so the values are bit-identical); host-to-device traffic drops 4x.  fp32 input keeps working."""
        self._input_norm = (tuple(float(m) for m in mean), tuple(float(v) for v in std))
        self._input_luts = {}

    def _input_lut(self, device):
        norm = getattr(self, '_input_norm', None)
        if norm is None:
            raise ValueError('uint8 images need set_input_normalization(mean, std) first')
        lut = self._input_luts.get(device)
        if lut is None:
            with torch.inference_mode(False):
                lut = ops.normalize_lut(norm[0], norm[1], device)
            self._input_luts[device] = lut
        return lut

    @torch.no_grad()
    def analyze_to_symbols(self, x):
        """g_a + round(y - median) on the device: image batch -> int32 symbols [B, C, H, W] (coder order)."""
        ops.require_cuda(x, 'FPBasedResNetBottleneck.encode')
        medians = self.entropy_bottleneck._get_medians().detach().reshape(-1)
        lut = self._input_lut(x.device) if x.dtype == torch.uint8 else None
        why = 'encoder_precision = %r' % self.encoder_precision if self.encoder_precision != 'split-tc' else \
            TensorCoreAnalysis.why_not(self.encoder, x.shape)
        if why is None:
            if self._tc_encoder is None:
                with _PLAN_LOCK:
                    if self._tc_encoder is None:
                        self._tc_encoder = TensorCoreAnalysis(self.encoder)
            return self._tc_encoder(x, medians, lut=lut)
        warn_fallback('%s.encoder (g_a)' % type(self).__name__, why)
        if lut is not None:
            x = ops.normalize_u8(x, lut)
        return run_transform(self.encoder, x, final_epilogue=_native.EPI_QUANTIZE, final_aux=medians)

    @torch.no_grad()
    def synthesize(self, latent_hat):
        """g_s on the device: dequantised latent (fp32 NCHW) -> decoder features."""
        return _run_synthesis(self, self.decoder, latent_hat)

    @torch.no_grad()
    def decode_packed(self, streams, shape, check_status=False):
        eb = self.entropy_bottleneck
        nat = None
        if getattr(self, 'native_calls', True) and streams.packed.is_cuda:
            for (dev, bshape, _), codec in self.__dict__.get('_native_codecs', {}).items():
                if codec is not False and dev == streams.packed.device and bshape[0] == streams.batch and tuple(codec.latent_hw) == tuple(shape) \
                        and not codec.stale():
                    nat = self._native_state(streams.packed.new_empty((0, 0, 0, 0)), batch_shape=bshape)
                    break
        if nat is None:
            latent_hat = eb.decompress_packed(streams, tuple(shape), check_status=check_status)
            return self._on_transform_stream(lambda: self.synthesize(latent_hat), latent_hat)
        codec, slot, ws = nat
        dev = streams.packed.device
        cur = torch.cuda.current_stream(dev)
        ts = self._transform_stream or cur
        # (the result is allocated from the TRANSFORM stream's pool, like the per-kernel route does: one pool recycles the blocks of
        # every caller; per-caller pools of 0.8 GB blocks kept the allocator calling cudaMalloc)
        with torch.cuda.stream(ts):
            out = torch.empty((codec.batch, codec.out_hw[0], codec.out_hw[1], codec.d3), dtype=torch.float32, device=dev)
        if ts != cur and getattr(self, '_transform_host_wait', False):
            # one host thread per batch: decode on the caller's stream, WAIT for it on the host, then queue g_s -- a transform stream
            # that waited for this batch's decoder on the device would hold up the transforms of every other thread's batch
            latent_hat = eb.decompress_packed(streams, tuple(shape), check_status=check_status)
            if not check_status:
                cur.synchronize()
            out.record_stream(cur)
            latent_hat.record_stream(ts)
            feats = codec.decode(slot, ts, cur, out=out, ws=ws[1] if ws else None, latent_hat=latent_hat)
        else:
            status = torch.zeros(1, dtype=torch.int32, device=dev) if check_status else eb._fault_word(dev)
            if ts != cur:
                out.record_stream(cur)
            feats = codec.decode(slot, ts, cur, packed=streams.packed, offsets=streams.offsets, status=status, out=out,
                                 ws=ws[1] if ws else None)
            if check_status:
                slot.ev_mid2.synchronize() if ts != cur else cur.synchronize()
                ops.raise_on_decode_fault(int(status.item()))
        if ts != cur:
            cur.wait_event(slot.ev_out)
        return feats

    def encode(self, x, **kwargs):
        """-> {'strings': [list of B bytes objects], 'shape': latent (H, W)}  (reference contract, layer.py:496-507)"""
        streams, shape = self.encode_packed(x, _slot_buffers=True)
        return {'strings': [streams.tolist()], 'shape': shape}

    def decode(self, strings, shape):
        """strings[0]: list of B bytes objects (or a device-resident PackedStreams) -> decoder features."""
        first = strings[0]
        from_host = not isinstance(first, ops.PackedStreams)
        if from_host:
            first = ops.PackedStreams.from_list(first, self.entropy_bottleneck._quantized_cdf.device)
        return self.decode_packed(first, shape, check_status=from_host)  # bytes from outside are validated eagerly

    # ---- training-time branches (differentiable torch, off the hot path) ----------------------------
    def _get_means(self, x):
        medians = self.entropy_bottleneck._get_medians().detach()
        spatial_dims = x.dim() - 2
        medians = self.entropy_bottleneck._extend_ndims(medians, spatial_dims)
        return medians.expand(x.size(0), *([-1] * (spatial_dims + 1)))

    def _forward2train(self, x):
        y_hat, _ = self.entropy_bottleneck(self.encoder(x))
        return self.decoder(y_hat)

    def forward(self, x):
        if not self.updated:
            return self._forward2train(x)
        if not self.training:
            return self.decode(**self.encode(x))
        # fine-tuning after update(): hard rounding around the medians, no gradient through it
        latent = self.encoder(x)
        eb = self.entropy_bottleneck
        rounded = eb.dequantize(eb.quantize(latent, 'dequantize', self._get_means(latent)))
        return self.decoder(rounded.detach())


@register_layer_class
class SHPBasedResNetBottleneck(BaseBottleneck):
    """Scale-hyperprior bottleneck for ResNet-style students (mirrors sc2bench/models/layer.py:553-720).

    y = g_a(x); z = h_a(|y|) is coded with the factorized EntropyBottleneck; both sides decode z, derive per-element scales
    with h_s and code y with the Gaussian conditional (explicit CDF index per element).  strings = [y_strings, z_strings],
    shape = z's spatial size.  The transforms run on the exact-fp32 kernels; the coder on the generic (indexed) rANS kernels."""

    def __init__(self, num_input_channels=3, num_latent_channels=16, num_bottleneck_channels=24, num_target_channels=256,
                 h_a=None, h_s=None, g_a_channel_sizes=None, g_s_channel_sizes=None):
        if g_a_channel_sizes is None:
            b = num_bottleneck_channels
            g_a_channel_sizes = [num_input_channels, b * 4, b * 2, b]
        else:
            num_bottleneck_channels = g_a_channel_sizes[3]
        if g_s_channel_sizes is None:
            t = num_target_channels
            g_s_channel_sizes = [g_a_channel_sizes[-1], t * 2, t, t]
        super().__init__(entropy_bottleneck_channels=num_latent_channels)
        a, g, L, b = g_a_channel_sizes, g_s_channel_sizes, num_latent_channels, num_bottleneck_channels
        self.g_a = nn.Sequential(
            nn.Conv2d(a[0], a[1], kernel_size=5, stride=2, padding=2, bias=False), GDN1(a[1]),
            nn.Conv2d(a[1], a[2], kernel_size=5, stride=2, padding=2, bias=False), GDN1(a[2]),
            nn.Conv2d(a[2], a[3], kernel_size=2, stride=1, padding=0, bias=False))
        self.g_s = nn.Sequential(
            nn.Conv2d(g[0], g[1], kernel_size=2, stride=1, padding=1, bias=False), GDN1(g[1], inverse=True),
            nn.Conv2d(g[1], g[2], kernel_size=2, stride=1, padding=0, bias=False), GDN1(g[2], inverse=True),
            nn.Conv2d(g[2], g[3], kernel_size=2, stride=1, padding=1, bias=False))
        self.h_a = h_a if h_a is not None else nn.Sequential(
            nn.Conv2d(b, L, kernel_size=5, stride=2, padding=1, bias=False), nn.ReLU(inplace=True),
            nn.Conv2d(L, L, kernel_size=5, stride=2, padding=2, bias=False))
        self.h_s = h_s if h_s is not None else nn.Sequential(
            nn.ConvTranspose2d(L, L, kernel_size=5, stride=2, padding=1, bias=False), nn.LeakyReLU(inplace=True),
            nn.ConvTranspose2d(L, L, kernel_size=5, stride=2, padding=1, bias=False), nn.LeakyReLU(inplace=True),
            nn.Conv2d(L, b, kernel_size=5, stride=1, padding=0, bias=False))
        self.gaussian_conditional = GaussianConditional(None)
        self.num_latent_channels = L
        self.num_bottleneck_channels = b
        self.decoder_precision = 'fp16-tc'
        self._tc_decoder = None

    @torch.no_grad()
    def _analysis(self, x):
        """y = g_a(x) as fp32 NCHW: the fused tensor-core kernels of the factorized-prior bottleneck (same Conv - GDN1 - Conv - GDN1
        - Conv topology), else the fp32 CUDA-core kernels (logged)."""
        ops.require_cuda(x, type(self).__name__ + '.encode')
        why = 'encoder_precision = %r' % self.encoder_precision if getattr(self, 'encoder_precision', 'split-tc') != 'split-tc' else \
            TensorCoreAnalysis.why_not(self.g_a, x.shape)
        if why is None:
            if self.__dict__.get('_tc_encoder') is None:
                with _PLAN_LOCK:
                    if self.__dict__.get('_tc_encoder') is None:
                        self.__dict__['_tc_encoder'] = TensorCoreAnalysis(self.g_a)
            return self.__dict__['_tc_encoder'](x, _NO_QUANT)
        warn_fallback('%s.g_a' % type(self).__name__, why)
        return run_transform(self.g_a, x)

    @torch.no_grad()
    def _scales_to_indexes(self, z_hat):
        return self.gaussian_conditional.build_indexes(run_analysis(self, 'h_s', self.h_s, z_hat))

    @torch.no_grad()
    def encode(self, x, **kwargs):
        eb, gc = self.entropy_bottleneck, self.gaussian_conditional
        y = self._analysis(x)
        z_symbols = run_analysis(self, 'h_a', self.h_a, y, medians=eb._get_medians().detach().reshape(-1), out='symbols',
                                 in_abs=True)  # h_a(|y|)
        z_shape = z_symbols.size()[-2:]
        z_streams = eb.compress_symbols(z_symbols, spatial=z_symbols[0, 0].numel())
        z_hat = eb.decompress_packed(z_streams, tuple(z_shape))  # the encoder decodes z itself, like the decoder will
        indexes = self._scales_to_indexes(z_hat)
        y_symbols = ops.quantize_symbols(y.reshape(y.size(0), 1, -1))
        y_streams = ops.rans_encode(y_symbols, gc.coder_tables(), indexes=indexes)
        return {'strings': [y_streams.tolist(), z_streams.tolist()], 'shape': z_shape}

    @torch.no_grad()
    def decode(self, strings, shape):
        assert isinstance(strings, list) and len(strings) == 2
        eb, gc = self.entropy_bottleneck, self.gaussian_conditional
        device = eb._quantized_cdf.device
        z_hat = eb.decompress_packed(ops.PackedStreams.from_list(strings[1], device), tuple(shape), check_status=True)
        indexes = self._scales_to_indexes(z_hat)
        y_hat = ops.rans_decode(ops.PackedStreams.from_list(strings[0], device), indexes[0].numel(), gc.coder_tables(),
                                indexes=indexes, want='values').view(indexes.size())
        return _run_synthesis(self, self.g_s, y_hat)

    def _get_means(self, x):
        medians = self.entropy_bottleneck._get_medians().detach()
        spatial_dims = x.dim() - 2
        medians = self.entropy_bottleneck._extend_ndims(medians, spatial_dims)
        return medians.expand(x.size(0), *([-1] * (spatial_dims + 1)))

    def _forward2train(self, x):
        y = self.g_a(x)
        z_hat, _ = self.entropy_bottleneck(self.h_a(torch.abs(y)))
        y_hat, _ = self.gaussian_conditional(y, self.h_s(z_hat))
        return self.g_s(y_hat)

    def forward(self, x):
        if not self.updated:
            return self._forward2train(x)
        if not self.training:
            return self.decode(**self.encode(x))
        y = self.g_a(x)
        gc = self.gaussian_conditional
        y_hat = gc.dequantize(gc.quantize(y, 'dequantize', self._get_means(y)))
        return self.g_s(y_hat.detach())

    def update(self, scale_table=None, force=False):
        if scale_table is None:
            scale_table = get_scale_table()
        updated = self.gaussian_conditional.update_scale_table(scale_table, force=force)
        updated |= super().update(force=force)
        self.updated = True
        return updated

    def load_state_dict(self, state_dict, **kwargs):
        update_registered_buffers(self.gaussian_conditional, 'gaussian_conditional',
                                  ['_quantized_cdf', '_offset', '_cdf_length', 'scale_table'], state_dict)
        return super().load_state_dict(state_dict)


@register_layer_class
class MSHPBasedResNetBottleneck(SHPBasedResNetBottleneck):
    """Mean-scale hyperprior bottleneck (mirrors sc2bench/models/layer.py:723-817).

    As the scale hyperprior, except: h_a reads y itself (not |y|) through LeakyReLU; h_s emits 2·C channels that split into
    (scales_hat, means_hat); y is quantised around means_hat (one mean per element) and dequantised by adding it back."""

    def __init__(self, num_input_channels=3, num_latent_channels=16, num_bottleneck_channels=24, num_target_channels=256,
                 g_a_channel_sizes=None, g_s_channel_sizes=None):
        L, b = num_latent_channels, num_bottleneck_channels
        h_a = nn.Sequential(
            nn.Conv2d(b, L, kernel_size=5, stride=2, padding=1, bias=False), nn.LeakyReLU(inplace=True),
            nn.Conv2d(L, L, kernel_size=5, stride=2, padding=2, bias=False))
        h_s = nn.Sequential(
            nn.ConvTranspose2d(L, L, kernel_size=5, stride=2, padding=1, bias=False), nn.LeakyReLU(inplace=True),
            nn.ConvTranspose2d(L, L * 3 // 2, kernel_size=5, stride=2, padding=1, bias=False), nn.LeakyReLU(inplace=True),
            nn.Conv2d(L * 3 // 2, b * 2, kernel_size=5, stride=1, padding=0, bias=False))
        super().__init__(num_input_channels=num_input_channels, num_latent_channels=L, num_bottleneck_channels=b,
                         num_target_channels=num_target_channels, h_a=h_a, h_s=h_s,
                         g_a_channel_sizes=g_a_channel_sizes, g_s_channel_sizes=g_s_channel_sizes)

    @torch.no_grad()
    def _gaussian_params(self, z_hat):
        scales_hat, means_hat = run_analysis(self, 'h_s', self.h_s, z_hat).chunk(2, 1)
        return self.gaussian_conditional.build_indexes(scales_hat), means_hat.contiguous()

    @torch.no_grad()
    def encode(self, x, **kwargs):
        eb, gc = self.entropy_bottleneck, self.gaussian_conditional
        y = self._analysis(x)
        z_symbols = run_analysis(self, 'h_a', self.h_a, y, medians=eb._get_medians().detach().reshape(-1), out='symbols')
        z_shape = z_symbols.size()[-2:]
        z_streams = eb.compress_symbols(z_symbols, spatial=z_symbols[0, 0].numel())
        z_hat = eb.decompress_packed(z_streams, tuple(z_shape))
        indexes, means_hat = self._gaussian_params(z_hat)
        y_symbols = ops.quantize_symbols(y, means_hat)
        y_streams = ops.rans_encode(y_symbols, gc.coder_tables(), indexes=indexes)
        return {'strings': [y_streams.tolist(), z_streams.tolist()], 'shape': z_shape}

    @torch.no_grad()
    def decode(self, strings, shape):
        assert isinstance(strings, list) and len(strings) == 2
        eb, gc = self.entropy_bottleneck, self.gaussian_conditional
        device = eb._quantized_cdf.device
        z_hat = eb.decompress_packed(ops.PackedStreams.from_list(strings[1], device), tuple(shape), check_status=True)
        indexes, means_hat = self._gaussian_params(z_hat)
        y_symbols = ops.rans_decode(ops.PackedStreams.from_list(strings[0], device), indexes[0].numel(), gc.coder_tables(),
                                    indexes=indexes, want='symbols').view(indexes.size())
        y_hat = ops.dequantize(y_symbols, means_hat)
        return _run_synthesis(self, self.g_s, y_hat)

    def _forward2train(self, x):
        y = self.g_a(x)
        z_hat, _ = self.entropy_bottleneck(self.h_a(y))
        scales_hat, means_hat = self.h_s(z_hat).chunk(2, 1)
        y_hat, _ = self.gaussian_conditional(y, scales_hat, means=means_hat)
        return self.g_s(y_hat)

    def forward(self, x):
        if not self.updated:
            return self._forward2train(x)
        if not self.training:
            return self.decode(**self.encode(x))
        eb, gc = self.entropy_bottleneck, self.gaussian_conditional
        y = self.g_a(x)
        z = self.h_a(y)
        z_hat = eb.dequantize(eb.quantize(z, 'dequantize', self._get_means(z)))
        scales_hat, means_hat = self.h_s(z_hat).chunk(2, 1)
        y_hat = gc.dequantize(gc.quantize(y, 'dequantize', means_hat))
        return self.g_s(y_hat.detach())
