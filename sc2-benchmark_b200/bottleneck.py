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
import torch
from torch import nn

from . import _native, ops
from .entropy_models import GaussianConditional
from .layers import GDN1
from .models import CompressionModel, get_scale_table, run_transform, update_registered_buffers



class TensorCoreTransform:
    """Execution plan that runs a stride-1 `Conv2d / GDN1` transform (the bottleneck's synthesis transform g_s) on the
    tcgen05 kernels: activations NHWC fp16 between layers, fp32 accumulation in TMEM, GDN1 fused with its 1x1 gamma
    GEMM, last layer written as fp32.  Weights are repacked once and re-packed when the parameters change."""

    def __init__(self, seq):
        self.seq = seq
        self._key = None
        self._steps = None

    @staticmethod
    def supports(seq):
        mods = list(seq)
        if not mods or not isinstance(mods[-1], nn.Conv2d):
            return False
        for m in mods:
            if isinstance(m, nn.Conv2d):
                k = m.kernel_size
                if (m.bias is not None or m.groups != 1 or tuple(m.stride) != (1, 1) or tuple(m.dilation) != (1, 1)
                        or k[0] != k[1] or m.padding[0] != m.padding[1] or isinstance(m.padding, str) or m.out_channels % 64):
                    return False
            elif type(m) is GDN1:
                if m.beta.numel() % 64 or m.beta.numel() > 512:  # (the kernel keeps beta in shared memory: <= 512 channels)
                    return False
            else:
                return False
        return True

    def _prepare(self):
        params = [p for p in self.seq.parameters()]
        key = tuple((p.data_ptr(), p._version, p.device) for p in params)
        if key == self._key:
            return
        steps, mods = [], list(self.seq)
        for i, m in enumerate(mods):
            if isinstance(m, nn.Conv2d):
                c_in_pad = (m.in_channels + 63) // 64 * 64
                mode = _native.TC_STORE_F32 if i == len(mods) - 1 else _native.TC_STORE_F16
                steps.append(('conv', ops.pack_conv_weight_f16(m.weight, c_in_pad), m.kernel_size[0], m.padding[0], mode, None))
            else:
                gamma, beta = m.effective_params()
                C = beta.numel()
                mode = _native.TC_IGDN1_F16 if m.inverse else _native.TC_GDN1_F16
                steps.append(('gdn', gamma.detach().reshape(1, C, C).half().contiguous(), 1, 0, mode, beta.detach().float().contiguous()))
        self._steps, self._key = steps, key
        self.c_in_pad = (mods[0].in_channels + 63) // 64 * 64

    @torch.no_grad()
    def __call__(self, x_nchw):
        """fp32 NCHW in -> fp32 output, logically NCHW (physically channels-last: what cuDNN prefers for the tail)."""
        self._prepare()
        x = ops.nchw_to_nhwc_f16(x_nchw, self.c_in_pad)
        for kind, w, k, pad, mode, beta in self._steps:
            x = ops.tc_conv(x, w, k, k, pad, mode=mode, beta=beta, gdn_x=x if kind == 'gdn' else None)
        return x.permute(0, 3, 1, 2)


class TensorCoreAnalysis:
    """Execution plan for the bottleneck's analysis transform g_a (Conv s2 - GDN1 - Conv s2 - GDN1 - Conv s1) on the
    fp32-grade "split fp16" tcgen05 kernels, ending in the fused quantise-to-symbols epilogue:

        image (fp32 NCHW) --patchify--> patches (parity-plane order) --1x1 GEMM--> x1 --GDN1--> y1 (parity planes)
        --5x5 s2 conv as 25 shifted boxes--> x2 --GDN1--> y2 --2x2 conv + round(y - median)--> int32 symbols (NCHW)

    Every intermediate is a (hi, lo) pair of NHWC fp16 planes.  Weights / gammas are split and packed once."""

    def __init__(self, seq):
        self.seq = seq
        self._key = None
        self.fuse_first_layer = True

    @staticmethod
    def supports(seq, x_shape):
        mods = list(seq)
        if len(mods) != 5 or not all(isinstance(mods[i], nn.Conv2d) for i in (0, 2, 4)) or not all(type(mods[i]) is GDN1 for i in (1, 3)):
            return False
        c1, c2, c3 = mods[0], mods[2], mods[4]
        for c in (c1, c2, c3):
            if (c.bias is not None or c.groups != 1 or tuple(c.dilation) != (1, 1) or c.kernel_size[0] != c.kernel_size[1]
                    or isinstance(c.padding, str) or c.padding[0] != c.padding[1] or c.stride[0] != c.stride[1]):
                return False
        if c1.stride[0] != 2 or c2.stride[0] != 2 or c3.stride[0] != 1 or mods[1].inverse or mods[3].inverse:
            return False
        if c1.in_channels * c1.kernel_size[0] ** 2 > 128 or c2.kernel_size[0] ** 2 > 25 or c3.kernel_size[0] ** 2 > 25:
            return False
        if c1.out_channels % 16 or c2.out_channels % 16 or max(c1.out_channels, c2.out_channels, c3.out_channels) > 128:
            return False
        H, W = x_shape[-2:]
        k, p = c1.kernel_size[0], c1.padding[0]
        h1, w1 = (H + 2 * p - k) // 2 + 1, (W + 2 * p - k) // 2 + 1
        return h1 >= 2 and w1 >= 2 and h1 % 2 == 0 and w1 % 2 == 0

    def _prepare(self):
        params = list(self.seq.parameters())
        key = tuple((q.data_ptr(), q._version, q.device) for q in params)
        if key == self._key:
            return
        c1, g1, c2, g2, c3 = list(self.seq)
        self.k1_pad = (c1.in_channels * c1.kernel_size[0] ** 2 + 15) // 16 * 16
        self.w1 = ops.pack_conv_weight_split(c1.weight, c_in_pad=self.k1_pad, as_patches=True)
        self.w2 = ops.pack_conv_weight_split(c2.weight)
        self.w3 = ops.pack_conv_weight_split(c3.weight)
        self.gdn = []
        for g in (g1, g2):
            gamma, beta = g.effective_params()
            C = beta.numel()
            self.gdn.append((ops.pack_conv_weight_split(gamma.detach().reshape(C, C, 1, 1)), beta.detach().float().contiguous()))
        self._key = key

    @torch.no_grad()
    def __call__(self, x, medians):
        self._prepare()
        c1, _, c2, _, c3 = list(self.seq)
        T = _native
        if c1.out_channels <= 96 and self.fuse_first_layer:
            # im2col fused into the kernel: no patch tensor in HBM
            h, l = ops.tc_first_layer(x, self.w1[0], self.w1[1], c1.out_channels, c1.kernel_size[0], c1.kernel_size[0], c1.padding[0])
        else:
            ph, pl = ops.patchify_split(x, c1.kernel_size[0], c1.kernel_size[0], 2, c1.padding[0], self.k1_pad)
            h, l = ops.tc_split_conv(ph, pl, self.w1[0], self.w1[1], c1.out_channels, 1, 1, 1, 0, T.TCS_STORE)
        (gh, gl), beta = self.gdn[0]
        h, l = ops.tc_split_conv(h, l, gh, gl, c1.out_channels, 1, 1, 1, 0, T.TCS_GDN1, beta=beta, gdn=True)
        h, l = ops.tc_split_conv(h, l, self.w2[0], self.w2[1], c2.out_channels, c2.kernel_size[0], c2.kernel_size[0], 2,
                                 c2.padding[0], T.TCS_STORE)
        (gh, gl), beta = self.gdn[1]
        h, l = ops.tc_split_conv(h, l, gh, gl, c2.out_channels, 1, 1, 1, 0, T.TCS_GDN1, beta=beta, gdn=True)
        return ops.tc_split_conv(h, l, self.w3[0], self.w3[1], c3.out_channels, c3.kernel_size[0], c3.kernel_size[0], 1,
                                 c3.padding[0], T.TCS_QUANT, medians=medians)


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
        # batches in flight: the coder layout that leaves the SMs to the transforms (sc2_rans_encode_batch, `layout`)
        self.entropy_bottleneck.coder_layout = 'lanes' if self._transform_stream is not None else None
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
    def encode_packed(self, x):
        """g_a + quantise + rANS, bitstreams left on the device: (PackedStreams, latent (H, W))."""
        eb = self.entropy_bottleneck
        symbols = self._on_transform_stream(lambda: self.analyze_to_symbols(x), x)
        return eb.compress_symbols(symbols, spatial=symbols[0, 0].numel()), symbols.size()[-2:]

    @torch.no_grad()
    def analyze_to_symbols(self, x):
        """g_a + round(y - median) on the device: image batch -> int32 symbols [B, C, H, W] (coder order)."""
        ops.require_cuda(x, 'FPBasedResNetBottleneck.encode')
        medians = self.entropy_bottleneck._get_medians().detach().reshape(-1)
        if self.encoder_precision == 'split-tc' and TensorCoreAnalysis.supports(self.encoder, x.shape):
            if self._tc_encoder is None:
                self._tc_encoder = TensorCoreAnalysis(self.encoder)
            return self._tc_encoder(x, medians)
        return run_transform(self.encoder, x, final_epilogue=_native.EPI_QUANTIZE, final_aux=medians)

    @torch.no_grad()
    def synthesize(self, latent_hat):
        """g_s on the device: dequantised latent (fp32 NCHW) -> decoder features."""
        if self.decoder_precision == 'fp16-tc' and TensorCoreTransform.supports(self.decoder):
            if self._tc_decoder is None:
                self._tc_decoder = TensorCoreTransform(self.decoder)
            return self._tc_decoder(latent_hat)
        return run_transform(self.decoder, latent_hat)

    @torch.no_grad()
    def decode_packed(self, streams, shape, check_status=False):
        latent_hat = self.entropy_bottleneck.decompress_packed(streams, tuple(shape), check_status=check_status)
        return self._on_transform_stream(lambda: self.synthesize(latent_hat), latent_hat)

    def encode(self, x, **kwargs):
        """-> {'strings': [list of B bytes objects], 'shape': latent (H, W)}  (reference contract, layer.py:496-507)"""
        streams, shape = self.encode_packed(x)
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
    def _scales_to_indexes(self, z_hat):
        return self.gaussian_conditional.build_indexes(run_transform(self.h_s, z_hat))

    @torch.no_grad()
    def encode(self, x, **kwargs):
        eb, gc = self.entropy_bottleneck, self.gaussian_conditional
        y = run_transform(self.g_a, x)
        z_symbols = run_transform(self.h_a, y, final_epilogue=_native.EPI_QUANTIZE,
                                  final_aux=eb._get_medians().detach().reshape(-1), in_abs=True)  # h_a(|y|), |.| on load
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
        if self.decoder_precision == 'fp16-tc' and TensorCoreTransform.supports(self.g_s):
            if self._tc_decoder is None:
                self._tc_decoder = TensorCoreTransform(self.g_s)
            return self._tc_decoder(y_hat)
        return run_transform(self.g_s, y_hat)

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
        scales_hat, means_hat = run_transform(self.h_s, z_hat).chunk(2, 1)
        return self.gaussian_conditional.build_indexes(scales_hat), means_hat.contiguous()

    @torch.no_grad()
    def encode(self, x, **kwargs):
        eb, gc = self.entropy_bottleneck, self.gaussian_conditional
        y = run_transform(self.g_a, x)
        z_symbols = run_transform(self.h_a, y, final_epilogue=_native.EPI_QUANTIZE,
                                  final_aux=eb._get_medians().detach().reshape(-1))
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
        if self.decoder_precision == 'fp16-tc' and TensorCoreTransform.supports(self.g_s):
            if self._tc_decoder is None:
                self._tc_decoder = TensorCoreTransform(self.g_s)
            return self._tc_decoder(y_hat)
        return run_transform(self.g_s, y_hat)

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
