"""CompressionModel and the two CompressAI zoo architectures on the path, with the CompressAI contract.

Drop-in for `compressai.models.CompressionModel` (sc2bench/models/layer.py:4,346,401; backbone.py:4,154,276),
`compressai.models.google.get_scale_table` (layer.py:5), `compressai.models.utils.update_registered_buffers`
(layer.py:6,714-719) and `compressai.zoo.image.{bmshj2018_factorized,bmshj2018_hyperprior,model_architectures}`
(sc2bench/models/registry.py:2,12-14,73).  `compress()` / `decompress()` run g_a / g_s and the coder through
libsc2b200.so; `forward()` (training) stays differentiable torch.
"""
import math
import warnings

import torch
from torch import nn

from . import _native, ops
from .entropy_models import EntropyBottleneck, GaussianConditional
from .layers import GDN

SCALES_MIN, SCALES_MAX, SCALES_LEVELS = 0.11, 256, 64


def get_scale_table(min=SCALES_MIN, max=SCALES_MAX, levels=SCALES_LEVELS):
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


def _resize_registered_buffer(module, buffer_name, state_dict_key, state_dict, policy, dtype):
    new_size = state_dict[state_dict_key].size()
    current = dict(module.named_buffers()).get(buffer_name)
    if policy in ('resize_if_empty', 'resize'):
        if current is None:
            raise RuntimeError(f'buffer "{buffer_name}" was not registered')
        if policy == 'resize' or current.numel() == 0:
            current.resize_(new_size)
    elif policy == 'register':
        if current is not None:
            raise RuntimeError(f'buffer "{buffer_name}" was already registered')
        module.register_buffer(buffer_name, torch.zeros(new_size, dtype=dtype))
    else:
        raise ValueError(f'Invalid policy "{policy}"')


def update_registered_buffers(module, module_name, buffer_names, state_dict, policy='resize_if_empty', dtype=torch.int):
    """Resizes the (initially empty) CDF buffers so that a checkpoint's tables can be loaded into them."""
    if not module:
        return
    valid = [n for n, _ in module.named_buffers()]
    for name in buffer_names:
        if name not in valid:
            raise ValueError(f'Invalid buffer name "{name}"')
    for name in buffer_names:
        _resize_registered_buffer(module, name, f'{module_name}.{name}', state_dict, policy, dtype)


def conv(in_channels, out_channels, kernel_size=5, stride=2):
    return nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=kernel_size // 2)


def deconv(in_channels, out_channels, kernel_size=5, stride=2):
    return nn.ConvTranspose2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                              output_padding=stride - 1, padding=kernel_size // 2)


def _single(v):
    if isinstance(v, (tuple, list)):
        if len(set(v)) != 1:
            raise NotImplementedError('anisotropic stride / padding is not on the bottleneck path')
        return int(v[0])
    return int(v)


@torch.no_grad()
def run_transform(seq, x, final_epilogue=_native.EPI_NONE, final_aux=None, in_abs=False):
    """Inference executor for an analysis / synthesis transform (an nn.Sequential of Conv2d / ConvTranspose2d /
    GDN / GDN1 / ReLU / LeakyReLU): every layer is one libsc2b200 launch, the activation folds into the conv before
    it, the last conv can take a fused epilogue (quantise-to-symbols, clamp) and the first can read |x| (in_abs)."""
    ops.require_cuda(x, 'run_transform')
    mods = list(seq)
    if in_abs and not (mods and isinstance(mods[0], (nn.Conv2d, nn.ConvTranspose2d))):
        x, in_abs = torch.abs(x), False
    i = 0
    while i < len(mods):
        m = mods[i]
        last = i == len(mods) - 1
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            if m.groups != 1 or _single(m.dilation) != 1:
                raise NotImplementedError('grouped / dilated convolutions are not on the bottleneck path')
            epi, aux, slope = _native.EPI_NONE, None, 0.0
            first = i == 0
            if i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU):
                epi = _native.EPI_RELU
                i += 1
                last = i == len(mods) - 1
            elif i + 1 < len(mods) and isinstance(mods[i + 1], nn.LeakyReLU):
                epi, slope = _native.EPI_LEAKY_RELU, mods[i + 1].negative_slope
                i += 1
                last = i == len(mods) - 1
            if last and final_epilogue != _native.EPI_NONE:
                if epi != _native.EPI_NONE:
                    raise NotImplementedError('cannot stack two epilogues')
                epi, aux = final_epilogue, final_aux
            tr = isinstance(m, nn.ConvTranspose2d)
            x = ops.conv2d(x, m.weight, m.bias, stride=_single(m.stride), padding=_single(m.padding), transposed=tr,
                           output_padding=_single(m.output_padding) if tr else 0, epilogue=epi, aux=aux,
                           in_abs=in_abs and first, epi_param=slope)
        elif isinstance(m, GDN):
            gamma, beta = m.effective_params()
            x = ops.gdn(x, gamma, beta, kind=m._kind, inverse=m.inverse)
        elif isinstance(m, nn.ReLU):
            x = torch.relu_(x) if x.is_floating_point() else x
        else:
            x = m(x)
        i += 1
    return x


class CompressionModel(nn.Module):
    """Base class holding entropy models; `update()`, `aux_loss()`, table-aware `load_state_dict()`."""

    def __init__(self, entropy_bottleneck_channels=None, init_weights=None):
        super().__init__()
        if entropy_bottleneck_channels is not None:
            # deprecated in CompressAI 1.2 but relied on by sc2bench (layer.py:356,409)
            self.entropy_bottleneck = EntropyBottleneck(entropy_bottleneck_channels)
        if init_weights is not None:
            warnings.warn('init_weights was removed.', DeprecationWarning, stacklevel=2)

    def load_state_dict(self, state_dict, strict=True):
        for name, module in self.named_modules():
            if not any(k.startswith(name) for k in state_dict.keys()):
                continue
            if isinstance(module, EntropyBottleneck):
                update_registered_buffers(module, name, ['_quantized_cdf', '_offset', '_cdf_length'], state_dict)
            if isinstance(module, GaussianConditional):
                update_registered_buffers(module, name, ['_quantized_cdf', '_offset', '_cdf_length', 'scale_table'], state_dict)
        return nn.Module.load_state_dict(self, state_dict, strict=strict)

    def update(self, scale_table=None, force=False):
        if scale_table is None:
            scale_table = get_scale_table()
        updated = False
        for _, module in self.named_modules():
            if isinstance(module, EntropyBottleneck):
                updated |= module.update(force=force)
            if isinstance(module, GaussianConditional):
                updated |= module.update_scale_table(scale_table, force=force)
        return updated

    def aux_loss(self):
        return sum(m.loss() for m in self.modules() if isinstance(m, EntropyBottleneck))


class SplitAnalysisPlan:
    """An analysis-side transform (Conv2d / GDN / GDN1 / ReLU / LeakyReLU in sequence: g_a and h_a of the CompressAI zoo codecs,
    compressai.models.google [mem], reached from sc2bench/models/wrapper.py:108-126) on the tcgen05 tensor cores in fp32-grade
    split-fp16 arithmetic (conv_tc_split.cu) -- the symbols must match the fp32 reference, so the fp16 route of g_s is not enough.
        first layer (3 input channels): im2col (sc2_patchify_split_nhwc) + a 1x1 GEMM over K = 75 -> 80
        Conv2d(k <= 5, stride 1 | 2) with bias, optional ReLU / LeakyReLU in the epilogue; stride-2 layers read the NHWC planes of the
            layer before through a 5-D tensor map (pixel parity = a coordinate), layers wider than 128 channels run as N tiles
        GDN / GDN1: a 1x1 gamma GEMM on x^2 / |x| formed in shared memory, y = x * rsqrt(beta + acc) / x / (beta + acc)
        ConvTranspose2d(k5, s2, p2, op1 | p1, op0) (hyper-synthesis h_s): 4 parity sub-convolutions writing interleaved pixels
        channel counts that are not multiples of 16 and odd sizes in front of a stride-2 layer are zero-padded (torch pads of small planes)
        the last conv can quantise straight to coder symbols (round(y + bias - median), NCHW order)."""

    def __init__(self, seq):
        self.seq = seq
        self._plan = None

    @staticmethod
    def why_not(seq, x_shape, planes_in=False):
        """None when the plan covers `seq` for an input of shape x_shape (NCHW), else the reason."""
        mods = list(seq)
        if not mods or not isinstance(mods[0], (nn.Conv2d, nn.ConvTranspose2d)):
            return 'does not start with a convolution'
        C, H, W = x_shape[-3:]
        prev_conv = False
        for i, m in enumerate(mods):
            if isinstance(m, nn.ConvTranspose2d):
                geo = (tuple(m.kernel_size), tuple(m.stride), tuple(m.padding), tuple(m.output_padding), m.groups, tuple(m.dilation))
                if geo not in (((5, 5), (2, 2), (2, 2), (1, 1), 1, (1, 1)), ((5, 5), (2, 2), (1, 1), (0, 0), 1, (1, 1))):
                    return 'transposed convolution %d is not (k5, s2, p2, op1) or (k5, s2, p1, op0)' % i
                if m.in_channels != C:
                    return 'transposed convolution %d expects %d channels, gets %d' % (i, m.in_channels, C)
                grow = 0 if m.padding[0] == 2 else 1
                H, W, C = 2 * H + grow, 2 * W + grow, m.out_channels
                prev_conv = True
            elif isinstance(m, nn.Conv2d):
                k, st = m.kernel_size[0], m.stride[0]
                if (m.groups != 1 or tuple(m.dilation) != (1, 1) or m.kernel_size[0] != m.kernel_size[1] or isinstance(m.padding, str)
                        or m.padding[0] != m.padding[1] or m.stride[0] != m.stride[1] or st not in (1, 2) or k > 5
                        or m.padding_mode != 'zeros'):
                    return 'convolution %d: groups / dilation / kernel / stride outside the kernel' % i
                if m.in_channels != C:
                    return 'convolution %d expects %d channels, gets %d' % (i, m.in_channels, C)
                H, W = (H + 2 * m.padding[0] - k) // st + 1, (W + 2 * m.padding[0] - k) // st + 1
                if H < 1 or W < 1:
                    return 'empty output'
                C = m.out_channels
                prev_conv = True
            elif isinstance(m, GDN):
                if m.inverse:
                    return 'layer %d is an inverse GDN' % i
                if C % 16 or m.beta.numel() != C:
                    return 'GDN over %d channels (not a multiple of 16)' % C
                prev_conv = False
            elif isinstance(m, (nn.ReLU, nn.LeakyReLU)):
                if not prev_conv:
                    return 'activation %d does not follow a convolution' % i
                prev_conv = False
            else:
                return 'layer %d is %s' % (i, type(m).__name__)
        return None

    def _prepare(self, planes_in):
        key = (tuple((q.data_ptr(), q._version, q.device) for q in self.seq.parameters()), planes_in)
        plan = self._plan
        if plan is not None and plan[0] == key:
            return plan
        steps, mods = [], list(self.seq)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                act, slope = _native.TCS_ACT_NONE, 0.0
                if i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU):
                    act = _native.TCS_ACT_RELU
                elif i + 1 < len(mods) and isinstance(mods[i + 1], nn.LeakyReLU):
                    act, slope = _native.TCS_ACT_LEAKY, mods[i + 1].negative_slope
                bias = m.bias.detach().float().contiguous() if m.bias is not None else None
                c_in_pad = (m.in_channels + 15) // 16 * 16  # K of the tensor-core kernel: planes are zero-padded to it
                if isinstance(m, nn.ConvTranspose2d):
                    steps.append(('deconv', m, ops.pack_deconv5_weight_split_tiles(m.weight, padding=m.padding[0], c_in_pad=c_in_pad),
                                  bias, act, slope, None))
                else:
                    patches = i == 0 and not planes_in and m.in_channels * m.kernel_size[0] ** 2 <= 128
                    k_pad = (m.in_channels * m.kernel_size[0] ** 2 + 15) // 16 * 16 if patches else None
                    tiles = ops.pack_conv_weight_split_tiles(m.weight, c_in_pad=k_pad if patches else c_in_pad, as_patches=patches)
                    steps.append(('conv', m, tiles, bias, act, slope, k_pad))
                i += 2 if act != _native.TCS_ACT_NONE else 1
            else:
                gamma, beta = m.effective_params()
                C = beta.numel()
                tiles = ops.pack_conv_weight_split_tiles(gamma.detach().reshape(C, C, 1, 1))
                steps.append(('gdn', m, tiles, beta.detach().float().contiguous(), _native.TCS_GDN1 if m._kind == 0 else _native.TCS_GDN))
                i += 1
        plan = (key, steps)
        self._plan = plan
        return plan

    @torch.no_grad()
    def __call__(self, x, medians=None, out='nchw'):
        """x: fp32 NCHW, or split planes (hi, lo) [B, H, W, C].  out: 'symbols' (the last layer must be a Conv2d; int32 NCHW,
        round(y - medians)) | 'nchw' (fp32) | 'planes' (hi, lo)."""
        planes_in = isinstance(x, tuple)
        _, steps = self._prepare(planes_in)
        T = _native
        h = l = None
        if planes_in:
            h, l = x
        elif steps[0][6] is None:  # an fp32 NCHW activation (not an image): split NHWC planes
            h, l = ops.split_f16(x.permute(0, 2, 3, 1))

        def fit(h, l, c_in, even):
            # zero-pad the planes to the K the kernel wants (c_in rounded up to 16) and, in front of a stride-2 layer, to even sizes
            # (the pad row / column is the zero the convolution's own padding would have read)
            pc = (c_in + 15) // 16 * 16 - h.shape[3]
            ph, pw = (h.shape[1] & 1, h.shape[2] & 1) if even else (0, 0)
            if pc < 0:
                h, l, pc = h[..., :(c_in + 15) // 16 * 16], l[..., :(c_in + 15) // 16 * 16], 0
            if pc or ph or pw:
                h, l = torch.nn.functional.pad(h, (0, pc, 0, pw, 0, ph)), torch.nn.functional.pad(l, (0, pc, 0, pw, 0, ph))
            return h.contiguous(), l.contiguous()

        for si, step in enumerate(steps):
            last = si == len(steps) - 1
            if step[0] == 'conv':
                _, m, tiles, bias, act, slope, k_pad = step
                k, st, pad = m.kernel_size[0], m.stride[0], m.padding[0]
                mode = T.TCS_QUANT if (last and out == 'symbols') else T.TCS_STORE
                med = medians if mode == T.TCS_QUANT else None
                if k_pad is not None:
                    ph, pl = ops.patchify_split_nhwc(x, k, k, st, pad, k_pad)
                    res = ops.tc_split_conv_tiled(ph, pl, tiles, 1, 1, 1, 0, mode, vec=bias, medians=med, act=act, slope=slope,
                                                  name='tcs_conv_first')
                else:
                    H, W = h.shape[1], h.shape[2]
                    h, l = fit(h, l, m.in_channels, st == 2)
                    res = ops.tc_split_conv_tiled(h, l, tiles, k, k, st, pad, mode, vec=bias, medians=med, act=act, slope=slope,
                                                  in_nhwc=st == 2, name='tcs_conv',
                                                  out_hw=((H + 2 * pad - k) // st + 1, (W + 2 * pad - k) // st + 1))
                if mode == T.TCS_QUANT:
                    return res
                h, l = res
            elif step[0] == 'deconv':
                _, m, packs, bias, act, slope, _ = step
                h, l = fit(h, l, m.in_channels, False)
                B, H, W, _ = h.shape
                pitch = (m.out_channels + 7) // 8 * 8
                grow = 0 if m.padding[0] == 2 else 1  # (k5, s2, p1, op0): 2H + 1 outputs
                out = (torch.empty((B, 2 * H + grow, 2 * W + grow, pitch), dtype=torch.float16, device=h.device),
                       torch.empty((B, 2 * H + grow, 2 * W + grow, pitch), dtype=torch.float16, device=h.device))
                taps = ops.deconv5_parity_taps(m.padding[0])
                for (py, px), tiles in packs.items():
                    (ky, pad_y), (kx, pad_x) = taps[py], taps[px]
                    ops.tc_split_conv_tiled(h, l, tiles, len(ky), len(kx), 1, pad_y, T.TCS_STORE, vec=bias, act=act, slope=slope,
                                            pad_x=pad_x, out=out, out_parity=(py, px), name='tcs_deconv5')
                h, l = out
            else:
                _, m, tiles, beta, mode = step
                h, l = ops.tc_split_conv_tiled(h, l, tiles, 1, 1, 1, 0, mode, vec=beta, gdn_x=(h, l), name='tcs_gdn')
        if out == 'symbols':
            raise NotImplementedError('the transform does not end with a convolution')
        if out == 'planes':
            return h, l
        c_out = [st for st in steps if st[0] in ('conv', 'deconv')][-1][1].out_channels
        return ops.unsplit_to_nchw(h, l, c_out)


def run_analysis(model, name, seq, x, medians=None, out='nchw', in_abs=False):
    """g_a / h_a of a zoo codec: the split tensor-core plan when it covers `seq` (and model.encoder_precision allows), else the fp32
    CUDA-core kernels (logged once).  x: fp32 NCHW, or split planes (hi, lo) [B, H, W, C] from a previous plan."""
    planes_in = isinstance(x, tuple)
    shape = (x[0].shape[3], x[0].shape[1], x[0].shape[2]) if planes_in else tuple(x.shape[-3:])
    prec = getattr(model, 'encoder_precision', 'split-tc')
    why = 'encoder_precision = %r' % prec if prec != 'split-tc' else SplitAnalysisPlan.why_not(seq, shape, planes_in=planes_in)
    if why is None:
        plans = model.__dict__.setdefault('_tc_analysis', {})
        plan = plans.get(name)
        if plan is None:
            plan = plans[name] = SplitAnalysisPlan(seq)
        if in_abs:
            x = ops.abs_split(*x) if planes_in else torch.abs(x)
        return plan(x, medians=medians, out=out)
    _warn_fp32(model, name, why)
    if planes_in:
        c = list(seq)[0].in_channels
        x = ops.unsplit_to_nchw(x[0], x[1], c)
    if out == 'symbols':
        return run_transform(seq, x, final_epilogue=_native.EPI_QUANTIZE, final_aux=medians, in_abs=in_abs)
    y = run_transform(seq, x, in_abs=in_abs)
    if out == 'planes':
        c = y.shape[1]
        return ops.split_f16(torch.nn.functional.pad(y.permute(0, 2, 3, 1), (0, -c % 8)))
    return y


class ZooSynthesisPlan:
    """g_s of the CompressAI zoo codecs (deconv - IGDN - deconv - IGDN - deconv - IGDN - deconv, compressai.models.google [mem];
    reached from sc2bench/models/wrapper.py:130) on the tcgen05 kernels, fp16 operands / fp32 accumulation like the bottleneck's
    g_s (the 1e-3 tolerance on x_hat):
        ConvTranspose2d(k5, s2, p2, op1) = 4 parity sub-convolutions (ops.pack_deconv5_weight_f16), bias in the epilogue;
        the deconv in front of an inverse GDN stores x and x^2 / 256; the GDN's 1x1 gamma GEMM reads the squares,
        y = x * sqrt(beta + 256 * acc); the last deconv writes clamp(x_hat, 0, 1) as fp32 NCHW.
    15 launches instead of 7 CUDA-core convolutions that ran at 0.1-1 % of the tensor peak."""

    def __init__(self, seq):
        self.seq = seq
        self._plan = None
        self.split_last_stage = True

    @staticmethod
    def why_not(seq):
        mods = list(seq)
        if len(mods) < 1 or len(mods) % 2 == 0:
            return 'not deconv (- IGDN - deconv)*'
        for i, m in enumerate(mods):
            if i % 2 == 0:
                if not isinstance(m, nn.ConvTranspose2d):
                    return 'layer %d is %s' % (i, type(m).__name__)
                if (tuple(m.kernel_size), tuple(m.stride), tuple(m.padding), tuple(m.output_padding), m.groups, tuple(m.dilation)) != \
                        ((5, 5), (2, 2), (2, 2), (1, 1), 1, (1, 1)):
                    return 'transposed convolution %d is not (k5, s2, p2, op1)' % i
                if m.in_channels % 64:
                    return 'c_in %d is not a multiple of 64' % m.in_channels
                last = i == len(mods) - 1
                if last and m.out_channels > 32:
                    return 'last layer has %d > 32 output channels' % m.out_channels
                if not last and m.out_channels % 64:
                    return 'c_out %d is not a multiple of 64' % m.out_channels
            else:
                if type(m) is not GDN or not m.inverse:
                    return 'layer %d is not an inverse GDN' % i
                if m.beta.numel() % 64 or m.beta.numel() > 512:
                    return 'GDN over %d channels' % m.beta.numel()
        return None

    def _prepare(self):
        key = tuple((q.data_ptr(), q._version, q.device) for q in self.seq.parameters())
        plan = self._plan
        if plan is not None and plan[0] == key:
            return plan
        steps, mods = [], list(self.seq)
        for i, m in enumerate(mods):
            if i % 2 == 0:
                last = i == len(mods) - 1
                packs = ops.pack_deconv5_weight_f16(m.weight, rows_pad=32 if last else None)
                bias = None
                if m.bias is not None:
                    bias = torch.zeros(32 if last else m.out_channels, dtype=torch.float32, device=m.weight.device)
                    bias[:m.out_channels] = m.bias.detach().float()
                steps.append(('deconv', packs, bias, m.in_channels, m.out_channels, last))
            else:
                gamma, beta = m.effective_params()
                C = beta.numel()
                steps.append(('igdn', gamma.detach().reshape(1, C, C).half().contiguous(), beta.detach().float().contiguous(), C))
        plan = (key, steps, mods[0].in_channels)
        self._plan = plan
        return plan

    @torch.no_grad()
    def __call__(self, y_hat):
        """fp32 NCHW latent -> clamp(x_hat, 0, 1) as fp32 NCHW"""
        _, steps, c_in0 = self._prepare()
        T = _native
        x = ops.nchw_to_nhwc_f16(y_hat, c_in0)
        sq = x_lo = None
        n_deconv = sum(1 for st in steps if st[0] == 'deconv')
        seen = 0
        for step in steps:
            if step[0] == 'deconv':
                _, packs, bias, c_in, c_out, last = step
                seen += 1
                # The LAST stage carries split activations (x = hi + lo): fp16 rounding of the last deconv's input and of the
                # last IGDN's operands is what costs 1.5e-3 on x_hat (6e-4 without; every earlier rounding is harmless).
                split_out = self.split_last_stage and seen == n_deconv - 1
                B, H, W, _ = x.shape
                lo_out = None
                if last:
                    out = torch.empty((B, c_out, 2 * H, 2 * W), dtype=torch.float32, device=x.device)
                    sq = None
                else:
                    out = torch.empty((B, 2 * H, 2 * W, c_out), dtype=torch.float16, device=x.device)
                    sq = torch.empty_like(out)
                    lo_out = torch.empty_like(out) if split_out else None
                for (py, px), w in packs.items():
                    (ky, pad_y), (kx, pad_x) = ops.DECONV5_TAPS[py], ops.DECONV5_TAPS[px]
                    ops.tc_conv_ex(x, w, len(ky), len(kx), pad_y, pad_x, T.TC_NCHW_F32_CLAMP if last else T.TC_STORE_SQ_F16, (H, W), out,
                                   out_stride=2, out_py=py, out_px=px, vec=bias, out2=sq, c_out=c_out, c_in=c_in, x_lo=x_lo, out3=lo_out,
                                   tag='tc_deconv5[%d->%d,p%d%d%s]' % (c_in, c_out, py, px, ',2pass' if x_lo is not None else ''))
                x, x_lo = out, lo_out
            else:
                _, gamma, beta, C = step
                B, H, W, _ = x.shape
                out = torch.empty_like(x)
                out_lo = torch.empty_like(x) if x_lo is not None else None
                ops.tc_conv_ex(sq, gamma, 1, 1, 0, 0, T.TC_IGDN_SQ_F16, (H, W), out, vec=beta, gdn_x=x, c_out=C, c_in=C,
                               gdn_x_lo=x_lo, out2=out_lo, tag='tc_igdn[%d%s]' % (C, ',split' if x_lo is not None else ''))
                x, x_lo, sq = out, out_lo, None
        return x


_warned_fallback = set()


def run_synthesis(model, seq, y_hat):
    """g_s of a zoo codec: the tensor-core plan when it covers `seq` (and model.decoder_precision allows), else the fp32 kernels."""
    why = 'decoder_precision = %r' % getattr(model, 'decoder_precision', 'fp16-tc') if getattr(model, 'decoder_precision', 'fp16-tc') != 'fp16-tc' \
        else ZooSynthesisPlan.why_not(seq)
    if why is None:
        plan = model.__dict__.get('_tc_synthesis')
        if plan is None:
            plan = ZooSynthesisPlan(seq)
            model.__dict__['_tc_synthesis'] = plan
        return plan(y_hat)
    _warn_fp32(model, 'g_s', why)
    return run_transform(seq, y_hat, final_epilogue=_native.EPI_CLAMP01)


def _warn_fp32(model, name, why):
    key = (type(model).__name__, name, why)
    if key not in _warned_fallback:
        _warned_fallback.add(key)
        import logging
        logging.getLogger('sc2bench_b200').warning('%s.%s runs on the fp32 CUDA-core kernels (conv2d_f32_kernel), not on the tensor cores: %s',
                                                    type(model).__name__, name, why)


class FactorizedPrior(CompressionModel):
    """bmshj2018-factorized: g_a (4 conv, 3 GDN) -> EntropyBottleneck -> g_s (4 deconv, 3 IGDN)."""

    def __init__(self, N, M, **kwargs):
        super().__init__(**kwargs)
        self.entropy_bottleneck = EntropyBottleneck(M)
        self.g_a = nn.Sequential(conv(3, N), GDN(N), conv(N, N), GDN(N), conv(N, N), GDN(N), conv(N, M))
        self.g_s = nn.Sequential(deconv(M, N), GDN(N, inverse=True), deconv(N, N), GDN(N, inverse=True),
                                 deconv(N, N), GDN(N, inverse=True), deconv(N, 3))
        self.N, self.M = N, M

    @property
    def downsampling_factor(self):
        return 2 ** 4

    def forward(self, x):
        y = self.g_a(x)
        y_hat, y_likelihoods = self.entropy_bottleneck(y)
        return {'x_hat': self.g_s(y_hat), 'likelihoods': {'y': y_likelihoods}}

    @torch.no_grad()
    def compress_packed(self, x):
        eb = self.entropy_bottleneck
        medians = eb._get_medians().detach().reshape(-1)
        symbols = run_analysis(self, 'g_a', self.g_a, x, medians=medians, out='symbols')
        return eb.compress_symbols(symbols, spatial=symbols[0, 0].numel()), symbols.size()[-2:]

    def compress(self, x):
        streams, shape = self.compress_packed(x)
        return {'strings': [streams.tolist()], 'shape': shape}

    @torch.no_grad()
    def decompress(self, strings, shape):
        assert isinstance(strings, (list, tuple)) and len(strings) == 1
        first = strings[0]
        streams = first if isinstance(first, ops.PackedStreams) else \
            ops.PackedStreams.from_list(first, self.entropy_bottleneck._quantized_cdf.device)
        y_hat = self.entropy_bottleneck.decompress_packed(streams, tuple(shape), check_status=not isinstance(first, ops.PackedStreams))
        return {'x_hat': run_synthesis(self, self.g_s, y_hat)}


class ScaleHyperprior(CompressionModel):
    """bmshj2018-hyperprior: adds h_a / h_s and a Gaussian conditional whose scales are decoded from z."""

    def __init__(self, N, M, **kwargs):
        super().__init__(**kwargs)
        self.entropy_bottleneck = EntropyBottleneck(N)
        self.g_a = nn.Sequential(conv(3, N), GDN(N), conv(N, N), GDN(N), conv(N, N), GDN(N), conv(N, M))
        self.g_s = nn.Sequential(deconv(M, N), GDN(N, inverse=True), deconv(N, N), GDN(N, inverse=True),
                                 deconv(N, N), GDN(N, inverse=True), deconv(N, 3))
        self.h_a = nn.Sequential(conv(M, N, stride=1, kernel_size=3), nn.ReLU(inplace=True), conv(N, N),
                                 nn.ReLU(inplace=True), conv(N, N))
        self.h_s = nn.Sequential(deconv(N, N), nn.ReLU(inplace=True), deconv(N, N), nn.ReLU(inplace=True),
                                 conv(N, M, stride=1, kernel_size=3), nn.ReLU(inplace=True))
        self.gaussian_conditional = GaussianConditional(None)
        self.N, self.M = int(N), int(M)

    @property
    def downsampling_factor(self):
        return 2 ** (4 + 2)

    def forward(self, x):
        y = self.g_a(x)
        z = self.h_a(torch.abs(y))
        z_hat, z_likelihoods = self.entropy_bottleneck(z)
        scales_hat = self.h_s(z_hat)
        y_hat, y_likelihoods = self.gaussian_conditional(y, scales_hat)
        return {'x_hat': self.g_s(y_hat), 'likelihoods': {'y': y_likelihoods, 'z': z_likelihoods}}

    @torch.no_grad()
    def compress(self, x):
        (y_streams, z_streams), shape = self.compress_packed(x)
        return {'strings': [y_streams.tolist(), z_streams.tolist()], 'shape': shape}

    @torch.no_grad()
    def compress_packed(self, x):
        """compress() with the bitstreams left on the device: ((y PackedStreams, z PackedStreams), latent (H, W)); nothing synchronises,
        so batches issued on different CUDA streams overlap (a batch's coder is a serial chain per image)."""
        eb, gc = self.entropy_bottleneck, self.gaussian_conditional
        y_planes = run_analysis(self, 'g_a', self.g_a, x, out='planes')
        y = ops.unsplit_to_nchw(y_planes[0], y_planes[1], self.M)
        z_symbols = run_analysis(self, 'h_a', self.h_a, y_planes, medians=eb._get_medians().detach().reshape(-1), out='symbols',
                                 in_abs=True)
        z_streams = eb.compress_symbols(z_symbols, spatial=z_symbols[0, 0].numel())
        # the encoder decodes z itself so that both sides derive the scales from identical values
        z_hat = eb.decompress_packed(z_streams, tuple(z_symbols.size()[-2:]))
        indexes = gc.build_indexes(run_analysis(self, 'h_s', self.h_s, z_hat))
        y_symbols = ops.quantize_symbols(y.reshape(y.size(0), 1, -1))
        y_streams = ops.rans_encode(y_symbols, gc.coder_tables(), indexes=indexes)
        return (y_streams, z_streams), z_symbols.size()[-2:]

    @torch.no_grad()
    def decompress(self, strings, shape):
        """strings: [y, z] as list[bytes] each (the CompressAI contract; decode faults raise here) or as PackedStreams from
        compress_packed (device resident; faults accumulate in the entropy bottleneck's fault word, see check_faults())."""
        assert isinstance(strings, (list, tuple)) and len(strings) == 2
        eb, gc = self.entropy_bottleneck, self.gaussian_conditional
        device = eb._quantized_cdf.device
        packed = isinstance(strings[0], ops.PackedStreams) and isinstance(strings[1], ops.PackedStreams)
        z_streams = strings[1] if packed else ops.PackedStreams.from_list(strings[1], device)
        y_streams = strings[0] if packed else ops.PackedStreams.from_list(strings[0], device)
        z_hat = eb.decompress_packed(z_streams, tuple(shape), check_status=not packed)
        indexes = gc.build_indexes(run_analysis(self, 'h_s', self.h_s, z_hat))
        y_hat = ops.rans_decode(y_streams, indexes[0].numel(), gc.coder_tables(), indexes=indexes, want='values', check_status=not packed,
                                status=eb._fault_word(device) if packed else None)
        return {'x_hat': run_synthesis(self, self.g_s, y_hat.view(indexes.size()))}


# ---- zoo (compressai.zoo.image) --------------------------------------------------------------------
_QUALITY_TO_NM = {q: (128, 192) for q in range(1, 6)}
_QUALITY_TO_NM.update({q: (192, 320) for q in range(6, 9)})
model_architectures = {'bmshj2018-factorized': FactorizedPrior, 'bmshj2018-hyperprior': ScaleHyperprior}


def _zoo(arch, quality, metric, pretrained, **kwargs):
    if metric not in ('mse', 'ms-ssim'):
        raise ValueError(f'Invalid metric "{metric}"')
    if quality not in _QUALITY_TO_NM:
        raise ValueError(f'Invalid quality "{quality}", should be between (1, 8)')
    if pretrained:
        raise RuntimeError('pretrained CompressAI weights have to be downloaded; load a local checkpoint with '
                           'load_state_dict() instead (no network access here)')
    return model_architectures[arch](*_QUALITY_TO_NM[quality], **kwargs)


def bmshj2018_factorized(quality, metric='mse', pretrained=False, progress=True, **kwargs):
    return _zoo('bmshj2018-factorized', quality, metric, pretrained, **kwargs)


def bmshj2018_hyperprior(quality, metric='mse', pretrained=False, progress=True, **kwargs):
    return _zoo('bmshj2018-hyperprior', quality, metric, pretrained, **kwargs)
