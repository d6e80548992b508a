"""EntropyBottleneck / GaussianConditional with the CompressAI call contract, coder on the GPU.

Drop-in for the classes sc2bench imports from `compressai.entropy_models` (sc2bench/models/layer.py:2) and
reaches through `CompressionModel` (layer.py:346,401): same constructor arguments, parameter / buffer names
(`quantiles`, `matrices.N`, `biases.N`, `factors.N`, `_offset`, `_quantized_cdf`, `_cdf_length`), same
`compress` / `decompress` / `forward` / `update` / `loss` signatures and the same `ValueError`s.

What differs underneath: `compress` / `decompress` do not loop over samples through Python
lists on the CPU; the whole batch is quantised, rANS-coded and packed on the device by libsc2b200.so and
crosses PCIe once.  The differentiable / training-time branches (`forward`, `quantize("noise")`, `loss`) are
off the hot path and stay plain torch ops.  Tables are always built on the host in fp32 so that they are
bit-reproducible (checkpoints carry them as buffers anyway).
"""
import math
import threading

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import ops

_FAULT_LOCK = threading.Lock()


class _LowerBoundFn(torch.autograd.Function):
    """max(x, bound) whose gradient also passes where it pulls x back above the bound."""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, grad_output):
        x, bound = ctx.saved_tensors
        keep = (x >= bound) | (grad_output < 0)
        return keep.type(grad_output.dtype) * grad_output, None


class LowerBound(nn.Module):
    def __init__(self, bound):
        super().__init__()
        self.register_buffer('bound', torch.Tensor([float(bound)]))

    def forward(self, x):
        return _LowerBoundFn.apply(x, self.bound)


class EntropyModel(nn.Module):
    """Shared machinery: quantisation, CDF buffers, batched device coder."""

    def __init__(self, likelihood_bound=1e-9, entropy_coder=None, entropy_coder_precision=16):
        super().__init__()
        if entropy_coder not in (None, 'ans'):
            raise ValueError(f'Unknown entropy coder "{entropy_coder}" (only the rANS coder exists here)')
        self.entropy_coder_precision = int(entropy_coder_precision)
        if self.entropy_coder_precision != 16:
            raise ValueError('the device coder is specialised for 16-bit CDFs')
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer('_offset', torch.IntTensor())
        self.register_buffer('_quantized_cdf', torch.IntTensor())
        self.register_buffer('_cdf_length', torch.IntTensor())
        self._tables = None
        self._tables_key = None

    # ---- quantisation (torch ops: training-time / generic entry points) -----------------------
    def quantize(self, inputs, mode, means=None):
        if mode not in ('noise', 'dequantize', 'symbols'):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == 'noise':
            return inputs + torch.empty_like(inputs).uniform_(-0.5, 0.5)
        centred = inputs - means if means is not None else inputs
        rounded = torch.round(centred)
        if mode == 'dequantize':
            return rounded + means if means is not None else rounded
        return rounded.int()

    @staticmethod
    def dequantize(inputs, means=None, dtype=torch.float):
        if means is not None:
            return inputs.type_as(means) + means
        return inputs.type(dtype)

    # ---- tables -------------------------------------------------------------------------------
    def _pmf_to_cdf(self, pmf, tail_mass, pmf_length, max_length):
        rows = len(pmf_length)
        cdf = torch.zeros((rows, max_length + 2), dtype=torch.int32)
        pmf, tail_mass = pmf.detach().cpu(), tail_mass.detach().cpu()
        for r in range(rows):
            n = int(pmf_length[r])
            row = ops.pmf_to_quantized_cdf(torch.cat((pmf[r, :n], tail_mass[r].reshape(1))), self.entropy_coder_precision)
            cdf[r, :row.numel()] = row
        return cdf.to(pmf_length.device)

    def _check_tables(self):
        if self._quantized_cdf.numel() == 0:
            raise ValueError('Uninitialized CDFs. Run update() first')
        if self._quantized_cdf.dim() != 2:
            raise ValueError(f'Invalid CDF size {self._quantized_cdf.size()}')
        if self._cdf_length.numel() == 0:
            raise ValueError('Uninitialized CDF lengths. Run update() first')
        if self._cdf_length.dim() != 1:
            raise ValueError(f'Invalid offsets size {self._cdf_length.size()}')
        if self._offset.numel() == 0:
            raise ValueError('Uninitialized offsets. Run update() first')
        if self._offset.dim() != 1:
            raise ValueError(f'Invalid offsets size {self._offset.size()}')

    def coder_tables(self):
        """Device coder tables for the current buffers (rebuilt when the buffers change)."""
        self._check_tables()
        key = (self._quantized_cdf.data_ptr(), self._quantized_cdf._version, self._cdf_length._version,
               self._offset._version, tuple(self._quantized_cdf.shape))
        if self._tables is None or self._tables_key != key:
            self._tables = ops.CoderTables(self._quantized_cdf, self._cdf_length, self._offset)
            self._tables_key = key
        return self._tables

    # ---- batched device coder -----------------------------------------------------------------
    def compress(self, inputs, indexes, means=None):
        """Generic CompressAI entry point: explicit per-element `indexes` (and optional `means`)."""
        if inputs.dim() < 2:
            raise ValueError('Invalid `inputs` size. Expected a tensor with at least 2 dimensions.')
        if inputs.size() != indexes.size():
            raise ValueError('`inputs` and `indexes` should have the same size.')
        tables = self.coder_tables()
        ops.require_cuda(inputs, 'EntropyModel.compress')
        centred = inputs - means if means is not None else inputs
        symbols = ops.quantize_symbols(centred.reshape(inputs.size(0), 1, -1))
        return ops.rans_encode(symbols, tables, indexes=indexes.to(inputs.device)).tolist()

    def decompress(self, strings, indexes, dtype=torch.float, means=None):
        if not isinstance(strings, (tuple, list)):
            raise ValueError('Invalid `strings` parameter type.')
        if not len(strings) == indexes.size(0):
            raise ValueError('Invalid strings or indexes parameters')
        if indexes.dim() < 2:
            raise ValueError('Invalid `indexes` size. Expected a tensor with at least 2 dimensions.')
        tables = self.coder_tables()
        if means is not None:
            if means.size()[:2] != indexes.size()[:2]:
                raise ValueError('Invalid means or indexes parameters')
            if means.size() != indexes.size():
                for i in range(2, indexes.dim()):
                    if means.size(i) != 1:
                        raise ValueError('Invalid means parameters')
        ops.require_cuda(indexes, 'EntropyModel.decompress')
        streams = ops.PackedStreams.from_list(strings, indexes.device)
        n = indexes[0].numel()
        symbols = ops.rans_decode(streams, n, tables, indexes=indexes, want='symbols').view(indexes.size())
        return self.dequantize(symbols, means, dtype)


class EntropyBottleneck(EntropyModel):
    """Factorised-prior entropy model (Balle et al. 2018, appendix 6.1) -- CompressAI 1.2.x parametrisation."""

    def __init__(self, channels, *args, tail_mass=1e-9, init_scale=10, filters=(3, 3, 3, 3), **kwargs):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        widths = (1,) + self.filters + (1,)
        depth = len(self.filters) + 1
        scale = self.init_scale ** (1 / depth)
        self.matrices = nn.ParameterList()
        self.biases = nn.ParameterList()
        self.factors = nn.ParameterList()
        for i in range(depth):
            fan_out, fan_in = widths[i + 1], widths[i]
            self.matrices.append(nn.Parameter(torch.full((self.channels, fan_out, fan_in),
                                                         float(np.log(np.expm1(1 / scale / fan_out))))))
            self.biases.append(nn.Parameter(torch.empty(self.channels, fan_out, 1).uniform_(-0.5, 0.5)))
            if i < depth - 1:
                self.factors.append(nn.Parameter(torch.zeros(self.channels, fan_out, 1)))
        self.quantiles = nn.Parameter(torch.Tensor([-self.init_scale, 0, self.init_scale]).repeat(self.channels, 1, 1))
        t = math.log(2 / self.tail_mass - 1)
        self.register_buffer('target', torch.Tensor([-t, 0, t]))

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        # CompressAI <= 1.1 stored the density parameters as `_matrix0`, `_bias0`, `_factor0`, ...
        for old, new in (('_matrix', 'matrices.'), ('_bias', 'biases.'), ('_factor', 'factors.')):
            for i in range(len(self.filters) + 1):
                key = f'{prefix}{old}{i}'
                if key in state_dict:
                    state_dict[f'{prefix}{new}{i}'] = state_dict.pop(key)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    def _get_medians(self):
        return self.quantiles[:, :, 1:2]

    def _logits_cumulative(self, inputs, stop_gradient):
        logits = inputs
        last = len(self.filters)
        for i in range(last + 1):
            m, b = self.matrices[i], self.biases[i]
            if stop_gradient:
                m, b = m.detach(), b.detach()
            logits = torch.matmul(F.softplus(m), logits) + b
            if i < last:
                f = self.factors[i].detach() if stop_gradient else self.factors[i]
                logits = logits + torch.tanh(f) * torch.tanh(logits)
        return logits

    def _likelihood(self, inputs, stop_gradient=False):
        lower = self._logits_cumulative(inputs - 0.5, stop_gradient)
        upper = self._logits_cumulative(inputs + 0.5, stop_gradient)
        return torch.sigmoid(upper) - torch.sigmoid(lower), lower, upper

    def loss(self):
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        return torch.abs(logits - self.target).sum()

    @torch.no_grad()
    def update(self, force=False):
        """Builds `_quantized_cdf` / `_cdf_length` / `_offset` from the learned density.  Host fp32 arithmetic."""
        if self._offset.numel() > 0 and not force:
            return False
        device = self.quantiles.device
        host = {k: v.detach().cpu().float() for k, v in
                (('q', self.quantiles),) + tuple((f'm{i}', p) for i, p in enumerate(self.matrices)) +
                tuple((f'b{i}', p) for i, p in enumerate(self.biases)) + tuple((f'f{i}', p) for i, p in enumerate(self.factors))}
        q = host['q']
        medians = q[:, 0, 1]
        minima = torch.clamp(torch.ceil(medians - q[:, 0, 0]).int(), min=0)
        maxima = torch.clamp(torch.ceil(q[:, 0, 2] - medians).int(), min=0)
        pmf_start = medians - minima
        pmf_length = maxima + minima + 1
        max_length = int(pmf_length.max())
        samples = torch.arange(max_length)[None, :] + pmf_start[:, None, None]

        def logits_cumulative(v):
            for i in range(len(self.filters) + 1):
                v = torch.matmul(F.softplus(host[f'm{i}']), v) + host[f'b{i}']
                if i < len(self.filters):
                    v = v + torch.tanh(host[f'f{i}']) * torch.tanh(v)
            return v

        lower, upper = logits_cumulative(samples - 0.5), logits_cumulative(samples + 0.5)
        pmf = (torch.sigmoid(upper) - torch.sigmoid(lower))[:, 0, :]
        tail = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
        self._offset = (-minima).to(device)
        self._quantized_cdf = self._pmf_to_cdf(pmf, tail, pmf_length, max_length).to(device)
        self._cdf_length = (pmf_length + 2).to(device)
        self._tables = None
        return True

    def forward(self, x, training=None):
        """(B, C, ...) -> (y_hat, likelihoods); noise while training, rounding around the medians otherwise."""
        if training is None:
            training = self.training
        # channels first: the density is per channel
        xc = x.transpose(0, 1).contiguous()
        shape = xc.size()
        values = xc.reshape(shape[0], 1, -1)
        outputs = self.quantize(values, 'noise' if training else 'dequantize', self._get_medians())
        likelihood, _, _ = self._likelihood(outputs)
        if self.use_likelihood_bound:
            likelihood = self.likelihood_lower_bound(likelihood)
        outputs = outputs.reshape(shape).transpose(0, 1).contiguous()
        likelihood = likelihood.reshape(shape).transpose(0, 1).contiguous()
        return outputs, likelihood

    @staticmethod
    def _build_indexes(size):
        N, C = size[0], size[1]
        shape = [1] * len(size)
        shape[1] = C
        return torch.arange(C).view(*shape).int().repeat(N, 1, *size[2:])

    @staticmethod
    def _extend_ndims(tensor, n):
        return tensor.reshape(-1, *([1] * n)) if n > 0 else tensor.reshape(-1)

    # ---- hot path: channel-indexed batched coder ------------------------------------------------
    def compress_symbols(self, symbols, spatial):
        """int32 symbols [B, C, ...] already centred on the medians -> PackedStreams (device resident)."""
        return ops.rans_encode(symbols, self.coder_tables(), spatial=spatial, layout=getattr(self, 'coder_layout', None))

    def compress_packed(self, x):
        """Like `compress` but leaves the bitstreams on the device (PackedStreams)."""
        if x.dim() < 2:
            raise ValueError('Invalid `inputs` size. Expected a tensor with at least 2 dimensions.')
        tables = self.coder_tables()
        if x.size(1) != tables.n_rows:
            raise ValueError('`inputs` and `indexes` should have the same size.')
        ops.require_cuda(x, 'EntropyBottleneck.compress')
        medians = self._get_medians().detach().reshape(-1)
        symbols = ops.quantize_symbols(x, medians)
        spatial = x[0, 0].numel() if x.dim() > 2 else 1
        return ops.rans_encode(symbols, tables, spatial=spatial, layout=getattr(self, 'coder_layout', None))

    def compress(self, x):
        return self.compress_packed(x).tolist()

    def _fault_word(self, device):
        """One persistent int32 per device that every unchecked decode of this module ORs its fault flags into: batches in
        flight (several streams / host threads) share it, so no fault is overwritten by a later batch."""
        words = self.__dict__.setdefault('_fault_words', {})
        w = words.get(device)
        if w is None:
            with _FAULT_LOCK:
                w = words.get(device)
                if w is None:
                    with torch.inference_mode(False):
                        w = torch.zeros(1, dtype=torch.int32, device=device)
                    words[device] = w
        return w

    def decompress_packed(self, streams, size, want='values', check_status=False):
        """Device-resident decode.  With check_status=False (default) nothing synchronises: device fault flags
        (truncated / malformed stream) accumulate in a per-device fault word and are raised by `check_faults()`."""
        tables = self.coder_tables()
        C = self._quantized_cdf.size(0)
        spatial = int(np.prod(size)) if len(size) else 1
        medians = self._get_medians().detach().reshape(-1)
        status = None if check_status else self._fault_word(streams.packed.device)
        out = ops.rans_decode(streams, C * spatial, tables, spatial=spatial, means=medians, want=want, check_status=check_status,
                              layout=getattr(self, 'coder_layout', None), status=status)
        return out.view(streams.batch, C, *size)

    def check_faults(self):
        """Synchronises and raises if any device-resident decode since the last call flagged a truncated or malformed stream."""
        for dev, w in list(self.__dict__.get('_fault_words', {}).items()):
            torch.cuda.synchronize(dev)
            st = int(w.item())
            if st:
                w.zero_()
                ops.raise_on_decode_fault(st)

    def decompress(self, strings, size):
        if not isinstance(strings, (tuple, list)):
            raise ValueError('Invalid `strings` parameter type.')
        self._check_tables()
        device = self._quantized_cdf.device
        if device.type != 'cuda':
            raise RuntimeError('EntropyBottleneck.decompress: the sc2bench_b200 coder runs on CUDA only; move the model to a GPU')
        return self.decompress_packed(ops.PackedStreams.from_list(strings, device), tuple(size), check_status=True)


class GaussianConditional(EntropyModel):
    """Zero-mean (or `means`-shifted) Gaussian conditional with a quantised scale table (CompressAI contract)."""

    def __init__(self, scale_table, *args, scale_bound=0.11, tail_mass=1e-9, **kwargs):
        super().__init__(*args, **kwargs)
        if not isinstance(scale_table, (type(None), list, tuple)):
            raise ValueError(f'Invalid type for scale_table "{type(scale_table)}"')
        if isinstance(scale_table, (list, tuple)) and len(scale_table) < 1:
            raise ValueError(f'Invalid scale_table length "{len(scale_table)}"')
        if scale_table and (scale_table != sorted(scale_table) or any(s <= 0 for s in scale_table)):
            raise ValueError(f'Invalid scale_table "({scale_table})"')
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            scale_bound = scale_table[0]
        if scale_bound is None or scale_bound <= 0:
            raise ValueError('Invalid parameters')
        self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer('scale_table', self._prepare_scale_table(scale_table) if scale_table else torch.Tensor())
        self.register_buffer('scale_bound', torch.Tensor([float(scale_bound)]))

    @staticmethod
    def _prepare_scale_table(scale_table):
        return torch.Tensor(tuple(float(s) for s in scale_table))

    @staticmethod
    def _standardized_cumulative(inputs):
        return 0.5 * torch.erfc(float(-(2 ** -0.5)) * inputs)

    @staticmethod
    def _standardized_quantile(quantile):
        import scipy.stats
        return scipy.stats.norm.ppf(quantile)

    def update_scale_table(self, scale_table, force=False):
        if self._offset.numel() > 0 and not force:
            return False
        device = self.scale_table.device
        self.scale_table = self._prepare_scale_table(scale_table).to(device)
        self.update()
        return True

    @torch.no_grad()
    def update(self):
        device = self.scale_table.device
        table = self.scale_table.detach().cpu().float()
        multiplier = -self._standardized_quantile(self.tail_mass / 2)
        pmf_center = torch.ceil(table * multiplier).int()
        pmf_length = 2 * pmf_center + 1
        max_length = int(pmf_length.max())
        samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
        scale = table.unsqueeze(1)
        upper = self._standardized_cumulative((0.5 - samples) / scale)
        lower = self._standardized_cumulative((-0.5 - samples) / scale)
        self._quantized_cdf = self._pmf_to_cdf(upper - lower, 2 * lower[:, :1], pmf_length, max_length).to(device)
        self._offset = (-pmf_center).to(device)
        self._cdf_length = (pmf_length + 2).to(device)
        self._tables = None

    def _likelihood(self, inputs, scales, means=None):
        values = torch.abs(inputs - means if means is not None else inputs)
        scales = self.lower_bound_scale(scales)
        upper = self._standardized_cumulative((0.5 - values) / scales)
        lower = self._standardized_cumulative((-0.5 - values) / scales)
        return upper - lower

    def forward(self, inputs, scales, means=None, training=None):
        if training is None:
            training = self.training
        outputs = self.quantize(inputs, 'noise' if training else 'dequantize', means)
        likelihood = self._likelihood(outputs, scales, means)
        if self.use_likelihood_bound:
            likelihood = self.likelihood_lower_bound(likelihood)
        return outputs, likelihood

    def build_indexes(self, scales):
        """Index of the first table entry >= max(scale, bound) (one pass on the device instead of 63)."""
        # (scale_bound is a registered buffer: float() of it on the device is a blocking copy that drains the stream -- 10 ms per call in
        # the hyperprior's compress / decompress; its host value is cached until the buffer is written again)
        sb = self.scale_bound
        key = (sb.data_ptr(), sb._version, sb.device)
        cached = self.__dict__.get('_scale_bound_host')
        if cached is None or cached[0] != key:
            cached = (key, float(sb))
            self.__dict__['_scale_bound_host'] = cached
        return ops.gc_build_indexes(scales, self.scale_table, cached[1])
