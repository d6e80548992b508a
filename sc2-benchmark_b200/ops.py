"""Tensor-level front-end of the C ABI (include/sc2b200.h).

PyTorch is used here for device memory (caching allocator), streams and pinned host buffers only; every
computation below is a call into libsc2b200.so on the current CUDA stream.  There is no CPU fallback:
a non-CUDA tensor on the hot path raises.
"""
import ctypes
import os
import threading

import numpy as np
import torch

from . import _native
from ._native import ConvDesc, GaHaloDesc, TcConvDesc, TcConvExDesc, TcSplitDesc, TcSplitExDesc, check


def _lib():
    return _native.load()


# ---- launch accounting (bench.py reads these; cheap enough to keep on) ----------------------------
STATS = {'launches': 0}
_PROFILE = {'tags': None, 'events': []}


def profile_kernels(tags):
    """Bracket launches whose tag is in `tags` (or every launch for tags == 'all') with CUDA events on the launching
    stream.  `profile_kernels(None)` turns it off.  Read the result with `profile_results()`."""
    _PROFILE['tags'] = tags
    _PROFILE['events'] = []


def profile_results():
    """{tag: [ms, ...]} -- call after a synchronize."""
    out = {}
    for tag, e0, e1, _ in _PROFILE['events']:
        out.setdefault(tag, []).append(e0.elapsed_time(e1))
    return out


def profile_work():
    """{tag: (flops, bytes, symbols) of ONE launch} for the launches profiled so far (None where an op does not state it)."""
    return {tag: work for tag, _, _, work in _PROFILE['events']}


class _launch:
    """Counts a launch and, when profiling is on, brackets it with CUDA events.  flops / nbytes: the ALGORITHMIC work of the
    launch (2 x MAC of the reference's layers; bytes = every input read once and every output written once at the reference's
    4 bytes per activation) -- what bench.py's roofline fractions are computed from."""

    def __init__(self, tag, kernels=1, flops=None, nbytes=None, symbols=None):
        self.tag, self.kernels, self.e0 = tag, kernels, None
        self.work = (flops, nbytes, symbols)

    def __enter__(self):
        STATS['launches'] += self.kernels
        tags = _PROFILE['tags']
        if tags is not None and (tags == 'all' or self.tag in tags):
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _PROFILE['events'].append((self.tag, self.e0, e1, self.work))
        return False


def _stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _TileCounters:
    """Zeroed int32 slots for the dynamic tile schedule of the persistent tensor-core kernels (include/sc2b200.h,
    `tile_counter`): one slot per launch, handed out from a per-stream block that is re-zeroed (stream-ordered, so after
    every kernel that used it) when it runs out.  SC2_TC_STATIC=1 selects the static schedule (NULL counters)."""
    SLOTS = 4096

    def __init__(self):
        self._blocks = {}
        self._lock = threading.Lock()
        self.static = bool(int(os.environ.get('SC2_TC_STATIC', '0')))

    def next(self):
        if self.static:
            return ctypes.c_void_p(0)
        stream = torch.cuda.current_stream()
        key = (stream.device.index, stream.cuda_stream)
        with self._lock:
            blk = self._blocks.get(key)
            if blk is None:
                with torch.inference_mode(False):
                    blk = [torch.zeros(self.SLOTS, dtype=torch.int32, device=stream.device), 0]
                self._blocks[key] = blk
            if blk[1] == self.SLOTS:
                blk[0].zero_()
                blk[1] = 0
            i = blk[1]
            blk[1] += 1
        return ctypes.c_void_p(blk[0].data_ptr() + 4 * i)


_TILE_COUNTERS = _TileCounters()


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def require_cuda(t, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError('%s: the sc2bench_b200 hot path runs on CUDA tensors only (got %s); there is no CPU fallback'
                           % (what, t.device if isinstance(t, torch.Tensor) else type(t)))


# ----------------------------------------------------------------------------------------------
# host: CDF quantisation (compressai._CXX.pmf_to_quantized_cdf)
# ----------------------------------------------------------------------------------------------
def pmf_to_quantized_cdf(pmf, precision=16):
    """float pmf (sequence / 1-D tensor) -> torch.IntTensor CDF of len(pmf) + 1 entries."""
    arr = np.ascontiguousarray(np.asarray(pmf.detach().cpu() if isinstance(pmf, torch.Tensor) else pmf, dtype=np.float32)).reshape(-1)
    out = np.empty(arr.size + 1, dtype=np.uint32)
    rc = _lib().sc2_pmf_to_quantized_cdf(arr.ctypes.data, arr.size, int(precision), out.ctypes.data)
    if rc == -4:
        raise ValueError('Invalid `pmf`: negative, non-finite or all-zero.')
    check(rc, 'sc2_pmf_to_quantized_cdf')
    return torch.from_numpy(out.astype(np.int32))


# ----------------------------------------------------------------------------------------------
# coder tables
# ----------------------------------------------------------------------------------------------
class CoderTables:
    """Device-side coder tables built from CompressAI's (_quantized_cdf, _cdf_length, _offset) buffers."""

    def __init__(self, quantized_cdf, cdf_length, offset):
        cdf = np.ascontiguousarray(quantized_cdf.detach().cpu().numpy().astype(np.int32))
        sizes = np.ascontiguousarray(cdf_length.detach().cpu().numpy().astype(np.int32).reshape(-1))
        offs = np.ascontiguousarray(offset.detach().cpu().numpy().astype(np.int32).reshape(-1))
        if cdf.ndim != 2:
            raise ValueError(f'Invalid CDF size {tuple(cdf.shape)}')
        if sizes.shape[0] != cdf.shape[0] or offs.shape[0] != cdf.shape[0]:
            raise ValueError('Invalid CDF lengths / offsets size')
        self.n_rows, self.cdf_stride = int(cdf.shape[0]), int(cdf.shape[1])
        self.max_size = int(sizes.max())
        nbytes = _lib().sc2_rans_table_bytes(self.n_rows, self.cdf_stride)
        blob = np.zeros(nbytes, dtype=np.uint8)
        rc = _lib().sc2_rans_build_tables(cdf.ctypes.data, sizes.ctypes.data, offs.ctypes.data, self.n_rows,
                                          self.cdf_stride, blob.ctypes.data)
        if rc == -1:
            raise ValueError('Invalid CDF table (each row must start at 0, end at 65536 and be strictly increasing)')
        check(rc, 'sc2_rans_build_tables')
        self._blob = torch.from_numpy(blob)
        self._on_device = {}

    def on(self, device):
        device = torch.device(device)
        t = self._on_device.get(device)
        if t is None:
            t = self._blob.to(device)
            self._on_device[device] = t
        return t


# ----------------------------------------------------------------------------------------------
# packed bitstreams
# ----------------------------------------------------------------------------------------------
class _PinnedPool(threading.local):
    """Per-thread, grow-only pinned staging buffers (cudaHostAlloc per call costs milliseconds)."""

    def __init__(self):
        self.d2h = None
        self.h2d = None
        self.h2d_event = None

    def get_d2h(self, nbytes):
        if self.d2h is None or self.d2h.numel() < nbytes:
            with torch.inference_mode(False):  # a plain tensor: it outlives the caller's inference_mode block
                self.d2h = torch.empty(max(nbytes, 1 << 20) * 5 // 4, dtype=torch.uint8, pin_memory=True)
        return self.d2h[:nbytes]

    def get_h2d(self, nbytes):
        if self.h2d_event is not None:
            self.h2d_event.synchronize()  # the previous upload from this buffer has finished
        if self.h2d is None or self.h2d.numel() < nbytes:
            with torch.inference_mode(False):
                self.h2d = torch.empty(max(nbytes, 1 << 20) * 5 // 4, dtype=torch.uint8, pin_memory=True)
        return self.h2d[:nbytes]


_PINNED = _PinnedPool()


class PackedStreams:
    """B CompressAI bitstreams held back to back in one device buffer (+ int64 offsets[B + 1]).

    `tolist()` materialises the reference contract's `list[bytes]` with ONE D2H copy."""

    def __init__(self, packed, offsets, batch, status=None, capacity=None):
        self.packed, self.offsets, self.batch, self.status = packed, offsets, batch, status
        self._offs = None
        # The streams are complete once the PRODUCING stream reaches this point (the coder / the upload may run on a batch
        # stream of CodecPipeline while the consumer reads them from another stream): host-side readers wait on this event.
        self.ready = torch.cuda.Event()
        self.ready.record(torch.cuda.current_stream(packed.device))

    def _host_offsets(self):
        if self._offs is None:
            # Wait with an event synchronise FIRST: a blocking copy to pageable memory (.cpu(), .item()) holds a driver lock
            # while it waits for the GPU, and every other host thread's event record / stream wait then blocks behind it for
            # milliseconds (measured: scripts/diag_e2e_calls.py).
            self.ready.synchronize()
            offs = self.offsets.cpu()
            if self.status is not None:
                st = int(self.status.item())
                if st & _native.FAULT_ARENA_OVERFLOW:
                    raise RuntimeError('rANS encoder ran out of arena space (device fault flag)')
                if st & _native.FAULT_BAD_INDEX:
                    raise ValueError('Invalid `indexes`: a CDF index is outside [0, number of CDF rows)')
            self._offs = np.ascontiguousarray(offs.numpy(), dtype=np.int64)
        return self._offs

    def tolist(self):
        offs = self._host_offsets()
        total = int(offs[-1])
        host = _PINNED.get_d2h(total)  # per-thread staging buffer, reused by the next call: nothing of it is cached here
        host.copy_(self.packed[:total], non_blocking=False)  # (the producing stream has finished: _host_offsets waited)
        # one copy per stream, straight out of the pinned staging buffer, with the GIL released
        return _native.hostbytes().split(host.numpy(), offs)

    def lengths(self):
        return np.diff(self._host_offsets())

    def total_bytes(self):
        return int(self._host_offsets()[-1])

    @staticmethod
    def from_list(strings, device):
        if not isinstance(strings, (tuple, list)):
            raise ValueError('Invalid `strings` parameter type.')
        if not all(isinstance(b, bytes) for b in strings):
            strings = [bytes(b) for b in strings]
        total = sum(map(len, strings))
        # one pinned staging buffer per thread: [streams ... | pad to 8 | int64 offsets], uploaded with ONE H2D copy;
        # the gather runs with the GIL released (csrc/hostbytes.c)
        off_pos = (max(total, 4) + 7) // 8 * 8
        n_off = 8 * (len(strings) + 1)
        host = _PINNED.get_h2d(off_pos + n_off)
        hv = host.numpy()
        _native.hostbytes().join(strings, hv[:off_pos], hv[off_pos:off_pos + n_off])
        offs = hv[off_pos:off_pos + n_off].view(np.int64)
        lens = np.diff(offs)
        if len(strings) and ((lens < 8).any() or (lens & 3).any()):
            raise ValueError('Invalid bitstream: a CompressAI rANS stream is a multiple of 4 bytes and >= 8 bytes long')
        blob = host.to(device, non_blocking=True)
        _PINNED.h2d_event = torch.cuda.Event()
        _PINNED.h2d_event.record()
        packed = blob[:max(total, 4)]
        offsets = blob[off_pos:off_pos + offs.nbytes].view(torch.int64)
        return PackedStreams(packed, offsets, len(strings))


# ----------------------------------------------------------------------------------------------
# device ops
# ----------------------------------------------------------------------------------------------
def quantize_symbols(x, means=None):
    """EntropyModel.quantize(x, "symbols", means) for x [B, C, *spatial]; means: [C] tensor, a tensor of x's shape
    (one mean per element, mean-scale hyperprior) or None."""
    require_cuda(x, 'quantize_symbols')
    x = x.contiguous().float()
    B, C = x.shape[0], x.shape[1]
    spatial = x[0, 0].numel() if x.dim() > 2 else 1
    if means is not None and means.numel() != C:
        if means.shape != x.shape:
            raise ValueError('means must have one value per channel or per element')
        B, C, spatial = 1, x.numel(), 1  # flat: element i takes means[i]
        if C >= 2 ** 31:
            raise ValueError('tensor too large for per-element means')
    out = torch.empty(x.shape, dtype=torch.int32, device=x.device)
    m = means.contiguous().float() if means is not None else None
    with torch.cuda.device(x.device), _launch('quantize_symbols', nbytes=8.0 * x.numel()):
        check(_lib().sc2_quantize_symbols(_ptr(x), _ptr(m), _ptr(out), B, C, spatial, _stream_ptr()), 'sc2_quantize_symbols')
    return out


def rans_encode(symbols, tables, indexes=None, spatial=None, slot_bytes=None, layout=None):
    """symbols: int32 [B, n] (or [B, C, ...]) CUDA tensor -> PackedStreams.  layout (channel mode): None / 'warp' (a warp per
    stream, lowest latency) or 'lanes' (32 streams per warp, for batches in flight; include/sc2b200.h)."""
    require_cuda(symbols, 'rans_encode')
    dev = symbols.device
    B = symbols.shape[0]
    n = symbols[0].numel() if B else (symbols.numel() if symbols.dim() < 2 else int(np.prod(symbols.shape[1:])))
    sym = symbols.contiguous().view(B, n)
    if sym.dtype != torch.int32:
        sym = sym.int()
    idx = None
    if indexes is not None:
        idx = indexes.contiguous().view(B, n)
        if idx.dtype != torch.int32:
            idx = idx.int()
        if idx.shape != sym.shape:
            raise ValueError('`inputs` and `indexes` should have the same size.')
    elif spatial is None:
        raise ValueError('channel mode needs `spatial` (symbols per channel)')
    lib = _lib()
    if slot_bytes is None:
        slot_bytes = int(lib.sc2_rans_max_stream_bytes(n))
    arena = torch.empty(max(B, 1) * slot_bytes, dtype=torch.uint8, device=dev)
    lengths = torch.empty(max(B, 1), dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    packed = torch.empty(max(B, 1) * slot_bytes, dtype=torch.uint8, device=dev)
    offsets = torch.empty(B + 1, dtype=torch.int64, device=dev)
    tab = tables.on(dev)
    with torch.cuda.device(dev):
        st = _stream_ptr()
        with _launch('rans_encode', 1 if B else 0, nbytes=4.0 * B * n, symbols=(B, n)):
            check(lib.sc2_rans_encode_batch(_ptr(sym), _ptr(idx), B, n, int(spatial or 0), _ptr(tab), tables.n_rows,
                                            tables.cdf_stride, _ptr(arena), slot_bytes, _ptr(lengths), _ptr(status),
                                            _native.rans_layout(layout, B), st),
                  'sc2_rans_encode_batch')
        with _launch('rans_pack', 2 if B else 1):
            check(lib.sc2_rans_pack(_ptr(arena), slot_bytes, _ptr(lengths), B, _ptr(packed), _ptr(offsets), st), 'sc2_rans_pack')
    return PackedStreams(packed, offsets, B, status)


def raise_on_decode_fault(st):
    if st:
        if st & _native.FAULT_BAD_INDEX:
            raise ValueError('Invalid `indexes`: a CDF index is outside [0, number of CDF rows) (device fault flags 0x%x)' % st)
        raise ValueError('Invalid bitstream (device fault flags 0x%x: %s)' % (st, ', '.join(
            name for bit, name in ((2, 'stream truncated'), (4, 'bad stream length')) if st & bit)))


def rans_decode(streams, n_per_stream, tables, indexes=None, spatial=None, means=None, want='values', check_status=True,
                return_status=False, layout=None, status=None):
    """PackedStreams -> [B, n] float32 (symbol + means[row]) or int32 symbols.  `status`: a caller-owned int32 fault word the
    kernel ORs its flags into (several decodes may share one); default: a fresh zeroed word."""
    dev = streams.packed.device
    B = streams.batch
    idx = None
    if indexes is not None:
        require_cuda(indexes, 'rans_decode')
        idx = indexes.contiguous().view(B, n_per_stream)
        if idx.dtype != torch.int32:
            idx = idx.int()
    elif spatial is None:
        raise ValueError('channel mode needs `spatial` (symbols per channel)')
    out_sym = torch.empty((B, n_per_stream), dtype=torch.int32, device=dev) if want == 'symbols' else None
    out_val = torch.empty((B, n_per_stream), dtype=torch.float32, device=dev) if want == 'values' else None
    m = means.contiguous().float() if means is not None else None
    if status is None:
        status = torch.zeros(1, dtype=torch.int32, device=dev)
    tab = tables.on(dev)
    with torch.cuda.device(dev), _launch('rans_decode', 1 if B and n_per_stream else 0, nbytes=4.0 * B * n_per_stream,
                                         symbols=(B, n_per_stream)):
        check(_lib().sc2_rans_decode_batch(_ptr(streams.packed), _ptr(streams.offsets), B, n_per_stream, _ptr(idx),
                                           int(spatial or 0), _ptr(tab), tables.n_rows, tables.cdf_stride, _ptr(out_sym),
                                           _ptr(out_val), _ptr(m), _ptr(status), _native.rans_layout(layout, B), _stream_ptr()),
              'sc2_rans_decode_batch')
    if check_status:
        torch.cuda.current_stream(dev).synchronize()  # (not a blocking copy: see PackedStreams._host_offsets)
        raise_on_decode_fault(int(status.item()))
    out = out_sym if want == 'symbols' else out_val
    return (out, status) if return_status else out


def dequantize(symbols, means=None):
    """EntropyModel.dequantize(symbols, means) with one mean per element: float(symbols) + means."""
    require_cuda(symbols, 'dequantize')
    sym = symbols.contiguous()
    if sym.dtype != torch.int32:
        sym = sym.int()
    m = None
    if means is not None:
        if means.shape != sym.shape:
            raise ValueError('means must have the shape of symbols')
        m = means.contiguous().float()
    out = torch.empty(sym.shape, dtype=torch.float32, device=sym.device)
    with torch.cuda.device(sym.device), _launch('dequantize', 1 if sym.numel() else 0, nbytes=8.0 * sym.numel()):
        check(_lib().sc2_dequantize(_ptr(sym), _ptr(m), _ptr(out), sym.numel(), _stream_ptr()), 'sc2_dequantize')
    return out


def gc_build_indexes(scales, scale_table, scale_bound):
    require_cuda(scales, 'gc_build_indexes')
    s = scales.contiguous().float()
    out = torch.empty(s.shape, dtype=torch.int32, device=s.device)
    tab = scale_table.to(s.device).contiguous().float()
    with torch.cuda.device(s.device), _launch('gc_build_indexes', nbytes=8.0 * s.numel()):
        check(_lib().sc2_gc_build_indexes(_ptr(s), s.numel(), _ptr(tab), tab.numel(), float(scale_bound), _ptr(out),
                                          _stream_ptr()), 'sc2_gc_build_indexes')
    return out


def conv2d(x, weight, bias=None, stride=1, padding=0, transposed=False, output_padding=0, epilogue=_native.EPI_NONE, aux=None,
           in_abs=False, epi_param=0.0):
    """fp32 NCHW conv / transposed conv with a fused epilogue (sc2_conv2d_f32); in_abs convolves |x|."""
    require_cuda(x, 'conv2d')
    x = x.contiguous().float()
    w = weight.detach().contiguous().float()
    B, Cin, H, W = x.shape
    if transposed:
        if w.shape[0] != Cin:
            raise ValueError('weight / input channel mismatch')
        Cout = w.shape[1]
    else:
        if w.shape[1] != Cin:
            raise ValueError('weight / input channel mismatch (groups are not supported)')
        Cout = w.shape[0]
    d = ConvDesc(B, Cin, H, W, Cout, w.shape[2], w.shape[3], int(stride), int(padding), int(bool(transposed)),
                 int(output_padding), int(epilogue), _native.IN_ABS if in_abs else _native.IN_NONE, float(epi_param))
    ho, wo = ctypes.c_int(), ctypes.c_int()
    check(_lib().sc2_conv_out_size(ctypes.byref(d), ctypes.byref(ho), ctypes.byref(wo)), 'sc2_conv_out_size')
    out = torch.empty((B, Cout, ho.value, wo.value), dtype=torch.int32 if epilogue == _native.EPI_QUANTIZE else torch.float32,
                      device=x.device)
    b = bias.detach().contiguous().float() if bias is not None else None
    a = aux.detach().contiguous().float() if aux is not None else None
    tag = 'conv2d_f32[%d->%d,k%d,s%d%s]' % (Cin, Cout, w.shape[2], int(stride), ',T' if transposed else '')
    macs = (B * Cin * H * W * Cout if transposed else B * Cout * ho.value * wo.value * Cin) * w.shape[2] * w.shape[3]
    with torch.cuda.device(x.device), _launch(tag, flops=2.0 * macs, nbytes=4.0 * (x.numel() + out.numel())):
        check(_lib().sc2_conv2d_f32(ctypes.byref(d), _ptr(x), _ptr(w), _ptr(b), _ptr(a), _ptr(out), _stream_ptr()), 'sc2_conv2d_f32')
    return out


def gdn(x, gamma, beta, kind=0, inverse=False):
    """GDN1 (kind 0) / GDN (kind 1) on NCHW fp32 with EFFECTIVE gamma [C, C] and beta [C]."""
    require_cuda(x, 'gdn')
    x = x.contiguous().float()
    B, C = x.shape[0], x.shape[1]
    spatial = x[0, 0].numel()
    g = gamma.detach().reshape(C, C).contiguous().float()
    b = beta.detach().contiguous().float()
    out = torch.empty_like(x)
    with torch.cuda.device(x.device), _launch('gdn_f32[%d%s]' % (C, ',inv' if inverse else ''), flops=2.0 * B * C * C * spatial,
                                              nbytes=8.0 * x.numel()):
        check(_lib().sc2_gdn_f32(_ptr(x), _ptr(g), _ptr(b), _ptr(out), B, C, spatial, int(kind), int(bool(inverse)),
                                 _stream_ptr()), 'sc2_gdn_f32')
    return out


# ----------------------------------------------------------------------------------------------
# tensor-core path (NHWC fp16)
# ----------------------------------------------------------------------------------------------
def pack_conv_weight_f16(weight, c_in_pad=None):
    """Conv2d weight [c_out, c_in, kh, kw] -> [kh*kw, c_out, c_in_pad] fp16 (tap-major, K contiguous). Once per model."""
    w = weight.detach().float()
    c_out, c_in, kh, kw = w.shape
    if c_in_pad is None:
        c_in_pad = (c_in + 63) // 64 * 64
    packed = torch.zeros((kh * kw, c_out, c_in_pad), dtype=torch.float16, device=w.device)
    packed[:, :, :c_in] = w.permute(2, 3, 0, 1).reshape(kh * kw, c_out, c_in).half()
    return packed.contiguous()


def nchw_to_nhwc_f16(x, c_pad):
    """fp32 NCHW -> fp16 NHWC with zero-padded channels: returns a [B, H, W, c_pad] tensor."""
    require_cuda(x, 'nchw_to_nhwc_f16')
    x = x.contiguous().float()
    B, C, H, W = x.shape
    y = torch.empty((B, H, W, c_pad), dtype=torch.float16, device=x.device)
    with torch.cuda.device(x.device), _launch('nchw_to_nhwc_f16', nbytes=4.0 * x.numel() + 2.0 * y.numel()):
        check(_lib().sc2_nchw_f32_to_nhwc_f16(_ptr(x), _ptr(y), B, C, H * W, c_pad, _stream_ptr()), 'sc2_nchw_f32_to_nhwc_f16')
    return y


def tc_conv(x_nhwc, w_packed, kh, kw, pad, mode=_native.TC_STORE_F16, beta=None, gdn_x=None, c_in=None, signs=None):
    """tcgen05 implicit-GEMM conv on NHWC fp16 (sc2_tc_conv_nhwc). Returns NHWC fp16 / fp32.  c_in: the layer's real input
    channels (the activation may be zero-padded to a multiple of 64), for the work accounting only."""
    require_cuda(x_nhwc, 'tc_conv')
    assert x_nhwc.dtype == torch.float16 and x_nhwc.is_contiguous() and w_packed.dtype == torch.float16
    B, H, W, Cp = x_nhwc.shape
    taps, c_out, cp2 = w_packed.shape
    if taps != kh * kw or cp2 != Cp:
        raise ValueError('packed weight does not match the activation / kernel size')
    d = TcConvDesc(B, H, W, Cp, c_out, kh, kw, pad, mode)
    ho, wo = H + 2 * pad - kh + 1, W + 2 * pad - kw + 1
    out = torch.empty((B, ho, wo, c_out), dtype=torch.float32 if mode == _native.TC_STORE_F32 else torch.float16, device=x_nhwc.device)
    b = beta.detach().contiguous().float() if beta is not None else None
    if mode == _native.TC_STORE_ABS_F16:  # |x| + packed signs out (the pair an IGDN1 in mode TC_IGDN1_ABS_F16 reads)
        signs = torch.empty((B, ho, wo, c_out // 32), dtype=torch.int32, device=x_nhwc.device)
    tag = 'tc_conv[%d->%d,k%d,m%d]' % (Cp, c_out, kh, mode)
    c_real = c_in or Cp
    flops = 2.0 * B * ho * wo * c_out * c_real * kh * kw
    nbytes = 4.0 * B * (H * W * c_real + ho * wo * c_out)
    with torch.cuda.device(x_nhwc.device), _launch(tag, flops=flops, nbytes=nbytes):
        check(_lib().sc2_tc_conv_nhwc(ctypes.byref(d), _ptr(x_nhwc), _ptr(w_packed), _ptr(b), _ptr(gdn_x), _ptr(out), _ptr(signs),
                                      _TILE_COUNTERS.next(), _stream_ptr()), 'sc2_tc_conv_nhwc')
    return (out, signs) if mode == _native.TC_STORE_ABS_F16 else out


# ----------------------------------------------------------------------------------------------
# fp32-grade tensor-core path ("split fp16": value = hi + lo / 2048)
# ----------------------------------------------------------------------------------------------
LO_SCALE = 2048.0


def split_f16(t):
    """fp32 tensor -> (hi, lo) fp16 pair with t ~= hi + lo / 2048 (22 mantissa bits). Host-side prep of weights."""
    t = t.detach().float()
    hi = t.half()
    lo = ((t - hi.float()) * LO_SCALE).half()
    return hi.contiguous(), lo.contiguous()


def pack_conv_weight_split(weight, c_in_pad=None, as_patches=False):
    """Conv2d weight [c_out, c_in, kh, kw] -> (hi, lo) of shape [taps, n_tile, c_in_pad], rows beyond c_out zero.
    as_patches=True packs the whole receptive field into ONE tap with K = (c, dy, dx) (first layer, after im2col)."""
    w = weight.detach().float()
    c_out, c_in, kh, kw = w.shape
    n_tile = _lib().sc2_tc_split_n_tile(c_out)
    if n_tile == 0:
        raise ValueError('c_out %d is not supported by the split tensor-core kernel' % c_out)
    if as_patches:
        k = c_in * kh * kw
        c_in_pad = c_in_pad or (k + 15) // 16 * 16
        packed = torch.zeros((1, n_tile, c_in_pad), dtype=torch.float32, device=w.device)
        packed[0, :c_out, :k] = w.reshape(c_out, k)
    else:
        c_in_pad = c_in_pad or (c_in + 15) // 16 * 16
        packed = torch.zeros((kh * kw, n_tile, c_in_pad), dtype=torch.float32, device=w.device)
        packed[:, :c_out, :c_in] = w.permute(2, 3, 0, 1).reshape(kh * kw, c_out, c_in)
    return split_f16(packed)


def patchify_split(x, kh, kw, stride, pad, k_pad):
    """fp32 NCHW image -> split fp16 patches [B * 4, h_out/2, w_out/2, k_pad] in parity-plane pixel order."""
    require_cuda(x, 'patchify_split')
    x = x.contiguous().float()
    B, C, H, W = x.shape
    ho, wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    hi = torch.empty((B * 4, ho // 2, wo // 2, k_pad), dtype=torch.float16, device=x.device)
    lo = torch.empty_like(hi)
    with torch.cuda.device(x.device), _launch('patchify_split', nbytes=4.0 * (x.numel() + hi.numel())):
        check(_lib().sc2_patchify_split(_ptr(x), _ptr(hi), _ptr(lo), B, C, H, W, kh, kw, stride, pad, k_pad, _stream_ptr()),
              'sc2_patchify_split')
    return hi, lo


def tc_split_conv(x_hi, x_lo, w_hi, w_lo, c_out, kh, kw, stride, pad, mode, beta=None, medians=None, gdn=False):
    """sc2_tc_split_conv.  x planes: stride 1 [images, H, W, C]; stride 2 parity planes [images * 4, H, W, C].
    Returns (hi, lo) planes [images, h_out, w_out, c_out] or int32 symbols [images, c_out, h_out, w_out] (mode QUANT)."""
    require_cuda(x_hi, 'tc_split_conv')
    n_img_planes, H, W, C = x_hi.shape
    planes = 4 if stride == 2 else 1
    images = n_img_planes // planes
    if stride == 2:
        ho, wo = (2 * H + 2 * pad - kh) // 2 + 1, (2 * W + 2 * pad - kw) // 2 + 1
    else:
        ho, wo = H + 2 * pad - kh + 1, W + 2 * pad - kw + 1
    dev = x_hi.device
    out_hi = out_lo = out_sym = None
    if mode == _native.TCS_QUANT:
        out_sym = torch.empty((images, c_out, ho, wo), dtype=torch.int32, device=dev)
        out_c = c_out
    else:
        out_c = (c_out + 7) // 8 * 8
        out_hi = torch.empty((images, ho, wo, out_c), dtype=torch.float16, device=dev)
        out_lo = torch.empty_like(out_hi)
        if out_c != c_out:
            out_hi.zero_()
            out_lo.zero_()
    d = TcSplitDesc(images, H, W, C, c_out, kh, kw, stride, pad, mode, ho, wo, out_c)
    b = beta.detach().contiguous().float() if beta is not None else None
    m = medians.detach().contiguous().float() if medians is not None else None
    tag = 'tc_split[%d->%d,k%d,s%d,m%d]' % (C, c_out, kh, stride, mode)
    flops = 2.0 * images * ho * wo * c_out * C * kh * kw
    nbytes = 4.0 * (x_hi.numel() + images * ho * wo * c_out)
    with torch.cuda.device(dev), _launch(tag, flops=flops, nbytes=nbytes):
        check(_lib().sc2_tc_split_conv(ctypes.byref(d), _ptr(x_hi), _ptr(x_lo), _ptr(w_hi), _ptr(w_lo), _ptr(b), _ptr(m),
                                       _ptr(x_hi) if gdn else None, _ptr(x_lo) if gdn else None, _ptr(out_hi), _ptr(out_lo),
                                       _ptr(out_sym), _TILE_COUNTERS.next(), _stream_ptr()), 'sc2_tc_split_conv')
    return out_sym if mode == _native.TCS_QUANT else (out_hi, out_lo)


def patchify_split_nhwc(x, kh, kw, stride, pad, k_pad):
    """fp32 NCHW image -> split fp16 patches [B, h_out, w_out, k_pad], pixels in plain NHWC order (sc2_patchify_split_nhwc)."""
    require_cuda(x, 'patchify_split_nhwc')
    x = x.contiguous().float()
    B, C, H, W = x.shape
    ho, wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    hi = torch.empty((B, ho, wo, k_pad), dtype=torch.float16, device=x.device)
    lo = torch.empty_like(hi)
    with torch.cuda.device(x.device), _launch('patchify_split', nbytes=4.0 * (x.numel() + hi.numel())):
        check(_lib().sc2_patchify_split_nhwc(_ptr(x), _ptr(hi), _ptr(lo), B, C, H, W, kh, kw, stride, pad, k_pad, _stream_ptr()),
              'sc2_patchify_split_nhwc')
    return hi, lo


def split_n_tiles(c_out):
    """Output-channel tiles [(first channel, channels)] of a layer wider than one tensor-core tile (sc2_tc_split_conv_ex): equal
    tiles of the largest size <= 128 that divides c_out into multiples of 16, else 128s and a remainder."""
    if c_out <= 128:
        return [(0, c_out)]
    for n in (128, 96, 64, 48, 32):
        if c_out % n == 0 and c_out // n <= 3:
            return [(i * n, n) for i in range(c_out // n)]
    tiles, c0 = [], 0
    while c0 < c_out:
        n = min(128, c_out - c0)
        tiles.append((c0, n))
        c0 += n
    return tiles


def pack_conv_weight_split_tiles(weight, c_in_pad=None, as_patches=False):
    """pack_conv_weight_split per output-channel tile: [(first channel, channels, w_hi, w_lo)]."""
    return [(c0, n) + pack_conv_weight_split(weight[c0:c0 + n], c_in_pad=c_in_pad, as_patches=as_patches)
            for c0, n in split_n_tiles(weight.shape[0])]


def tc_split_conv_tiled(x_hi, x_lo, tiles, kh, kw, stride, pad, mode, vec=None, medians=None, gdn_x=None, act=_native.TCS_ACT_NONE,
                        slope=0.0, in_nhwc=False, name='tc_split', pad_x=None, out=None, out_parity=None, out_hw=None):
    """sc2_tc_split_conv_ex over the output-channel tiles of pack_conv_weight_split_tiles (one launch per tile, all writing the same
    planes).  x planes: [images, H, W, C] (stride 1, or stride 2 with in_nhwc) or parity planes [images * 4, H/2, W/2, C].
    vec: bias (STORE / QUANT) or GDN beta, the full vector; gdn_x: (hi, lo) planes of x for the GDN modes.
    out_parity=(py, px) with out=(hi, lo) planes [images, Ho, Wo, ceil8(c_out)]: the results are the pixels (2Y + py, 2X + px) of
    those planes (one sub-convolution of a ConvTranspose2d(k5, s2); kh, kw, pad, pad_x from deconv5_parity_taps).
    out_hw: the output size when it is not the one the (possibly zero-padded) input planes imply.
    Returns (hi, lo) planes [images, h_out, w_out, ceil8(c_out)], or int32 symbols [images, c_out, h_out, w_out] (TCS_QUANT)."""
    require_cuda(x_hi, 'tc_split_conv_tiled')
    n_img_planes, H, W, C = x_hi.shape
    planes = 4 if (stride == 2 and not in_nhwc) else 1
    images = n_img_planes // planes
    if stride == 2 and not in_nhwc:
        ho, wo = (2 * H + 2 * pad - kh) // 2 + 1, (2 * W + 2 * pad - kw) // 2 + 1
    else:
        ho, wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    c_out = tiles[-1][0] + tiles[-1][1]
    dev = x_hi.device
    out_hi = out_lo = out_sym = None
    pitch = (c_out + 7) // 8 * 8
    if out_hw is not None:
        ho, wo = out_hw
    full_h = full_w = 0
    if out_parity is not None:
        out_hi, out_lo = out
        full_h, full_w = out_hi.shape[1], out_hi.shape[2]
        ho, wo = (full_h - out_parity[0] + 1) // 2, (full_w - out_parity[1] + 1) // 2
        if out_hi.shape[0] != images or out_hi.shape[3] != pitch or stride != 1 or mode != _native.TCS_STORE:
            raise ValueError('out planes %s do not fit %d channels of %d images' % (tuple(out_hi.shape), c_out, images))
    elif mode == _native.TCS_QUANT:
        out_sym = torch.empty((images, c_out, ho, wo), dtype=torch.int32, device=dev)
    else:
        out_hi = torch.empty((images, ho, wo, pitch), dtype=torch.float16, device=dev)
        out_lo = torch.empty_like(out_hi)
    v = vec.detach().contiguous().float() if vec is not None else None
    m = medians.detach().contiguous().float() if medians is not None else None
    gx_hi, gx_lo = gdn_x if gdn_x is not None else (None, None)
    if gx_hi is not None and (gx_hi.shape[-1] != pitch or tuple(gx_hi.shape[:3]) != (images, ho, wo)):
        raise ValueError('gdn_x planes %s do not match the output planes' % (tuple(gx_hi.shape),))
    for c0, n, w_hi, w_lo in tiles:
        d = TcSplitExDesc(images, H, W, C, n, kh, kw, stride, pad, mode, ho, wo, pitch, c0, c_out, 1 if in_nhwc else 0, act, float(slope),
                          -1 if pad_x is None else pad_x, 2 if out_parity is not None else 1,
                          out_parity[0] if out_parity is not None else 0, out_parity[1] if out_parity is not None else 0, full_h, full_w)
        tag = '%s[%d->%d/%d,k%d,s%d,m%d]' % (name, C, n, c_out, kh, stride, mode)
        flops = 2.0 * images * ho * wo * n * C * kh * kw
        nbytes = 4.0 * (x_hi.numel() + images * ho * wo * n)
        with torch.cuda.device(dev), _launch(tag, flops=flops, nbytes=nbytes):
            check(_lib().sc2_tc_split_conv_ex(ctypes.byref(d), _ptr(x_hi), _ptr(x_lo), _ptr(w_hi), _ptr(w_lo), _ptr(v), _ptr(m),
                                              _ptr(gx_hi), _ptr(gx_lo), _ptr(out_hi), _ptr(out_lo), _ptr(out_sym),
                                              _TILE_COUNTERS.next(), _stream_ptr()), 'sc2_tc_split_conv_ex')
    return out_sym if mode == _native.TCS_QUANT else (out_hi, out_lo)


def deconv5_parity_taps(padding):
    """{output parity: (kernel indexes of the taps, padding)} of the stride-1 sub-convolutions of ConvTranspose2d(k5, s2, padding):
    from o = 2 i - padding + k, out[2Y + r] = sum_j x[Y + j - pad] * w[taps[j]].  padding 2 (with output_padding 1: 2H outputs) is
    DECONV5_TAPS; padding 1 (output_padding 0: 2H + 1 outputs, the hyperprior bottlenecks' h_s) has H + 1 even and H odd outputs."""
    if padding == 2:
        return DECONV5_TAPS
    if padding == 1:
        return {0: ([3, 1], 1), 1: ([4, 2, 0], 1)}
    raise ValueError('ConvTranspose2d(k5, s2) with padding %d is not covered' % padding)


def pack_deconv5_weight_split_tiles(weight, padding=2, c_in_pad=None):
    """ConvTranspose2d(k5, s2, padding) weight [c_in, c_out, 5, 5] -> {(py, px): tiles of the stride-1 sub-convolution that produces
    the output pixels of that parity} (pack_deconv5_weight_f16 in split precision)."""
    w = weight.detach().float()
    if tuple(w.shape[2:]) != (5, 5):
        raise ValueError('pack_deconv5_weight_split_tiles is for 5x5 kernels')
    taps = deconv5_parity_taps(padding)
    packs = {}
    for py in (0, 1):
        for px in (0, 1):
            ky, kx = taps[py][0], taps[px][0]
            sub = w[:, :, ky][:, :, :, kx].permute(1, 0, 2, 3).contiguous()     # conv weight [c_out, c_in, Ty, Tx]
            packs[(py, px)] = pack_conv_weight_split_tiles(sub, c_in_pad=c_in_pad)
    return packs


def unsplit_to_nchw(hi, lo, channels):
    """split planes [B, H, W, C_pad] -> fp32 NCHW [B, channels, H, W] (value = hi + lo / 2048)"""
    y = torch.add(hi[..., :channels].float(), lo[..., :channels].float(), alpha=1.0 / LO_SCALE)
    return y.permute(0, 3, 1, 2).contiguous()


def abs_split(hi, lo):
    """|a| of split planes: |hi|, lo with the sign of hi folded in (a = hi + lo / 2048; hi = 0 leaves |lo|)"""
    neg = (hi < 0) | ((hi == 0) & (lo < 0))
    return hi.abs(), torch.where(neg, -lo, lo)


def tc_first_layer(x, w_hi, w_lo, c_out, kh, kw, pad):
    """Stride-2 first conv on an fp32 NCHW image with the im2col fused into the tensor-core kernel (sc2_tc_first_layer).
    Returns split parity planes (hi, lo) of shape [B * 4, h_out/2, w_out/2, c_out rounded up to 8]."""
    require_cuda(x, 'tc_first_layer')
    x = x.contiguous().float()
    B, C, H, W = x.shape
    ho, wo = (H + 2 * pad - kh) // 2 + 1, (W + 2 * pad - kw) // 2 + 1
    out_c = (c_out + 7) // 8 * 8
    hi = torch.empty((B * 4, ho // 2, wo // 2, out_c), dtype=torch.float16, device=x.device)
    lo = torch.empty_like(hi)
    with torch.cuda.device(x.device), _launch('tc_first[%d->%d,k%d,s2]' % (C, c_out, kh), flops=2.0 * B * ho * wo * c_out * C * kh * kw,
                                              nbytes=4.0 * (x.numel() + B * ho * wo * c_out)):
        check(_lib().sc2_tc_first_layer(_ptr(x), B, C, H, W, kh, kw, pad, c_out, _ptr(w_hi), _ptr(w_lo), _ptr(hi), _ptr(lo), out_c,
                                        _TILE_COUNTERS.next(), _stream_ptr()), 'sc2_tc_first_layer')
    return hi, lo


# ----------------------------------------------------------------------------------------------
# fused g_a kernels (round 2): conv + GDN1 back to back, stacked (hi; lo) weights
# ----------------------------------------------------------------------------------------------
def pack_conv_weight_stacked(weight, n=None, c_in_pad=None):
    """Conv2d weight [c_out, c_in, kh, kw] -> [kh*kw, 2n, c_in_pad] fp16: per tap, rows [0, n) hold the hi halves and rows
    [n, 2n) the lo halves (value = hi + lo / 2048) of the K-contiguous weights; rows beyond c_out are zero.  One stacked
    N = 2n MMA then yields hi.hi and hi.lo from a single read of the activations (conv_ga_halo.cu)."""
    w = weight.detach().float()
    c_out, c_in, kh, kw = w.shape
    n = n or _lib().sc2_ga_halo_n(c_out)
    if not n or n < c_out:
        raise ValueError('c_out %d is not supported by the fused g_a kernels' % c_out)
    c_in_pad = c_in_pad or (c_in + 15) // 16 * 16
    full = torch.zeros((kh * kw, n, c_in_pad), dtype=torch.float32, device=w.device)
    full[:, :c_out, :c_in] = w.permute(2, 3, 0, 1).reshape(kh * kw, c_out, c_in)
    hi, lo = split_f16(full)
    return torch.cat([hi, lo], dim=1).contiguous()


def ga_halo_conv_gdn(x_hi, x_lo, w_stack, gamma_stack, beta, c_out, kh, kw, pad):
    """sc2_ga_halo_conv_gdn: stride-2 conv on parity planes [images * 4, H, W, C] + GDN1, one kernel.
    Returns split planes (hi, lo) [images, h_out, w_out, c_out rounded up to 8]."""
    require_cuda(x_hi, 'ga_halo_conv_gdn')
    planes, H, W, C = x_hi.shape
    images = planes // 4
    ho, wo = (2 * H + 2 * pad - kh) // 2 + 1, (2 * W + 2 * pad - kw) // 2 + 1
    out_c = (c_out + 7) // 8 * 8
    dev = x_hi.device
    out_hi = torch.empty((images, ho, wo, out_c), dtype=torch.float16, device=dev)
    out_lo = torch.empty_like(out_hi)
    d = GaHaloDesc(images, H, W, C, c_out, kh, kw, pad, ho, wo, out_c)
    b = beta.detach().contiguous().float()
    flops = 2.0 * images * ho * wo * c_out * (C * kh * kw + c_out)
    nbytes = 4.0 * (x_hi.numel() + images * ho * wo * c_out)
    with torch.cuda.device(dev), _launch('ga_halo[%d->%d,k%d,s2]+gdn1' % (C, c_out, kh), flops=flops, nbytes=nbytes):
        check(_lib().sc2_ga_halo_conv_gdn(ctypes.byref(d), _ptr(x_hi), _ptr(x_lo), _ptr(w_stack), _ptr(gamma_stack), _ptr(b),
                                          _ptr(out_hi), _ptr(out_lo), _TILE_COUNTERS.next(), _stream_ptr()), 'sc2_ga_halo_conv_gdn')
    return out_hi, out_lo


def pack_first_layer_stacked(weight):
    """Conv2d(3 -> c_out, k5) weight -> [2n, 80] fp16 stacked (hi; lo) of weight.reshape(c_out, 75) (K order (c, dy, dx))."""
    w = weight.detach().float()
    c_out, c_in, kh, kw = w.shape
    if (c_in, kh, kw) != (3, 5, 5):
        raise ValueError('the fused first layer is Conv2d(3 -> C, k5, s2, p2)')
    n = _lib().sc2_ga_halo_n(c_out)
    if not n:
        raise ValueError('c_out %d is not supported by the fused g_a kernels' % c_out)
    full = torch.zeros((n, 80), dtype=torch.float32, device=w.device)
    full[:c_out, :75] = w.reshape(c_out, 75)
    hi, lo = split_f16(full)
    return torch.cat([hi, lo], dim=0).contiguous()


def normalize_lut(mean, std, device):
    """float32 [3, 256]: the value of byte v of channel c after torchvision's ToTensor (v / 255) and Normalize
    ((x - mean) / std), computed with the same torch ops in the same order so that the table is bit-identical to the loader."""
    v = torch.arange(256, dtype=torch.float32).div(255)
    m = torch.as_tensor(mean, dtype=torch.float32).view(-1, 1)
    s = torch.as_tensor(std, dtype=torch.float32).view(-1, 1)
    return v.view(1, 256).sub(m).div(s).contiguous().to(device)


def normalize_u8(x, lut):
    """uint8 NCHW image -> fp32 through the look-up table (the route for shapes the fused first layer does not cover)."""
    require_cuda(x, 'normalize_u8')
    C = x.shape[1]
    idx = x.long() + (torch.arange(C, device=x.device) * 256).view(1, C, 1, 1)
    return lut.reshape(-1)[idx]


def ga_first_conv_gdn(x, w_stack, gamma_stack, beta, c_out, lut=None):
    """sc2_ga_first_conv_gdn: Conv2d(3 -> c_out, k5, s2, p2) + GDN1 on an NCHW image (fp32, or uint8 with `lut`).
    Returns split parity planes (hi, lo) [B * 4, h_out/2, w_out/2, c_out rounded up to 8]."""
    require_cuda(x, 'ga_first_conv_gdn')
    u8 = x.dtype == torch.uint8
    if u8 and lut is None:
        raise ValueError('uint8 images need the normalisation look-up table (ops.normalize_lut)')
    x = x.contiguous() if u8 else x.contiguous().float()
    B, C, H, W = x.shape
    ho, wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out_c = (c_out + 7) // 8 * 8
    hi = torch.empty((B * 4, ho // 2, wo // 2, out_c), dtype=torch.float16, device=x.device)
    lo = torch.empty_like(hi)
    b = beta.detach().contiguous().float()
    flops = 2.0 * B * ho * wo * c_out * (75 + c_out)
    nbytes = 4.0 * (x.numel() + B * ho * wo * c_out)
    with torch.cuda.device(x.device), _launch('ga_first[3->%d,k5,s2]+gdn1%s' % (c_out, ',u8' if u8 else ''), flops=flops, nbytes=nbytes):
        check(_lib().sc2_ga_first_conv_gdn(_ptr(x), int(u8), _ptr(lut), B, H, W, c_out, _ptr(w_stack), _ptr(gamma_stack), _ptr(b),
                                           _ptr(hi), _ptr(lo), out_c, _TILE_COUNTERS.next(), _stream_ptr()), 'sc2_ga_first_conv_gdn')
    return hi, lo


# ----------------------------------------------------------------------------------------------
# extended tensor-core convolution (round 2): transposed convolutions as parity sub-convolutions, GDN proper
# ----------------------------------------------------------------------------------------------
DECONV5_TAPS = {0: ([4, 2, 0], 1), 1: ([3, 1], 0)}  # output parity -> (kernel indexes of the taps, padding) of ConvTranspose2d(k5, s2, p2, op1)


def pack_deconv5_weight_f16(weight, rows_pad=None):
    """ConvTranspose2d(k5, s2, p2, output_padding 1) weight [c_in, c_out, 5, 5] -> {(py, px): [taps, rows, c_in_pad] fp16}: the four
    stride-1 sub-convolutions that produce the output pixels of parity (py, px).  out[2Y + py] = sum_t x[Y + t - pad] * w[k_t] with
    (k, pad) = ([4, 2, 0], 1) for py = 0 and ([3, 1], 0) for py = 1 (from o = 2 i - 2 + k), the same along x."""
    w = weight.detach().float()
    c_in, c_out, kh, kw = w.shape
    if (kh, kw) != (5, 5):
        raise ValueError('pack_deconv5_weight_f16 is for 5x5 kernels')
    c_in_pad = (c_in + 63) // 64 * 64
    rows = rows_pad or c_out
    packs = {}
    for py in (0, 1):
        for px in (0, 1):
            ky, kx = DECONV5_TAPS[py][0], DECONV5_TAPS[px][0]
            sub = w[:, :, ky][:, :, :, kx]                                   # [c_in, c_out, Ty, Tx]
            packed = torch.zeros((len(ky) * len(kx), rows, c_in_pad), dtype=torch.float16, device=w.device)
            packed[:, :c_out, :c_in] = sub.permute(2, 3, 1, 0).reshape(len(ky) * len(kx), c_out, c_in).half()
            packs[(py, px)] = packed.contiguous()
    return packs


def tc_conv_ex(x_nhwc, w_packed, kh, kw, pad_y, pad_x, mode, grid, out, out_stride=1, out_py=0, out_px=0, vec=None, gdn_x=None,
               out2=None, c_out=None, c_in=None, tag=None, x_lo=None, gdn_x_lo=None, out3=None):
    """sc2_tc_conv_ex: one launch of the generalised tcgen05 convolution.  `out` (and `out2`) are caller-allocated full output
    tensors ([B, out_h, out_w, c_out] NHWC, or [B, c_out, out_h, out_w] fp32 for mode TC_NCHW_F32_CLAMP); `grid` = (h_out, w_out)
    output pixels computed by this launch, placed at (oy * out_stride + out_py, ox * out_stride + out_px)."""
    require_cuda(x_nhwc, 'tc_conv_ex')
    assert x_nhwc.dtype == torch.float16 and x_nhwc.is_contiguous() and w_packed.dtype == torch.float16
    B, H, W, Cp = x_nhwc.shape
    taps, rows, cp2 = w_packed.shape
    if taps != kh * kw or cp2 != Cp:
        raise ValueError('packed weight does not match the activation / kernel size')
    nchw = mode == _native.TC_NCHW_F32_CLAMP
    out_h, out_w = (out.shape[2], out.shape[3]) if nchw else (out.shape[1], out.shape[2])
    c_out = c_out or rows
    d = TcConvExDesc(B, H, W, Cp, c_out, kh, kw, pad_y, pad_x, mode, grid[0], grid[1], out_h, out_w, out_stride, out_py, out_px)
    v = vec.detach().contiguous().float() if vec is not None else None
    c_real = c_in or Cp
    flops = 2.0 * B * grid[0] * grid[1] * c_out * c_real * kh * kw  # (algorithmic: the second pass over x_lo is not counted)
    nbytes = 4.0 * B * (H * W * c_real / (out_stride * out_stride) + grid[0] * grid[1] * c_out)
    tag = tag or 'tc_conv_ex[%d->%d,k%dx%d,m%d]' % (Cp, c_out, kh, kw, mode)
    with torch.cuda.device(x_nhwc.device), _launch(tag, flops=flops, nbytes=nbytes):
        check(_lib().sc2_tc_conv_ex(ctypes.byref(d), _ptr(x_nhwc), _ptr(x_lo), _ptr(w_packed), _ptr(v), _ptr(gdn_x), _ptr(gdn_x_lo), _ptr(out),
                                    _ptr(out2), _ptr(out3), None, _TILE_COUNTERS.next(), _stream_ptr()), 'sc2_tc_conv_ex')
    return out
