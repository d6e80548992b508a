"""One C call per batch for FPBasedResNetBottleneck (csrc/fp_codec.cu): plan, workspaces and slots.

`FpNativeCodec` packs what `FPBasedResNetBottleneck.encode` / `.decode` (sc2bench/models/layer.py:496-521) need on the device into
a `sc2_fp_plan`, allocates the two transform workspaces ONCE, and hands out `Slot`s: per-batch buffers (symbols, coder arena,
packed bitstreams, dequantised latent, decoder features, events) that are reused batch after batch.  `encode(x, slot, ...)` and
`decode(slot, ...)` are then one ctypes call each -- no tensor allocation, no per-kernel Python (the host side of a pipelined step
drops from ~1 ms, and 3-8 ms on a busy core, to ~0.1 ms).  Results are bit-identical to the per-kernel route (same kernels).
"""
import ctypes

import torch

from . import _native, ops
from ._native import FpPlan, check


class Slot:
    """Buffers of ONE batch in flight.  `streams` / `features` are views of these buffers: valid until the slot is used again."""

    def __init__(self, codec, stream=None):
        dev, B = codec.device, codec.batch
        n = codec.symbols_per_image
        self.stream = stream
        self.symbols = torch.empty((B, codec.c3, codec.latent_hw[0], codec.latent_hw[1]), dtype=torch.int32, device=dev)
        self.slot_bytes = codec.slot_bytes
        self.arena = torch.empty(B * self.slot_bytes, dtype=torch.uint8, device=dev)
        self._packed = self._out = None  # (allocated on first use: callers that bring their own never pay for them)
        self._codec_shape = (B, codec.out_hw[0], codec.out_hw[1], codec.d3)
        self.lengths = torch.empty(B, dtype=torch.int32, device=dev)
        self.offsets = torch.empty(B + 1, dtype=torch.int64, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self.latent_hat = torch.empty((B, codec.c3, codec.latent_hw[0], codec.latent_hw[1]), dtype=torch.float32, device=dev)
        self.tile_counters = torch.zeros(8, dtype=torch.int32, device=dev)
        self.ev_in, self.ev_mid, self.ev_enc, self.ev_mid2, self.ev_out = (torch.cuda.Event() for _ in range(5))
        for ev in (self.ev_in, self.ev_mid, self.ev_enc, self.ev_mid2, self.ev_out):
            ev.record(torch.cuda.current_stream(dev))  # (creates the CUDA event: the handle is passed across the C ABI)
        self.n_symbols = n

    @property
    def packed(self):
        if self._packed is None:
            with torch.inference_mode(False):
                self._packed = torch.empty(self.arena.numel(), dtype=torch.uint8, device=self.arena.device)
        return self._packed

    @property
    def out(self):
        if self._out is None:
            with torch.inference_mode(False):
                self._out = torch.empty(self._codec_shape, dtype=torch.float32, device=self.arena.device)
        return self._out

    @property
    def features(self):
        """decoder output, logically NCHW (channels-last storage), like the per-kernel route returns it"""
        return self.out.permute(0, 3, 1, 2)


def _ev(e):
    return ctypes.c_void_p(e.cuda_event) if e is not None else ctypes.c_void_p(0)


class FpNativeCodec:
    def __init__(self, layer, batch, h_in, w_in, device, coder_layout='throughput'):
        from .bottleneck import TensorCoreAnalysis, TensorCoreTransform, _param_key
        device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError('FpNativeCodec: the sc2bench_b200 hot path runs on CUDA only')
        why = TensorCoreAnalysis.why_not(layer.encoder, (batch, 3, h_in, w_in)) or TensorCoreTransform.why_not(layer.decoder)
        if why is not None or layer.encoder_precision != 'split-tc' or layer.decoder_precision != 'fp16-tc':
            raise ValueError('the fused tensor-core plans do not cover this layer: %s' % (why or 'precision settings'))
        self.layer, self.batch, self.device = layer, int(batch), device
        self.coder_layout = _native.rans_layout(coder_layout, int(batch))
        if layer._tc_encoder is None:
            layer._tc_encoder = TensorCoreAnalysis(layer.encoder)
        if layer._tc_decoder is None:
            layer._tc_decoder = TensorCoreTransform(layer.decoder)
        enc = layer._tc_encoder._prepare()
        _, steps, _ = layer._tc_decoder._prepare()
        if enc['first'] is None or enc['mid'] is None:
            raise ValueError('the fused g_a kernels do not cover this encoder')
        kinds = [s[0] for s in steps]
        modes = [s[4] for s in steps]
        T = _native
        if kinds != ['conv', 'gdn', 'conv', 'gdn', 'conv'] or modes != [T.TC_STORE_ABS_F16, T.TC_IGDN1_ABS_F16, T.TC_STORE_ABS_F16,
                                                                          T.TC_IGDN1_ABS_F16, T.TC_STORE_F32]:
            raise ValueError('the decoder is not Conv - IGDN1 - Conv - IGDN1 - Conv on the |x| + sign plan')
        self._key = (_param_key(layer.encoder), _param_key(layer.decoder))
        c1, _, c2, _, c3 = list(layer.encoder)
        d1, _, d2, _, d3 = list(layer.decoder)
        eb = layer.entropy_bottleneck
        tables = eb.coder_tables()
        with torch.inference_mode(False):
            self._keep = {
                'w1': enc['first'][0], 'g1': enc['first'][1], 'beta1': enc['gdn'][0][1],
                'w2': enc['mid'][0], 'g2': enc['mid'][1], 'beta2': enc['gdn'][1][1],
                'w3h': enc['w3'][0], 'w3l': enc['w3'][1],
                'medians': eb._get_medians().detach().reshape(-1).float().contiguous().clone(),
                'tables': tables.on(device),
                'wd1': steps[0][1], 'gd1': steps[1][1], 'betad1': steps[1][5], 'wd2': steps[2][1], 'gd2': steps[3][1], 'betad2': steps[3][5],
                'wd3': steps[4][1],
                'lut': layer._input_lut(device) if getattr(layer, '_input_norm', None) is not None else None,
            }
        k = self._keep
        p = FpPlan()
        p.batch, p.h_in, p.w_in = self.batch, int(h_in), int(w_in)
        p.c1, p.c2, p.c3 = c1.out_channels, c2.out_channels, c3.out_channels
        p.k1, p.k2, p.k3, p.p3 = c1.kernel_size[0], c2.kernel_size[0], c3.kernel_size[0], c3.padding[0]
        p.d1, p.d2, p.d3 = d1.out_channels, d2.out_channels, d3.out_channels
        p.kd1, p.pd1, p.kd2, p.pd2, p.kd3, p.pd3 = (d1.kernel_size[0], d1.padding[0], d2.kernel_size[0], d2.padding[0],
                                                    d3.kernel_size[0], d3.padding[0])
        p.n_rows, p.cdf_stride = tables.n_rows, tables.cdf_stride
        for name, key in (('w1_stack', 'w1'), ('g1_stack', 'g1'), ('beta1', 'beta1'), ('w2_stack', 'w2'), ('g2_stack', 'g2'), ('beta2', 'beta2'),
                          ('w3_hi', 'w3h'), ('w3_lo', 'w3l'), ('medians', 'medians'), ('lut', 'lut'), ('tables', 'tables'), ('wd1', 'wd1'),
                          ('gd1', 'gd1'), ('wd2', 'wd2'), ('gd2', 'gd2'), ('wd3', 'wd3'), ('betad1', 'betad1'), ('betad2', 'betad2')):
            t = k[key]
            setattr(p, name, t.data_ptr() if t is not None else None)
        self.plan = p
        ga, gs, nsym = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        lh, lw, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        lib = _native.load()
        check(lib.sc2_fp_workspace_bytes(ctypes.byref(p), ctypes.byref(ga), ctypes.byref(gs), ctypes.byref(nsym), ctypes.byref(lh),
                                         ctypes.byref(lw), ctypes.byref(oh), ctypes.byref(ow)), 'sc2_fp_workspace_bytes')
        self.symbols_per_image, self.latent_hw, self.out_hw = int(nsym.value), (lh.value, lw.value), (oh.value, ow.value)
        self.c3, self.d3 = p.c3, p.d3
        self.slot_bytes = int(lib.sc2_rans_max_stream_bytes(self.symbols_per_image))
        with torch.inference_mode(False):
            # ONE workspace per transform (the transforms of all batches run on one stream, in order)
            self.ws_ga = torch.empty(ga.value, dtype=torch.uint8, device=device)
            self.ws_gs = torch.empty(gs.value, dtype=torch.uint8, device=device)
        self._lib = lib

    def stale(self):
        from .bottleneck import _param_key
        return self._key != (_param_key(self.layer.encoder), _param_key(self.layer.decoder))

    def new_slot(self, stream=None):
        with torch.inference_mode(False), torch.cuda.device(self.device):
            return Slot(self, stream)

    def make_workspaces(self):
        """A private (g_a, g_s) workspace pair, for callers whose transforms do NOT share one stream."""
        with torch.inference_mode(False), torch.cuda.device(self.device):
            return (torch.empty(self.ws_ga.numel(), dtype=torch.uint8, device=self.device),
                    torch.empty(self.ws_gs.numel(), dtype=torch.uint8, device=self.device))

    def encode(self, x, slot, transform_stream, coder_stream, ev_in=None, packed=None, offsets=None, status=None, ws=None):
        """g_a + coder of batch x into `slot` (or into the caller's packed / offsets / status tensors).  transform_stream waits for
        ev_in; slot.ev_enc marks the packed streams ready."""
        if tuple(x.shape) != (self.batch, 3, self.plan.h_in, self.plan.w_in) or not x.is_contiguous():
            raise ValueError('batch shape %s does not match the plan' % (tuple(x.shape),))
        u8 = x.dtype == torch.uint8
        if u8 and self._keep['lut'] is None:
            raise ValueError('uint8 images need set_input_normalization(mean, std) before the codec is built')
        if not u8 and x.dtype != torch.float32:
            raise ValueError('images must be float32 or uint8')
        ops.STATS['launches'] += 6
        packed = slot.packed if packed is None else packed
        offsets = slot.offsets if offsets is None else offsets
        status = slot.status if status is None else status
        ws = self.ws_ga if ws is None else ws
        with torch.cuda.device(self.device):
            check(self._lib.sc2_fp_encode_batch(ctypes.byref(self.plan), x.data_ptr(), int(u8), ws.data_ptr(), slot.symbols.data_ptr(),
                                                slot.arena.data_ptr(), slot.slot_bytes, slot.lengths.data_ptr(), packed.data_ptr(),
                                                offsets.data_ptr(), status.data_ptr(), slot.tile_counters.data_ptr(),
                                                self.coder_layout, ctypes.c_void_p(transform_stream.cuda_stream),
                                                ctypes.c_void_p(coder_stream.cuda_stream), _ev(ev_in), _ev(slot.ev_mid), _ev(slot.ev_enc)),
                  'sc2_fp_encode_batch')
        streams = ops.PackedStreams.__new__(ops.PackedStreams)
        streams.packed, streams.offsets, streams.batch, streams.status, streams._offs, streams.ready = \
            packed, offsets, self.batch, status, None, slot.ev_enc
        return streams

    def decode(self, slot, transform_stream, coder_stream, ev_in=None, packed=None, offsets=None, status=None, out=None, ws=None,
               latent_hat=None):
        """coder + g_s of the streams in `slot` (or of `packed` / `offsets`); slot.ev_out marks the output ready.  `out`: the
        caller's [B, H, W, C] fp32 tensor instead of slot.out.  latent_hat: an already decoded latent -- only g_s runs."""
        only_gs = latent_hat is not None
        ops.STATS['launches'] += 6 if only_gs else 7
        packed = slot.packed if packed is None else packed
        offsets = slot.offsets if offsets is None else offsets
        status = slot.status if status is None else status
        out = slot.out if out is None else out
        ws = self.ws_gs if ws is None else ws
        lat = latent_hat if only_gs else slot.latent_hat
        with torch.cuda.device(self.device):
            check(self._lib.sc2_fp_decode_batch(ctypes.byref(self.plan), None if only_gs else packed.data_ptr(), offsets.data_ptr(),
                                                lat.data_ptr(), ws.data_ptr(), out.data_ptr(), status.data_ptr(),
                                                slot.tile_counters.data_ptr() + 12, self.coder_layout, ctypes.c_void_p(coder_stream.cuda_stream),
                                                ctypes.c_void_p(transform_stream.cuda_stream), _ev(ev_in), _ev(slot.ev_mid2), _ev(slot.ev_out)),
                  'sc2_fp_decode_batch')
        return out.permute(0, 3, 1, 2)
