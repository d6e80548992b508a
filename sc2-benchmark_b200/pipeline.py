"""Software pipeline over batches for the bottleneck path (throughput mode).

`FPBasedResNetBottleneck.encode` / `.decode` (sc2bench/models/layer.py:496-521) are one batch at a time: g_a, coder, g_s in
sequence, and the coder -- one serial rANS chain per image -- leaves the tensor cores idle for most of the step.  Batches are
independent, so a caller with a queue of batches can overlap them.  Doing that with one CUDA stream per batch is not enough:
the batches drift into lock-step (all in their transforms, then all in their coders; scripts/diag_trace.py).  CodecPipeline
fixes the order instead:

    transform stream:  g_a(0) g_a(1) ... g_a(d)  g_s(0) g_a(d+1)  g_s(1) g_a(d+2) ...
    batch streams:           coder(0) coder(1) ...          (lane-per-stream layout: a batch's coder is one block)

Results come back in submission order, `depth` submissions late.  The host is held back only so that it stays at most
`depth + max_ahead` batches ahead of the GPU: left alone it queues the whole run at once, the caching allocator then has to
cudaMalloc fresh blocks for every batch in flight, and those calls stall the host until the GPU starves (measured: the same
command ran at 12 k or 43 k images/s depending on who won).
Needs one hardware queue per stream: CUDA_DEVICE_MAX_CONNECTIONS >= depth + 3 (the package sets 32 at import unless the
variable is already set); with streams sharing queues a waiting coder kernel blocks the transforms queued behind it.
"""
import collections
import logging
import time

import torch

from . import _native
from .native_codec import FpNativeCodec, Slot

_log = logging.getLogger('sc2bench_b200')


class PipelineResult:
    """One batch out of the pipeline: device-resident bitstreams and decoder features, valid once `ready` has happened."""

    def __init__(self, streams, shape, features, ready):
        self.streams, self.shape, self.features, self.ready = streams, shape, features, ready

    def wait(self, stream=None):
        """Orders `stream` (default: the current stream) after this batch; no host synchronisation."""
        stream = stream or torch.cuda.current_stream()
        stream.wait_event(self.ready)
        self.features.record_stream(stream)
        return self


class CodecPipeline:
    """depth: batches whose coders overlap the transforms of the others; max_ahead: batches the host may additionally run ahead.
    native (default): a batch is two C calls into preallocated slots (native_codec.FpNativeCodec) -- `streams` / `features` of a
    result are then VIEWS of ring buffers, valid until `depth + 1` more batches have been submitted (clone to keep them).
    coder_sms: SMs left to the coder blocks while the pipeline exists (sc2_set_persistent_ctas; 0 = none)."""

    def __init__(self, layer, depth=16, max_ahead=4, native=True, coder_sms=12):
        if depth < 1:
            raise ValueError('depth must be >= 1')
        self.layer, self.depth, self.max_ahead = layer, depth, max(0, max_ahead)
        self._retired = collections.deque()  # ready events of retired batches the GPU may not have finished yet
        device = layer.entropy_bottleneck._quantized_cdf.device
        if device.type != 'cuda':
            raise RuntimeError('CodecPipeline: the sc2bench_b200 hot path runs on CUDA only; move the layer to a GPU')
        self.device = device
        self.transform_stream = layer.use_transform_stream(True)
        self.batch_streams = [torch.cuda.Stream(device=device) for _ in range(depth + 1)]
        self._pending = collections.deque()
        self._submitted = 0
        self.wait_s = 0.0  # host time spent in back-pressure waits (diagnostics: issue time = loop time - wait_s)
        self._want_native, self._native, self._slots, self._native_shape = bool(native), None, None, None
        self._coder_sms = int(coder_sms)
        if self._coder_sms > 0:
            _native.check(_native.load().sc2_set_persistent_ctas(148 - self._coder_sms), 'sc2_set_persistent_ctas')

    def _native_for(self, x):
        """The one-call-per-batch codec for batches shaped like x (built at the first batch), or None (per-kernel route)."""
        if not self._want_native:
            return None
        shape = (tuple(x.shape), x.dtype)
        if self._native is not None and shape == self._native_shape and not self._native.stale():
            return self._native
        if self._pending:
            return None if shape != self._native_shape else self._native  # (never switch routes with batches in flight)
        try:
            with torch.cuda.stream(self.transform_stream):
                self._native = FpNativeCodec(self.layer, x.shape[0], x.shape[2], x.shape[3], self.device)
                self._slots = [self._native.new_slot(s) for s in self.batch_streams]
            torch.cuda.current_stream(self.device).wait_stream(self.transform_stream)
            self._native_shape = shape
        except (ValueError, _native.NativeError) as e:
            _log.info('CodecPipeline: per-kernel route (%s)', e)
            self._want_native, self._native = False, None
        return self._native

    @torch.no_grad()
    def submit(self, x):
        """Queues g_a + coder of batch x (a CUDA tensor produced on the current stream).  Returns the PipelineResult of the
        batch submitted `depth` calls earlier, or None while the pipeline fills."""
        if len(self._retired) > self.max_ahead:  # back-pressure: bounded work (and memory) in flight
            t0 = time.perf_counter()
            while len(self._retired) > self.max_ahead:
                self._retired.popleft().synchronize()
            self.wait_s += time.perf_counter() - t0
        i = self._submitted % len(self.batch_streams)
        s = self.batch_streams[i]
        self._submitted += 1
        codec = self._native_for(x)
        if codec is not None:
            slot = self._slots[i]
            cur = torch.cuda.current_stream(self.device)
            slot.ev_in.record(cur)
            x.record_stream(self.transform_stream)
            streams = codec.encode(x, slot, self.transform_stream, s, ev_in=slot.ev_in)
            self._pending.append((slot, (streams, codec.latent_hw)))
        else:
            s.wait_stream(torch.cuda.current_stream())
            x.record_stream(s)
            with torch.cuda.stream(s):
                encoded = self.layer.encode_packed(x)
            self._pending.append((s, encoded))
        return self._retire() if len(self._pending) > self.depth else None

    @torch.no_grad()
    def _retire(self):
        s, (streams, shape) = self._pending.popleft()
        if isinstance(s, Slot):
            features = self._native.decode(s, self.transform_stream, s.stream)
            ready = s.ev_out
        else:
            with torch.cuda.stream(s):
                features = self.layer.decode_packed(streams, shape)
                ready = torch.cuda.Event()
                ready.record(s)
        self._retired.append(ready)
        return PipelineResult(streams, shape, features, ready)

    def drain(self):
        """Retires every batch still in flight, in order."""
        out = []
        while self._pending:
            out.append(self._retire())
        return out

    def close(self):
        self.drain()
        self.layer.use_transform_stream(None)
        if self._coder_sms > 0:
            _native.load().sc2_set_persistent_ctas(0)
