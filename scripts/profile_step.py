"""One batch through the hot path, twice per coder layout (first pass warms up) -- the target of the ncu captures.
python scripts/profile_step.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import sc2bench_b200 as s2  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device('cuda:0')
torch.manual_seed(0)
layer = s2.get_layer('FPBasedResNetBottleneck').eval()
layer.update()
layer.to(dev)
layer.native_calls = False  # per-kernel route: the coder layout below is honoured per call
x = torch.randn(batch, 3, 224, 224, device=dev)
with torch.inference_mode():
    for layout in (None, None, 'lanes', 'lanes'):
        layer.entropy_bottleneck.coder_layout = layout
        streams, shape = layer.encode_packed(x)
        out = layer.decode_packed(streams, shape)
        torch.cuda.synchronize()
print('ok', tuple(out.shape), streams.total_bytes())
