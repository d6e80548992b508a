"""Pipelined (depth 8, batch 256) results vs the serial path: per batch, do the streams / features match? (diagnostic)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sc2bench_b200 as s2
depth = int(sys.argv[1]); steps = int(sys.argv[2])
dev = torch.device('cuda:0')
torch.manual_seed(0)
layer = s2.get_layer('FPBasedResNetBottleneck', num_bottleneck_channels=24, num_target_channels=256).eval()
layer.update(); layer.to(dev)
xs = [torch.randn(256, 3, 224, 224, device=dev) for _ in range(2)]
with torch.inference_mode():
    want = []
    for x in xs:
        st, shape = layer.encode_packed(x)
        want.append((st.lengths().copy(), st.packed[:st.total_bytes()].clone(), layer.decode_packed(st, shape).clone()))
    torch.cuda.synchronize()
    print('serial ok, bytes', [int(w[0].sum()) for w in want], flush=True)
    pipe = s2.CodecPipeline(layer, depth=depth, max_ahead=2)
    res = []
    t0 = time.time()
    for i in range(steps):
        r = pipe.submit(xs[i & 1])
        if r is not None: res.append(r)
    res += pipe.drain()
    torch.cuda.synchronize()
    print('pipeline done %.1f ms/step' % ((time.time() - t0) * 1e3 / steps), flush=True)
    bad = 0
    for i, r in enumerate(res):
        w = want[i & 1]
        ln = r.streams.lengths()
        same_len = (ln == w[0]).all()
        same_bytes = same_len and torch.equal(r.streams.packed[:int(ln.sum())], w[1])
        same_feat = torch.equal(r.features, w[2])
        if not (same_bytes and same_feat):
            bad += 1
            print('batch', i, 'len', bool(same_len), 'bytes', bool(same_bytes), 'feat', bool(same_feat), 'total', int(ln.sum()), flush=True)
    print('mismatching batches:', bad, 'of', len(res))
