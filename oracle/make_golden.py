"""oracle/make_golden.py -- regenerates tests/golden/*.npz.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):
    python oracle/make_golden.py
It imports the REFERENCE'S OWN model code (/root/reference/sc2bench/models/{layer,backbone,wrapper}.py)
on top of oracle/shim (the CPU restatement of CompressAI + import stubs for torchdistill/timm), runs the
bottleneck path on seeded inputs and freezes inputs, tables, symbols, bitstreams and outputs.

PARITY UNPINNED: CompressAI itself cannot run here, so the CompressAI half of every vector comes from the
restatement (oracle/shim/compressai, oracle/rans_oracle.c).  The sc2bench half (topology, call order,
encode()/decode() contract) is the reference's real code.  Anyone with a real `compressai` install can
re-run this script with `--real-compressai` to close the loop: it then skips the shim for compressai.
"""
import argparse
import hashlib
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, 'tests', 'golden')


def _setup_paths(real_compressai):
    shim = os.path.join(HERE, 'shim')
    if real_compressai:
        # keep only the torchdistill/timm stubs visible
        import importlib.util
        if importlib.util.find_spec('compressai') is None:
            raise SystemExit('--real-compressai given but compressai is not importable')
        sys.path.append(shim)
    else:
        sys.path.insert(0, shim)
    sys.path.insert(0, '/root/reference')
    sys.path.insert(0, HERE)


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _bytes_to_u8(b):
    return np.frombuffer(b, dtype=np.uint8).copy()


def _pack_strings(strings):
    """list[bytes] -> (concatenated u8, int64 offsets[len+1])"""
    offs = np.zeros(len(strings) + 1, dtype=np.int64)
    for i, s in enumerate(strings):
        offs[i + 1] = offs[i] + len(s)
    return _bytes_to_u8(b''.join(strings)), offs


def rans_cases(torch):
    """Coder-level known-answer vectors on real EntropyBottleneck / GaussianConditional tables."""
    import cref
    import pyrans
    from compressai.entropy_models import EntropyBottleneck, GaussianConditional
    from compressai.models import get_scale_table
    out = {}
    torch.manual_seed(0)
    eb = EntropyBottleneck(24)
    eb.update()
    cdf, ln, off = eb._quantized_cdf.numpy(), eb._cdf_length.numpy(), eb._offset.numpy()
    out['eb24_cdf'], out['eb24_len'], out['eb24_off'] = cdf, ln, off
    out['eb24_medians'] = eb._get_medians().detach().reshape(-1).numpy()
    rng = np.random.RandomState(1234)
    names = []

    def add(name, table, symbols, indexes, check_py=True):
        tcdf, tln, toff = table
        symbols = np.asarray(symbols, dtype=np.int32)
        indexes = np.asarray(indexes, dtype=np.int32)
        s = cref.encode_with_indexes(symbols, indexes, tcdf, tln, toff)
        if check_py:
            s_py = pyrans.encode_with_indexes(symbols.tolist(), indexes.tolist(), tcdf.tolist(), tln.tolist(), toff.tolist())
            assert s == s_py, name
            assert pyrans.decode_with_indexes(s, indexes.tolist(), tcdf.tolist(), tln.tolist(), toff.tolist()) == symbols.tolist(), name
        assert (cref.decode_with_indexes(s, indexes, tcdf, tln, toff) == symbols).all(), name
        out[name + '_symbols'], out[name + '_indexes'], out[name + '_stream'] = symbols, indexes, _bytes_to_u8(s)
        names.append(name)

    idx_eb = np.repeat(np.arange(24, dtype=np.int32), 25)
    for sigma in (1, 3, 8):
        add('eb24_sigma%d' % sigma, (cdf, ln, off), np.round(rng.randn(600) * sigma), idx_eb)
    add('empty', (cdf, ln, off), [], [])
    add('single', (cdf, ln, off), [3], [7])
    add('single_escape', (cdf, ln, off), [-11], [0])
    add('edge_values', (cdf, ln, off), [-10, 10, 11, -11, 9, -9, 0, 12, -12, 26, -26], np.zeros(11))
    add('huge_escapes', (cdf, ln, off),
        [2 ** 30, -2 ** 30, 123456789, -123456789, 65535, -65536, 2 ** 31 - 12, -(2 ** 31 - 12) // 2, 17, -17, 0],
        np.arange(11) % 24)
    # Gaussian-conditional tables (64 rows, ragged lengths up to 3133)
    gc = GaussianConditional(None)
    gc.update_scale_table(get_scale_table())
    gcdf, gln, goff = gc._quantized_cdf.numpy(), gc._cdf_length.numpy(), gc._offset.numpy()
    out['gc_scale_table'] = gc.scale_table.numpy()
    out['gc_len'], out['gc_off'] = gln, goff
    out['gc_cdf_sha256'] = np.array(_sha(gcdf))
    out['gc_cdf_shape'] = np.array(gcdf.shape)
    out['gc_cdf_row0'], out['gc_cdf_row31'] = gcdf[0, :gln[0]], gcdf[31, :gln[31]]
    out['gc_cdf_row63_head'], out['gc_cdf_row63_tail'] = gcdf[63, :64], gcdf[63, gln[63] - 64:gln[63]]
    gidx = rng.randint(0, 64, size=2000).astype(np.int32)
    gsym = np.round(rng.randn(2000) * gc.scale_table.numpy()[gidx] * 1.5)
    add('gc_mixed', (gcdf, gln, goff), gsym, gidx, check_py=True)
    out['case_names'] = np.array(names)
    # pmf_to_quantized_cdf known answers (incl. the zero-frequency "steal" path)
    pmfs = [np.array([0.1, 0.2, 0.3, 0.4], np.float32),
            np.array([0.5, 0.0, 0.0, 0.5, 1e-9], np.float32),
            np.array([1e-7] * 5 + [0.999] + [1e-7] * 5, np.float32),
            np.array([0.25] * 4, np.float32),
            np.abs(rng.randn(300)).astype(np.float32) ** 4 / 1000]
    for i, p in enumerate(pmfs):
        out['pmf%d' % i] = p
        out['pmf%d_cdf' % i] = cref.pmf_to_quantized_cdf(p, 16).astype(np.int64)
    out['n_pmfs'] = np.array(len(pmfs))
    np.savez_compressed(os.path.join(GOLD, 'rans_cases.npz'), **out)
    print('rans_cases.npz:', names)


def perturb_entropy_bottleneck(torch, eb, seed):
    """Make an EntropyBottleneck look trained: ragged quantiles, non-zero medians/factors."""
    g = torch.Generator().manual_seed(seed)
    C = eb.channels
    with torch.no_grad():
        med = (torch.rand(C, generator=g) - 0.5) * 3
        lo = 1.5 + torch.rand(C, generator=g) * 30
        hi = 1.5 + torch.rand(C, generator=g) * 30
        eb.quantiles[:, 0, 0] = med - lo
        eb.quantiles[:, 0, 1] = med
        eb.quantiles[:, 0, 2] = med + hi
        for f in eb.factors:
            f.copy_((torch.rand(f.shape, generator=g) - 0.5))
        for m in eb.matrices:
            m.add_((torch.rand(m.shape, generator=g) - 0.5) * 0.5)
    # stretch the density so that the tails are really light where the quantiles claim they are
    return eb


def small_fp_bottleneck(torch):
    """The reference's FPBasedResNetBottleneck (sc2bench/models/layer.py:444-550) at a tiny size, weights stored."""
    from sc2bench.models.layer import get_layer
    torch.manual_seed(7)
    layer = get_layer('FPBasedResNetBottleneck', num_input_channels=3, num_bottleneck_channels=8, num_target_channels=32)
    with torch.no_grad():  # dense, non-trivial GDN gammas / betas like a trained model
        for mod in list(layer.encoder) + list(layer.decoder):
            if hasattr(mod, 'gamma'):
                C = mod.gamma.shape[0]
                mod.gamma.copy_(mod.gamma_reparam.init(0.1 * torch.eye(C) + 0.02 * torch.rand(C, C)))
                mod.beta.copy_(mod.beta_reparam.init(0.5 + torch.rand(C)))
        for mod in layer.encoder:
            if isinstance(mod, torch.nn.Conv2d):
                mod.weight.mul_(4.0)  # widen the latent distribution: exercises +-k symbols and escapes
    perturb_entropy_bottleneck(torch, layer.entropy_bottleneck, 11)
    with torch.no_grad():
        q = layer.entropy_bottleneck.quantiles
        q[:, 0, 0] = q[:, 0, 1] - 1.5 - torch.rand(8) * 4     # narrow tables -> escapes happen
        q[:, 0, 2] = q[:, 0, 1] + 1.5 + torch.rand(8) * 4
    layer.eval()
    layer.update(force=True)
    torch.manual_seed(8)
    x = torch.randn(3, 3, 64, 48) * 1.5
    with torch.inference_mode():
        latent = layer.encoder(x)
        enc = layer.encode(x)
        dec = layer.decode(**enc)
        fwd = layer(x)
        assert torch.equal(fwd, dec)
        eb = layer.entropy_bottleneck
        med = eb._get_medians().detach().reshape(1, -1, 1, 1)
        symbols = torch.round(latent - med).int()
        latent_hat = eb.decompress(enc['strings'][0], enc['shape'])
        assert torch.equal(latent_hat, symbols.float() + med)
    out = {'x': x.numpy(), 'latent': latent.numpy(), 'symbols': symbols.numpy(), 'latent_hat': latent_hat.numpy(),
           'decoded': dec.numpy(), 'shape': np.array(tuple(enc['shape']))}
    out['streams'], out['stream_offsets'] = _pack_strings(enc['strings'][0])
    for k, v in layer.state_dict().items():
        out['sd/' + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, 'fp_bottleneck_small.npz'), **out)
    n_esc = int(((symbols < eb._offset.view(1, -1, 1, 1)) | (symbols - eb._offset.view(1, -1, 1, 1) >= (eb._cdf_length - 2).view(1, -1, 1, 1))).sum())
    print('fp_bottleneck_small.npz: latent', tuple(latent.shape), 'bytes', [len(s) for s in enc['strings'][0]],
          'sym range', int(symbols.min()), int(symbols.max()), 'escapes', n_esc)


def config1(torch):
    """BASELINE.json configs[0]: Entropic Student splittable ResNet-50, batch 1, 3x224x224, CPU, random init
    (SURVEY.md 8d "Config 1").  Weights are regenerated from the seed by the tests (25 M parameters are not a
    'small fixture'); a checksum of the bottleneck weights guards against RNG drift."""
    from sc2bench.models.backbone import splittable_resnet
    torch.manual_seed(0)
    model = splittable_resnet(bottleneck_config={'key': 'FPBasedResNetBottleneck',
                                                 'kwargs': {'num_bottleneck_channels': 24, 'num_target_channels': 256}},
                              resnet_name='resnet50', skips_avgpool=False, skips_fc=False, weights=None)
    model.eval()
    model.update()
    torch.manual_seed(1)
    x = torch.randn(1, 3, 224, 224)
    bl = model.bottleneck_layer
    with torch.inference_mode():
        latent = bl.encoder(x)
        enc = bl.encode(x)
        dec = bl.decode(**enc)
        logits = model(x)
        med = bl.entropy_bottleneck._get_medians().detach().reshape(1, -1, 1, 1)
        symbols = torch.round(latent - med).int()
    wsum = hashlib.sha256()
    for k, v in bl.state_dict().items():
        wsum.update(k.encode())
        wsum.update(v.numpy().tobytes())
    out = {'weights_sha256': np.array(wsum.hexdigest()), 'x_sha256': np.array(_sha(x.numpy())),
           'symbols': symbols.numpy().astype(np.int8), 'latent_sub': latent.numpy()[0, :, ::5, ::5],
           'stream': _bytes_to_u8(enc['strings'][0][0]), 'shape': np.array(tuple(enc['shape'])),
           'decoded_sub': dec.numpy()[0, ::8, ::4, ::4], 'decoded_sha256': np.array(_sha(dec.numpy())),
           'decoded_absmax': np.array(float(dec.abs().max())), 'decoded_mean': np.array(float(dec.double().mean())),
           'logits': logits.numpy()[0], 'top1': np.array(int(logits.argmax()))}
    np.savez_compressed(os.path.join(GOLD, 'config1_entropic_student_resnet50.npz'), **out)
    print('config1: stream bytes', len(enc['strings'][0][0]), 'top1', int(logits.argmax()), 'symbols nz', int((symbols != 0).sum()))


def small_zoo_models(torch):
    """FactorizedPrior / ScaleHyperprior (configs[2], configs[3]) at a tiny (N, M) with stored weights, driven through
    the reference's NeuralInputCompressionClassifier.forward contract: compress -> analyze -> decompress['x_hat']
    (sc2bench/models/wrapper.py:119-135)."""
    from compressai.models import FactorizedPrior, ScaleHyperprior
    for name, cls in (('factorized_prior_small', FactorizedPrior), ('scale_hyperprior_small', ScaleHyperprior)):
        torch.manual_seed(21)
        net = cls(16, 24)
        with torch.no_grad():
            for mod in net.modules():
                if hasattr(mod, 'gamma'):
                    C = mod.gamma.shape[0]
                    mod.gamma.copy_(mod.gamma_reparam.init(0.1 * torch.eye(C) + 0.01 * torch.rand(C, C)))
                    mod.beta.copy_(mod.beta_reparam.init(0.5 + torch.rand(C)))
                if isinstance(mod, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)):
                    mod.bias.normal_(0, 0.05)
            for mod in net.g_a:
                if isinstance(mod, torch.nn.Conv2d):
                    mod.weight.mul_(3.0)
        perturb_entropy_bottleneck(torch, net.entropy_bottleneck, 5)
        net.eval()
        net.update(force=True)
        torch.manual_seed(22)
        x = torch.rand(2, 3, 128, 64)
        with torch.inference_mode():
            obj = net.compress(x)
            x_hat = net.decompress(**obj)['x_hat']
            y = net.g_a(x)
        out = {'x': x.numpy(), 'y': y.numpy(), 'x_hat': x_hat.numpy(), 'shape': np.array(tuple(obj['shape'])),
               'n_string_lists': np.array(len(obj['strings']))}
        for li, strings in enumerate(obj['strings']):
            out['streams%d' % li], out['stream_offsets%d' % li] = _pack_strings(strings)
        for k, v in net.state_dict().items():
            out['sd/' + k] = v.numpy()
        np.savez_compressed(os.path.join(GOLD, name + '.npz'), **out)
        print(name, 'y', tuple(y.shape), 'bytes', [[len(s) for s in l] for l in obj['strings']])


def small_shp_bottleneck(torch, key='SHPBasedResNetBottleneck', fname='shp_bottleneck_small.npz', seed=31):
    """The reference's SHPBasedResNetBottleneck / MSHPBasedResNetBottleneck (sc2bench/models/layer.py:553-817) at a tiny
    size, weights stored."""
    from sc2bench.models.layer import get_layer
    torch.manual_seed(seed)
    layer = get_layer(key, num_input_channels=3, num_latent_channels=8, num_bottleneck_channels=8,
                      num_target_channels=32)
    with torch.no_grad():
        for mod in list(layer.g_a) + list(layer.g_s):
            if hasattr(mod, 'gamma'):
                C = mod.gamma.shape[0]
                mod.gamma.copy_(mod.gamma_reparam.init(0.1 * torch.eye(C) + 0.02 * torch.rand(C, C)))
                mod.beta.copy_(mod.beta_reparam.init(0.5 + torch.rand(C)))
        for mod in layer.g_a:
            if isinstance(mod, torch.nn.Conv2d):
                mod.weight.mul_(3.0)
        for mod in list(layer.h_a) + list(layer.h_s):
            if hasattr(mod, 'weight'):
                mod.weight.mul_(2.0)
    perturb_entropy_bottleneck(torch, layer.entropy_bottleneck, 13)
    layer.eval()
    layer.update(force=True)
    torch.manual_seed(seed + 1)
    x = torch.randn(2, 3, 96, 80) * 1.5
    with torch.inference_mode():
        enc = layer.encode(x)
        dec = layer.decode(**enc)
        y = layer.g_a(x)
    out = {'x': x.numpy(), 'y': y.numpy(), 'decoded': dec.numpy(), 'shape': np.array(tuple(enc['shape']))}
    for li, strings in enumerate(enc['strings']):
        out['streams%d' % li], out['stream_offsets%d' % li] = _pack_strings(strings)
    for k, v in layer.state_dict().items():
        out['sd/' + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, fname), **out)
    print(fname, ': y', tuple(y.shape), 'z shape', tuple(enc['shape']), 'bytes', [[len(s) for s in l] for l in enc['strings']])


def small_mshp_bottleneck(torch):
    small_shp_bottleneck(torch, 'MSHPBasedResNetBottleneck', 'mshp_bottleneck_small.npz', seed=41)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--real-compressai', action='store_true')
    ap.add_argument('--only', default=None)
    args = ap.parse_args()
    _setup_paths(args.real_compressai)
    warnings.simplefilter('ignore')
    import torch
    torch.set_num_threads(1)  # bit-reproducible reductions
    os.makedirs(GOLD, exist_ok=True)
    jobs = {'rans': rans_cases, 'small_fp': small_fp_bottleneck, 'config1': config1, 'zoo': small_zoo_models, 'shp': small_shp_bottleneck,
            'mshp': small_mshp_bottleneck}
    for k, fn in jobs.items():
        if args.only in (None, k):
            fn(torch)


if __name__ == '__main__':
    main()
