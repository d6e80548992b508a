"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the committed golden vectors.

Bars (BASELINE.json north_star): bitstreams byte-identical given identical symbols; symbol mismatches < 1e-6 of
symbols; latents within 1e-4 relative; decoded features / logits within 1e-3 relative with identical top-1.
"Relative" is max|a - b| / max|b| over the tensor.
"""
import hashlib

import numpy as np
import pytest
import torch

import cref
from helpers import load_golden, rel_err, state_dict_from_golden, unpack_streams

pytestmark = pytest.mark.gpu

LATENT_TOL = 1e-4
FEATURE_TOL = 1e-3


@pytest.fixture(scope='module')
def s2():
    import sc2bench_b200
    return sc2bench_b200


@pytest.fixture(scope='module')
def dev():
    return torch.device('cuda:0')


@pytest.fixture(scope='module')
def g():
    return load_golden('rans_cases.npz')


def _tables(s2, cdf, ln, off):
    return s2.ops.CoderTables(torch.from_numpy(np.asarray(cdf)), torch.from_numpy(np.asarray(ln)), torch.from_numpy(np.asarray(off)))


# ---------------------------------------------------------------------------------------------------
# coder
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['eb24_sigma1', 'eb24_sigma3', 'eb24_sigma8', 'single', 'single_escape', 'edge_values',
                                  'huge_escapes', 'gc_mixed'])
def test_rans_golden_streams_explicit_indexes(s2, dev, g, name):
    if name.startswith('gc'):
        gc = s2.GaussianConditional(None)
        gc.update_scale_table(s2.get_scale_table())
        tables = gc.coder_tables()
    else:
        tables = _tables(s2, g['eb24_cdf'], g['eb24_len'], g['eb24_off'])
    sym = torch.from_numpy(g[name + '_symbols']).to(dev).view(1, -1)
    idx = torch.from_numpy(g[name + '_indexes']).to(dev).view(1, -1)
    streams = s2.ops.rans_encode(sym, tables, indexes=idx)
    got = streams.tolist()
    assert got[0] == g[name + '_stream'].tobytes()
    back = s2.ops.rans_decode(s2.ops.PackedStreams.from_list(got, dev), sym.shape[1], tables, indexes=idx, want='symbols')
    assert torch.equal(back, sym)


@pytest.mark.parametrize('layout', ['warp', 'lanes'])
def test_rans_empty_batch_and_empty_stream(s2, dev, g, layout):
    tables = _tables(s2, g['eb24_cdf'], g['eb24_len'], g['eb24_off'])
    sym = torch.zeros((3, 0), dtype=torch.int32, device=dev)
    out = s2.ops.rans_encode(sym, tables, spatial=1, layout=layout).tolist()
    assert out == [g['empty_stream'].tobytes()] * 3
    assert s2.ops.rans_encode(torch.zeros((0, 5), dtype=torch.int32, device=dev), tables, spatial=5).tolist() == []


@pytest.mark.parametrize('layout', ['warp', 'lanes'])
def test_rans_config1_golden_channel_mode(s2, dev, g, layout):
    c1 = load_golden('config1_entropic_student_resnet50.npz')
    tables = _tables(s2, g['eb24_cdf'], g['eb24_len'], g['eb24_off'])
    sym = torch.from_numpy(c1['symbols'].astype(np.int32)).to(dev)
    streams = s2.ops.rans_encode(sym, tables, spatial=55 * 55, layout=layout)
    got = streams.tolist()
    assert len(got) == 1 and got[0] == c1['stream'].tobytes()
    back = s2.ops.rans_decode(streams, 24 * 55 * 55, tables, spatial=55 * 55, layout=layout, want='symbols')
    assert torch.equal(back.view_as(sym), sym)


@pytest.mark.parametrize('layout', ['warp', 'lanes'])
def test_rans_full_size_batch_roundtrip_and_oracle_sample(s2, dev, g, layout):
    """BASELINE configs[1] size: 256 streams x 72,600 symbols, random symbols including escapes."""
    tables = _tables(s2, g['eb24_cdf'], g['eb24_len'], g['eb24_off'])
    gen = torch.Generator(device='cpu').manual_seed(5)
    sigma = torch.rand(256, 24, 1, 1, generator=gen) * 6 + 0.2
    sym = torch.round(torch.randn(256, 24, 55, 55, generator=gen) * sigma).int()
    sym[3, 2, 10, 10] = 100000
    sym[3, 2, 10, 11] = -100000
    sym_d = sym.to(dev)
    streams = s2.ops.rans_encode(sym_d, tables, spatial=55 * 55, layout=layout)
    back = s2.ops.rans_decode(streams, 24 * 55 * 55, tables, spatial=55 * 55, layout=layout, want='symbols')
    assert torch.equal(back.view_as(sym_d), sym_d)
    med = torch.linspace(-1, 1, 24, device=dev)
    vals = s2.ops.rans_decode(streams, 24 * 55 * 55, tables, spatial=55 * 55, layout=layout, means=med, want='values')
    assert torch.equal(vals.view_as(sym_d), sym_d.float() + med.view(1, 24, 1, 1))
    strings = streams.tolist()
    idx = np.repeat(np.arange(24, dtype=np.int32), 55 * 55)
    for b in (0, 3, 77, 255):
        assert strings[b] == cref.encode_with_indexes(sym[b].reshape(-1).numpy(), idx, g['eb24_cdf'], g['eb24_len'], g['eb24_off'])
    assert all(len(s) % 4 == 0 and len(s) >= 8 for s in strings)
    # host round trip of the contract object: list[bytes] -> device -> symbols
    again = s2.ops.rans_decode(s2.ops.PackedStreams.from_list(strings, dev), 24 * 55 * 55, tables, spatial=55 * 55, layout=layout, want='symbols')
    assert torch.equal(again.view_as(sym_d), sym_d)


@pytest.mark.parametrize('layout', ['warp', 'lanes'])
def test_rans_wide_tables_and_ragged_rows(s2, dev, oracle_compressai, layout):
    """Tables wider than one / two warps (trained-like EntropyBottleneck) and all 64 GaussianConditional rows."""
    from compressai.entropy_models import EntropyBottleneck as OracleEB
    torch.manual_seed(2)
    eb = OracleEB(40)
    with torch.no_grad():
        eb.quantiles[:, 0, 0] = -torch.linspace(2, 90, 40)
        eb.quantiles[:, 0, 2] = torch.linspace(3, 70, 40)
        for m in eb.matrices:
            m.sub_(1.2)  # flatten the density so that wide tables stay strictly positive
    eb.update(force=True)
    cdf, ln, off = eb._quantized_cdf.numpy(), eb._cdf_length.numpy(), eb._offset.numpy()
    assert ln.max() > 130 and ln.min() < 20
    tables = _tables(s2, cdf, ln, off)
    rng = np.random.RandomState(9)
    sym = np.round(rng.randn(5, 40, 7, 9) * np.linspace(1, 60, 40).reshape(1, 40, 1, 1)).astype(np.int32)
    sym_d = torch.from_numpy(sym).to(dev)
    strings = s2.ops.rans_encode(sym_d, tables, spatial=63, layout=layout).tolist()
    idx = np.repeat(np.arange(40, dtype=np.int32), 63)
    for b in range(5):
        assert strings[b] == cref.encode_with_indexes(sym[b].reshape(-1), idx, cdf, ln, off)
    back = s2.ops.rans_decode(s2.ops.PackedStreams.from_list(strings, dev), 40 * 63, tables, spatial=63, layout=layout, want='symbols')
    assert torch.equal(back.view_as(sym_d), sym_d)
    # GaussianConditional: explicit indexes, every row, tables up to 3133 entries (not staged in shared memory)
    gc = s2.GaussianConditional(None)
    gc.update_scale_table(s2.get_scale_table())
    gcdf, gln, goff = gc._quantized_cdf.numpy(), gc._cdf_length.numpy(), gc._offset.numpy()
    gidx = rng.randint(0, 64, size=(3, 4000)).astype(np.int32)
    gsym = np.round(rng.randn(3, 4000) * gc.scale_table.numpy()[gidx] * 1.3).astype(np.int32)
    strings = s2.ops.rans_encode(torch.from_numpy(gsym).to(dev), gc.coder_tables(), indexes=torch.from_numpy(gidx).to(dev)).tolist()
    for b in range(3):
        assert strings[b] == cref.encode_with_indexes(gsym[b], gidx[b], gcdf, gln, goff)
    back = s2.ops.rans_decode(s2.ops.PackedStreams.from_list(strings, dev), 4000, gc.coder_tables(),
                              indexes=torch.from_numpy(gidx).to(dev), want='symbols')
    assert (back.cpu().numpy() == gsym).all()


def test_rans_rejects_malformed_streams(s2, dev, g):
    tables = _tables(s2, g['eb24_cdf'], g['eb24_len'], g['eb24_off'])
    with pytest.raises(ValueError):
        s2.ops.PackedStreams.from_list([b'\x00' * 7], dev)
    good = g['eb24_sigma3_stream'].tobytes()
    idx = torch.from_numpy(g['eb24_sigma3_indexes']).to(dev).view(1, -1)
    with pytest.raises(ValueError, match='truncated'):
        s2.ops.rans_decode(s2.ops.PackedStreams.from_list([good[:16]], dev), idx.shape[1], tables, indexes=idx, want='symbols')


def test_quantize_rounds_half_to_even_like_torch(s2, dev):
    x = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 0.49999997, 1e-8, 7.5000005, -3.4999998], device=dev).view(1, 1, -1)
    med = torch.tensor([0.25], device=dev)
    got = s2.ops.quantize_symbols(x, med)
    assert torch.equal(got, torch.round(x - med.view(1, 1, 1)).int())
    big = torch.randn(4, 6, 33, 17, device=dev) * 5
    meds = torch.randn(6, device=dev)
    assert torch.equal(s2.ops.quantize_symbols(big, meds), torch.round(big - meds.view(1, 6, 1, 1)).int())


def test_gc_build_indexes_matches_reference_loop(s2, dev, oracle_compressai):
    from compressai.entropy_models import GaussianConditional as OracleGC
    from compressai.models import get_scale_table
    ref = OracleGC(None)
    ref.update_scale_table(get_scale_table())
    gc = s2.GaussianConditional(None)
    gc.update_scale_table(s2.get_scale_table())
    torch.manual_seed(0)
    scales = torch.cat([torch.rand(5000) * 300, ref.scale_table.clone(), ref.scale_table * (1 + 1e-7), torch.tensor([0.0, 0.05, 0.11, 1e9])])
    assert torch.equal(gc.build_indexes(scales.to(dev)).cpu(), ref.build_indexes(scales))


# ---------------------------------------------------------------------------------------------------
# transforms (fp32 path)
# ---------------------------------------------------------------------------------------------------
CONV_CASES = [  # (c_in, c_out, k, stride, pad, transposed, out_pad, H, W, bias)
    (3, 96, 5, 2, 2, False, 0, 64, 48, False),
    (96, 48, 5, 2, 2, False, 0, 33, 29, False),
    (48, 24, 2, 1, 0, False, 0, 17, 13, False),
    (24, 512, 2, 1, 1, False, 0, 15, 11, False),
    (512, 256, 2, 1, 0, False, 0, 16, 12, False),
    (256, 256, 2, 1, 1, False, 0, 15, 11, False),
    (3, 20, 5, 2, 2, False, 0, 37, 41, True),
    (24, 16, 5, 2, 2, True, 1, 9, 7, True),
    (16, 3, 5, 2, 2, True, 1, 18, 14, True),
    (16, 16, 5, 2, 1, True, 0, 6, 5, True),
    (24, 16, 3, 1, 1, False, 0, 8, 4, True),
]


@pytest.mark.parametrize('case', CONV_CASES)
def test_conv2d_f32_matches_torch(s2, dev, case):
    cin, cout, k, stride, pad, tr, op, H, W, has_bias = case
    torch.manual_seed(hash(case) % 1000)
    x = torch.randn(3, cin, H, W)
    if tr:
        w = torch.randn(cin, cout, k, k) / (cin * k * k) ** 0.5
    else:
        w = torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5
    b = torch.randn(cout) if has_bias else None
    ref = torch.nn.functional.conv_transpose2d(x, w, b, stride, pad, op) if tr else torch.nn.functional.conv2d(x, w, b, stride, pad)
    got = s2.ops.conv2d(x.to(dev), w.to(dev), b.to(dev) if has_bias else None, stride=stride, padding=pad, transposed=tr, output_padding=op)
    assert got.shape == ref.shape
    assert rel_err(got.cpu(), ref) < 5e-6
    relu = s2.ops.conv2d(x.to(dev), w.to(dev), b.to(dev) if has_bias else None, stride=stride, padding=pad, transposed=tr,
                         output_padding=op, epilogue=s2._native.EPI_RELU)
    assert torch.equal(relu, torch.relu(got))


@pytest.mark.parametrize('C,kind,inverse', [(96, 0, False), (48, 0, False), (512, 0, True), (256, 0, True), (20, 1, False), (20, 1, True)])
def test_gdn_matches_reference(s2, dev, oracle_compressai, C, kind, inverse):
    import compressai.layers as L
    torch.manual_seed(C + kind)
    ref = (L.GDN1 if kind == 0 else L.GDN)(C, inverse=inverse)
    with torch.no_grad():
        ref.gamma.copy_(ref.gamma_reparam.init(0.1 * torch.eye(C) + 0.02 * torch.rand(C, C)))
        ref.beta.copy_(ref.beta_reparam.init(0.5 + torch.rand(C)))
    mine = (s2.GDN1 if kind == 0 else s2.GDN)(C, inverse=inverse)
    mine.load_state_dict(ref.state_dict())
    mine.to(dev)
    x = torch.randn(2, C, 9, 13)
    with torch.no_grad():
        want = ref(x)
        got = mine(x.to(dev))
    assert rel_err(got.cpu(), want) < 3e-6


# ---------------------------------------------------------------------------------------------------
# whole path against the golden vectors produced by the reference's own model code
# ---------------------------------------------------------------------------------------------------
def test_fp_bottleneck_small_golden(s2, dev):
    gs = load_golden('fp_bottleneck_small.npz')
    layer = s2.get_layer('FPBasedResNetBottleneck', num_input_channels=3, num_bottleneck_channels=8, num_target_channels=32)
    layer.load_state_dict(state_dict_from_golden(gs))
    layer.update()
    layer.eval().to(dev)
    x = torch.from_numpy(gs['x']).to(dev)
    enc = layer.encode(x)
    assert tuple(enc['shape']) == tuple(gs['shape'])
    assert isinstance(enc['strings'], list) and len(enc['strings']) == 1 and all(isinstance(s, bytes) for s in enc['strings'][0])
    want_strings = unpack_streams(gs['streams'], gs['stream_offsets'])
    # symbols first: with identical symbols the streams must be byte-identical
    latent = s2.models.run_transform(layer.encoder, x)
    assert rel_err(latent.cpu(), torch.from_numpy(gs['latent'])) < LATENT_TOL
    symbols = s2.ops.quantize_symbols(latent, layer.entropy_bottleneck._get_medians().detach().reshape(-1))
    mismatches = int((symbols.cpu() != torch.from_numpy(gs['symbols'])).sum())
    assert mismatches == 0, 'symbol mismatches vs oracle: %d of %d' % (mismatches, symbols.numel())
    assert enc['strings'][0] == want_strings
    dec = layer.decode(**enc)
    assert rel_err(dec.cpu(), torch.from_numpy(gs['decoded'])) < FEATURE_TOL
    # decoding the GOLDEN bytes gives the golden latent exactly (integer + median)
    latent_hat = layer.entropy_bottleneck.decompress(want_strings, tuple(gs['shape']))
    assert torch.equal(latent_hat.cpu(), torch.from_numpy(gs['latent_hat']))
    assert torch.equal(layer(x), dec)


def test_config1_entropic_student_resnet50(s2, dev):
    """BASELINE configs[0]: splittable ResNet-50, batch 1, 3x224x224, random init (seeds as SURVEY.md 8d)."""
    c1 = load_golden('config1_entropic_student_resnet50.npz')
    torch.manual_seed(0)
    model = s2.splittable_resnet(bottleneck_config={'key': 'FPBasedResNetBottleneck',
                                                    'kwargs': {'num_bottleneck_channels': 24, 'num_target_channels': 256}},
                                 resnet_name='resnet50', skips_avgpool=False, skips_fc=False, weights=None)
    model.eval()
    model.update()
    h = hashlib.sha256()
    for k, v in model.bottleneck_layer.state_dict().items():
        h.update(k.encode())
        h.update(v.numpy().tobytes())
    if h.hexdigest() != str(c1['weights_sha256']):
        pytest.skip('torch RNG stream differs from the one the golden vectors were generated with')
    torch.manual_seed(1)
    x = torch.randn(1, 3, 224, 224)
    assert hashlib.sha256(x.numpy().tobytes()).hexdigest() == str(c1['x_sha256'])
    model.to(dev)
    xd = x.to(dev)
    bl = model.bottleneck_layer
    with torch.inference_mode():
        enc = bl.encode(xd)
        streams, _ = bl.encode_packed(xd)
        symbols = s2.ops.rans_decode(streams, 24 * 55 * 55, bl.entropy_bottleneck.coder_tables(), spatial=55 * 55, want='symbols')
        mism = int((symbols.view(1, 24, 55, 55).cpu() != torch.from_numpy(c1['symbols'].astype(np.int32))).sum())
        assert mism == 0, 'symbol mismatches: %d of 72600' % mism
        assert enc['strings'][0][0] == c1['stream'].tobytes()
        assert tuple(enc['shape']) == (55, 55)
        dec = bl.decode(**enc)
        assert dec.shape == (1, 256, 56, 56)
        assert rel_err(dec[0, ::8, ::4, ::4].cpu(), torch.from_numpy(c1['decoded_sub'])) < FEATURE_TOL
        assert abs(float(dec.abs().max()) - float(c1['decoded_absmax'])) < FEATURE_TOL * float(c1['decoded_absmax'])
        logits = model(xd)
        assert rel_err(logits[0].cpu(), torch.from_numpy(c1['logits'])) < FEATURE_TOL
        assert int(logits.argmax()) == int(c1['top1'])


@pytest.mark.parametrize('name', ['factorized_prior_small', 'scale_hyperprior_small'])
def test_zoo_models_small_golden(s2, dev, name):
    gz = load_golden(name + '.npz')
    cls = s2.FactorizedPrior if name.startswith('factorized') else s2.ScaleHyperprior
    net = cls(16, 24)
    net.load_state_dict(state_dict_from_golden(gz))
    net.update()
    net.eval().to(dev)
    x = torch.from_numpy(gz['x']).to(dev)
    y = s2.models.run_transform(net.g_a, x)
    assert rel_err(y.cpu(), torch.from_numpy(gz['y'])) < LATENT_TOL
    obj = net.compress(x)
    assert len(obj['strings']) == int(gz['n_string_lists']) and tuple(obj['shape']) == tuple(gz['shape'])
    for li in range(len(obj['strings'])):
        assert obj['strings'][li] == unpack_streams(gz['streams%d' % li], gz['stream_offsets%d' % li]), 'string list %d' % li
    x_hat = net.decompress(**obj)['x_hat']
    assert rel_err(x_hat.cpu(), torch.from_numpy(gz['x_hat'])) < FEATURE_TOL
    assert float(x_hat.min()) >= 0.0 and float(x_hat.max()) <= 1.0


def test_batch256_path_properties(s2, dev):
    """Full-size (BASELINE configs[1]) size-independent properties: decode(encode(x)) == g_s(dequant(symbols)),
    streams decode back to the symbols that produced them, per-sample independence of the batch."""
    torch.manual_seed(0)
    layer = s2.get_layer('FPBasedResNetBottleneck').eval()
    layer.update()
    layer.to(dev)
    torch.manual_seed(1)
    x = torch.randn(256, 3, 224, 224, device=dev)
    with torch.inference_mode():
        streams, shape = layer.encode_packed(x)
        assert tuple(shape) == (55, 55)
        eb = layer.entropy_bottleneck
        med = eb._get_medians().detach().reshape(-1)
        symbols = layer.analyze_to_symbols(x)
        # the fp32-grade tensor-core g_a against the exact-fp32 CUDA-core g_a: symbol mismatches stay below 1e-6
        layer.encoder_precision = 'fp32'
        exact_symbols = layer.analyze_to_symbols(x[:64])
        layer.encoder_precision = 'split-tc'
        mism = int((symbols[:64] != exact_symbols).sum())
        assert mism <= max(1, exact_symbols.numel() // 1000000), 'symbol mismatches tensor-core vs fp32: %d of %d' % (mism, exact_symbols.numel())
        back = s2.ops.rans_decode(streams, 24 * 55 * 55, eb.coder_tables(), spatial=55 * 55, want='symbols')
        assert torch.equal(back.view_as(symbols), symbols)
        out = layer.decode_packed(streams, shape)
        assert out.shape == (256, 256, 56, 56)
        direct = layer.synthesize(symbols.float() + med.view(1, -1, 1, 1))
        assert torch.equal(out, direct)
        # the tensor-core g_s (fp16 operands) against the exact-fp32 CUDA-core g_s at full size
        layer.decoder_precision = 'fp32'
        exact = layer.synthesize(symbols[:32].float() + med.view(1, -1, 1, 1))
        layer.decoder_precision = 'fp16-tc'
        assert rel_err(out[:32], exact) < FEATURE_TOL
        # sample 17 alone gives the same bytes and features as inside the batch
        s17, _ = layer.encode_packed(x[17:18])
        assert s17.tolist()[0] == streams.tolist()[17]
        assert torch.equal(layer.decode_packed(s17, shape)[0], out[17])
        total_bytes = streams.total_bytes()
        assert 256 * 8 < total_bytes < 256 * 24 * 55 * 55 * 2


@pytest.mark.parametrize('key,fname', [('SHPBasedResNetBottleneck', 'shp_bottleneck_small.npz'),
                                       ('MSHPBasedResNetBottleneck', 'mshp_bottleneck_small.npz')])
def test_shp_bottleneck_small_golden(s2, dev, key, fname):
    """golden vectors from the reference's own SHPBasedResNetBottleneck / MSHPBasedResNetBottleneck
    (layer.py:553-817)."""
    gs = load_golden(fname)
    layer = s2.get_layer(key, num_input_channels=3, num_latent_channels=8, num_bottleneck_channels=8,
                         num_target_channels=32)
    layer.load_state_dict(state_dict_from_golden(gs))
    layer.update()
    layer.eval().to(dev)
    x = torch.from_numpy(gs['x']).to(dev)
    y = s2.models.run_transform(layer.g_a, x)
    assert rel_err(y.cpu(), torch.from_numpy(gs['y'])) < LATENT_TOL
    # the route encode() takes: the fused tensor-core analysis kernels
    assert s2.bottleneck.TensorCoreAnalysis.why_not(layer.g_a, x.shape) is None
    y_tc = layer._analysis(x)
    assert layer.__dict__.get('_tc_encoder') is not None
    assert rel_err(y_tc.cpu(), torch.from_numpy(gs['y'])) < LATENT_TOL
    enc = layer.encode(x)
    # ... and h_a / h_s on the split tensor-core plan (odd sizes, channel counts that are not multiples of 16, k5 s2 p1 transposed convs)
    assert set(layer.__dict__['_tc_analysis']) == {'h_a', 'h_s'}
    assert tuple(enc['shape']) == tuple(gs['shape']) and len(enc['strings']) == 2
    want_y = unpack_streams(gs['streams0'], gs['stream_offsets0'])
    want_z = unpack_streams(gs['streams1'], gs['stream_offsets1'])
    assert enc['strings'][1] == want_z, 'z streams differ'
    assert enc['strings'][0] == want_y, 'y streams differ'
    dec = layer.decode(**enc)
    assert rel_err(dec.cpu(), torch.from_numpy(gs['decoded'])) < FEATURE_TOL
    # cross decode: the golden bytes through our decoder
    dec2 = layer.decode([want_y, want_z], tuple(gs['shape']))
    assert rel_err(dec2.cpu(), torch.from_numpy(gs['decoded'])) < FEATURE_TOL


def test_conv_leaky_relu_abs_and_dequantize(s2, dev):
    """LeakyReLU epilogue / |x| on load (h_a, h_s of the hyperprior bottlenecks, layer.py:606-617,755-770) and the
    per-element dequantise (layer.py:785) against torch on the CPU."""
    torch.manual_seed(5)
    x = torch.randn(2, 6, 13, 11)
    conv = torch.nn.Conv2d(6, 10, 5, stride=2, padding=1, bias=False)
    tconv = torch.nn.ConvTranspose2d(6, 9, 5, stride=2, padding=1, bias=False)
    with torch.no_grad():
        want = torch.nn.functional.leaky_relu(conv(torch.abs(x)), 0.01)
        want_t = torch.nn.functional.leaky_relu(tconv(x), 0.2)
    got = s2.ops.conv2d(x.to(dev), conv.weight.to(dev), stride=2, padding=1, epilogue=s2._native.EPI_LEAKY_RELU,
                        epi_param=0.01, in_abs=True)
    got_t = s2.ops.conv2d(x.to(dev), tconv.weight.to(dev), stride=2, padding=1, transposed=True,
                          epilogue=s2._native.EPI_LEAKY_RELU, epi_param=0.2)
    assert rel_err(got.cpu(), want) < 5e-6 and rel_err(got_t.cpu(), want_t) < 5e-6
    seq = torch.nn.Sequential(conv, torch.nn.LeakyReLU(inplace=True)).to(dev)
    assert torch.equal(s2.models.run_transform(seq, x.to(dev), in_abs=True), got)
    sym = torch.randint(-40, 40, (3, 5, 7), dtype=torch.int32)
    means = torch.randn(3, 5, 7)
    assert torch.equal(s2.ops.dequantize(sym.to(dev), means.to(dev)).cpu(), sym.float() + means)
    assert torch.equal(s2.ops.dequantize(sym.to(dev)).cpu(), sym.float())
    y = torch.randn(3, 5, 7) * 4
    assert torch.equal(s2.ops.quantize_symbols(y.to(dev), means.to(dev)).cpu(), torch.round(y - means).int())


def test_rans_lanes_layout_edge_cases(s2, dev, g, oracle_compressai):
    """Lane-per-stream layout (rans_lanes.cu): stream counts that are not a multiple of 32 or of the block size, stream
    lengths that are not a multiple of 4 or 8 (scalar head / tail of the vector stores, partial last block), CDF rows with
    more than 255 symbols (the decoder LUT's 8-bit symbol index saturates) and rows shorter than one LUT bucket step, and
    malformed / truncated streams.  Always byte-identical to the warp-per-stream layout and to the oracle."""
    from compressai.entropy_models import EntropyBottleneck as OracleEB
    torch.manual_seed(4)
    eb = OracleEB(6)
    with torch.no_grad():
        eb.quantiles[:, 0, 0] = -torch.tensor([2., 40., 160., 300., 5., 900.])
        eb.quantiles[:, 0, 2] = torch.tensor([3., 50., 170., 280., 4., 800.])
        for m in eb.matrices:
            m.sub_(2.0)
    eb.update(force=True)
    cdf, ln, off = eb._quantized_cdf.numpy(), eb._cdf_length.numpy(), eb._offset.numpy()
    assert ln.max() > 600 and ln.min() < 16
    tables = _tables(s2, cdf, ln, off)
    rng = np.random.RandomState(11)
    scale = np.array([1, 15, 60, 110, 2, 300], dtype=np.float64).reshape(1, 6, 1)
    for B, spatial in ((1, 7), (33, 13), (70, 101), (300, 9)):
        sym = np.round(rng.randn(B, 6, spatial) * scale).astype(np.int32)
        sym[0, 0, 0] = 5000       # escapes on both sides
        sym[B - 1, 5, spatial - 1] = -70000
        sym_d = torch.from_numpy(sym).to(dev)
        lanes = s2.ops.rans_encode(sym_d, tables, spatial=spatial, layout='lanes').tolist()
        warp = s2.ops.rans_encode(sym_d, tables, spatial=spatial, layout='warp').tolist()
        assert lanes == warp
        idx = np.repeat(np.arange(6, dtype=np.int32), spatial)
        for b in sorted({0, B // 2, B - 1}):
            assert lanes[b] == cref.encode_with_indexes(sym[b].reshape(-1), idx, cdf, ln, off)
        ps = s2.ops.PackedStreams.from_list(lanes, dev)
        back = s2.ops.rans_decode(ps, 6 * spatial, tables, spatial=spatial, want='symbols', layout='lanes')
        assert torch.equal(back.view_as(sym_d), sym_d)
        med = torch.linspace(-2, 2, 6, device=dev)
        vals = s2.ops.rans_decode(ps, 6 * spatial, tables, spatial=spatial, means=med, want='values', layout='lanes')
        assert torch.equal(vals.view_as(sym_d), sym_d.float() + med.view(1, 6, 1))
    # a truncated stream among good ones is reported, and does not disturb its neighbours' lanes
    good = lanes[:40]
    bad = list(good)
    bad[7] = bad[7][:8]
    with pytest.raises(ValueError, match='Invalid bitstream'):
        s2.ops.rans_decode(s2.ops.PackedStreams.from_list(bad, dev), 6 * spatial, tables, spatial=spatial, want='symbols', layout='lanes')
    out, st = s2.ops.rans_decode(s2.ops.PackedStreams.from_list(bad, dev), 6 * spatial, tables, spatial=spatial, want='symbols',
                                 layout='lanes', check_status=False, return_status=True)
    assert int(st.item()) != 0
    keep = [i for i in range(40) if i != 7]
    assert torch.equal(out.view(40, 6, spatial)[keep], sym_d[:40][keep])
