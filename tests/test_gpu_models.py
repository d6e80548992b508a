"""GPU: the callers either side of the path at the shapes BASELINE.json names (configs[2..4]): the input-compression
wrapper with the CompressAI zoo models at quality 8 (N=192, M=320; 224 -> AdaptivePad(64) -> 256), and the detection-style
FeatureExtractionBackbone at COCO shape (3x800x1344, one 1.6 M-symbol stream per image).  Checked against the oracle
(restated CompressAI on CPU torch + C rANS) on the same seeded weights and inputs."""
import numpy as np
import pytest
import torch
from torch import nn

import cref
from helpers import rel_err

pytestmark = pytest.mark.gpu
FEATURE_TOL = 1e-3   # decoded features / x_hat, max-abs over max-abs (north_star)
LATENT_TOL = 1e-4    # analysis-transform latents, max-abs over max-abs (north_star)


def symbol_budget(n_symbols):
    """north_star: symbol mismatches from fp32 rounding ties must stay below 1e-6 of the symbols.  For samples smaller than
    10^6 symbols that is a budget of less than one symbol; one is allowed (a tie is a property of the input, not of the size)."""
    return max(1, int(np.ceil(1e-6 * n_symbols)))


def _torch_seq(seq, x):
    """Runs an nn.Sequential of this package layer by layer with plain torch CPU ops (the product's GDN modules compute with
    the differentiable torch branch when gradients are enabled; here the formula is applied directly)."""
    import torch.nn.functional as F
    for m in seq:
        if hasattr(m, 'effective_params'):
            gamma, beta = m.effective_params()
            C = beta.numel()
            norm = F.conv2d(m._norm_input(x), gamma.reshape(C, C, 1, 1), beta)
            x = m._apply_norm(x, norm)
        else:
            x = m(x)
    return x


@pytest.fixture(scope='module')
def s2():
    import sc2bench_b200
    return sc2bench_b200


def _decode_symbols(strings, cdf, ln, off, idx):
    return np.stack([cref.decode_with_indexes(s, idx, cdf, ln, off) for s in strings])


@pytest.mark.parametrize('arch', ['bmshj2018_factorized', 'bmshj2018_hyperprior'])
def test_neural_input_compression_classifier_q8(s2, oracle_compressai, arch):
    import compressai.zoo as ozoo
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    ref = getattr(ozoo, arch)(quality=8, pretrained=False).eval()
    ref.update()
    codec = s2.get_compression_model({'key': arch, 'kwargs': {'quality': 8, 'pretrained': False}}, 'cpu')
    codec.load_state_dict(ref.state_dict())
    codec.update()
    assert torch.equal(codec.entropy_bottleneck._quantized_cdf, ref.entropy_bottleneck._quantized_cdf)
    classifier = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(3, 10))
    model = s2.NeuralInputCompressionClassifier(classifier, pre_transform=s2.AdaptivePad(fill=0, factor=64), compression_model=codec,
                                                post_transform=None,
                                                analysis_config={'analyzes_after_compress': True,
                                                                 'analyzer_configs': [{'key': 'FileSizeAnalyzer', 'kwargs': {'unit': 'KB'}}]})
    model.eval().to(dev)
    model.activate_analysis()
    torch.manual_seed(1)
    x = torch.rand(2, 3, 224, 224)
    with torch.inference_mode():
        xp = s2.AdaptivePad(fill=0, factor=64)(x)
        assert xp.shape == (2, 3, 256, 256)
        want_obj = ref.compress(xp)
        want = ref.decompress(**want_obj)['x_hat']
        got_obj = codec.compress(xp.to(dev))
        got = codec.decompress(**got_obj)['x_hat']
        logits = model(x.to(dev))
    assert logits.shape == (2, 10) and len(model.analyzers[0].file_size_list) == 1
    assert len(got_obj['strings']) == len(want_obj['strings']) and tuple(got_obj['shape']) == tuple(want_obj['shape'])
    # (1) analysis-transform latents: 1e-4 relative (north_star), whatever the symbols do
    with torch.inference_mode():
        y_got = s2.models.run_transform(codec.g_a, xp.to(dev)).cpu()
        y_want = ref.g_a(xp)
    assert rel_err(y_got, y_want) < LATENT_TOL
    # ... and on the route compress() takes: the split tensor-core plan (no fp32 CUDA-core fallback for g_a / h_a)
    assert set(codec.__dict__['_tc_analysis']) == ({'g_a'} if arch == 'bmshj2018_factorized' else {'g_a', 'h_a', 'h_s'})
    y_tc = s2.models.run_analysis(codec, 'g_a', codec.g_a, xp.to(dev)).cpu()
    assert rel_err(y_tc, y_want) < LATENT_TOL
    # (2) the last string list is the EntropyBottleneck stream (y for factorized, z for the hyperprior): decode both sides'
    # bytes with the oracle coder; symbol mismatches (fp32 rounding ties) have a budget of 1e-6 of the symbols, at least one
    eb = ref.entropy_bottleneck
    cdf, ln, off = eb._quantized_cdf.numpy(), eb._cdf_length.numpy(), eb._offset.numpy()
    C = cdf.shape[0]
    hw = int(np.prod(tuple(want_obj['shape'])))
    idx = np.repeat(np.arange(C, dtype=np.int32), hw)
    s_want = _decode_symbols(want_obj['strings'][-1], cdf, ln, off, idx)
    s_got = _decode_symbols(got_obj['strings'][-1], cdf, ln, off, idx)
    mism = int((s_want != s_got).sum())
    assert mism <= symbol_budget(s_want.size), 'symbol mismatches vs oracle: %d of %d' % (mism, s_want.size)
    # (3) the coder is bit-exact on ITS symbols, always: the oracle coder re-encodes the decoded symbols to the same bytes
    for b, stream in enumerate(got_obj['strings'][-1]):
        assert stream == cref.encode_with_indexes(s_got[b], idx, cdf, ln, off)
    # (4) decoder parity on this implementation's own streams: the ORACLE decodes them (for the hyperprior that includes
    # h_s + build_indexes + the Gaussian-conditional decode of y with the oracle's own indexes) ...
    with torch.inference_mode():
        oracle_on_got = ref.decompress(**got_obj)['x_hat']
        cross = codec.decompress(**want_obj)['x_hat']
    assert rel_err(got.cpu(), oracle_on_got) < FEATURE_TOL
    # (5) ... and this implementation decodes the ORACLE's bytes (cross-implementation decode)
    assert rel_err(cross.cpu(), want) < FEATURE_TOL
    assert float(got.min()) >= 0 and float(got.max()) <= 1
    # (6) the device-resident calls (bitstreams stay PackedStreams, nothing synchronises) give the same bytes and the same x_hat,
    # also when two batches are issued on different CUDA streams
    with torch.inference_mode():
        outs = []
        for st in (torch.cuda.Stream(dev), torch.cuda.Stream(dev)):
            st.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(st):
                strs, shp = codec.compress_packed(xp.to(dev))
                strs = list(strs) if isinstance(strs, tuple) else [strs]
                outs.append((strs, shp, codec.decompress(strs, shp)['x_hat']))
        torch.cuda.synchronize(dev)
        codec.entropy_bottleneck.check_faults()
    for strs, shp, x_hat in outs:
        assert [ps.tolist() for ps in strs] == got_obj['strings'] and tuple(shp) == tuple(got_obj['shape'])
        assert torch.equal(x_hat, got)


def test_feature_extraction_backbone_coco_shape(s2):
    """configs[4]: splittable ResNet-50 body for Faster R-CNN at 3x800x1344 (GeneralizedRCNNTransform batching):
    latent 24x199x335 = 1,599,960 symbols in ONE stream per image."""
    import ref_models
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    body = s2.splittable_resnet(bottleneck_config={'key': 'FPBasedResNetBottleneck', 'kwargs': {'num_bottleneck_channels': 24, 'num_target_channels': 256}},
                                resnet_name='resnet50', skips_avgpool=True, skips_fc=True, weights=None)
    fx = s2.FeatureExtractionBackbone(body, {'bottleneck_layer': '1', 'layer2': '2', 'layer3': '3', 'layer4': '4'},
                                      [{'key': 'FileSizeAnalyzer', 'kwargs': {'unit': 'KB'}}], analyzes_after_compress=True,
                                      analyzable_layer_key='bottleneck_layer')
    assert [n for n, _ in fx.named_children()] == ['bottleneck_layer', 'layer2', 'layer3', 'layer4']
    fx.eval()
    fx.update()
    fx.activate_analysis()
    oracle = ref_models.build_fp_bottleneck(3, 24, 256)
    oracle.load_state_dict(fx.bottleneck_layer.state_dict())
    oracle.eval()
    fx.to(dev)
    torch.manual_seed(1)
    x = torch.rand(1, 3, 800, 1344)
    with torch.inference_mode():
        feats = fx(x.to(dev))
        enc = fx.bottleneck_layer.encode(x.to(dev))
        latent, want_sym = oracle.symbols(x)
    assert list(feats.keys()) == ['1', '2', '3', '4'] and feats['1'].shape == (1, 256, 200, 336) and feats['4'].shape == (1, 2048, 25, 42)
    assert tuple(enc['shape']) == (199, 335) and len(fx.analyzers[0].file_size_list) == 1
    eb = oracle.entropy_bottleneck
    cdf, ln, off = eb._quantized_cdf.numpy(), eb._cdf_length.numpy(), eb._offset.numpy()
    idx = np.repeat(np.arange(24, dtype=np.int32), 199 * 335)
    got_sym = cref.decode_with_indexes(enc['strings'][0][0], idx, cdf, ln, off).reshape(1, 24, 199, 335)
    mism = int((got_sym != want_sym.numpy()).sum())
    assert mism <= symbol_budget(got_sym.size), 'symbol mismatches vs oracle at COCO shape: %d of %d' % (mism, got_sym.size)
    # the coder is bit-exact on its symbols, always (1.6 M-symbol stream)
    assert enc['strings'][0][0] == cref.encode_with_indexes(got_sym.reshape(-1), idx, cdf, ln, off)
    with torch.inference_mode():
        med = eb._get_medians().detach().reshape(1, -1, 1, 1)
        want_feat = oracle.decoder(torch.from_numpy(got_sym).float() + med)
    assert rel_err(feats['1'].cpu(), want_feat) < FEATURE_TOL


def test_shp_bottleneck_and_entropy_bottleneck_layer_vs_reference(s2, oracle_compressai):
    """the scale-hyperprior bottleneck (GaussianConditional coder with per-element indexes) against
    the REFERENCE'S OWN class when /root/reference is present (build container), else structure + round trip only; and
    EntropyBottleneckLayer on a large latent (256 x 56 x 56 = 802,816 symbols per image, 256 CDF rows)."""
    import os
    import sys
    dev = torch.device('cuda:0')
    # ---- EntropyBottleneckLayer at ResNet layer1 size ----
    from compressai.entropy_models import EntropyBottleneck as OracleEB
    torch.manual_seed(0)
    layer = s2.EntropyBottleneckLayer(entropy_bottleneck_channels=256)
    layer.update()
    oeb = OracleEB(256)
    oeb.load_state_dict({k: v for k, v in layer.entropy_bottleneck.state_dict().items() if not k.startswith('_')}, strict=False)
    oeb.update()
    assert torch.equal(oeb._quantized_cdf, layer.entropy_bottleneck._quantized_cdf)
    layer.eval().to(dev)
    torch.manual_seed(1)
    x = torch.randn(2, 256, 56, 56) * 3
    obj = layer.compress(x.to(dev))
    assert tuple(obj['shape']) == (56, 56) and len(obj['strings'][0]) == 2
    cdf, ln, off = oeb._quantized_cdf.numpy(), oeb._cdf_length.numpy(), oeb._offset.numpy()
    idx = np.repeat(np.arange(256, dtype=np.int32), 56 * 56)
    med = oeb._get_medians().detach().reshape(1, -1, 1, 1)
    want_sym = torch.round(x - med).int().numpy()
    for b in range(2):
        assert obj['strings'][0][b] == cref.encode_with_indexes(want_sym[b].reshape(-1), idx, cdf, ln, off)
    back = layer.decompress(**obj)
    assert torch.equal(back.cpu(), torch.from_numpy(want_sym).float() + med)
    # ---- SHP bottleneck ----
    torch.manual_seed(3)
    shp = s2.get_layer('SHPBasedResNetBottleneck', num_latent_channels=16, num_bottleneck_channels=24, num_target_channels=256)
    shp.eval()
    shp.update()
    assert shp.updated and tuple(shp.gaussian_conditional._quantized_cdf.shape) == (64, 3133)
    shp.to(dev)
    torch.manual_seed(4)
    x = torch.randn(2, 3, 224, 224)
    with torch.inference_mode():
        enc = shp.encode(x.to(dev))
        dec = shp.decode(**enc)
    assert len(enc['strings']) == 2 and all(len(l) == 2 for l in enc['strings']) and dec.shape == (2, 256, 56, 56)
    assert set(shp.__dict__['_tc_analysis']) == {'h_a', 'h_s'} and shp.__dict__.get('_tc_encoder') is not None  # no fp32 CUDA-core fallback
    # Oracle pipeline for SHPBasedResNetBottleneck.encode / decode (sc2bench/models/layer.py:631-666), runnable on the GPU box:
    # the module's own layers as plain torch CPU ops + the restated CompressAI entropy models + the C coder.
    import copy
    from compressai.entropy_models import GaussianConditional as OracleGC
    cpu = copy.deepcopy(shp).cpu()
    oeb = OracleEB(16)
    oeb.load_state_dict({k: v.cpu() for k, v in shp.entropy_bottleneck.state_dict().items() if not k.startswith('_')}, strict=False)
    oeb.update()
    ogc = OracleGC(None)
    ogc.update_scale_table(s2.models.get_scale_table())
    assert torch.equal(oeb._quantized_cdf, shp.entropy_bottleneck._quantized_cdf.cpu())
    assert torch.equal(ogc._quantized_cdf, shp.gaussian_conditional._quantized_cdf.cpu())
    with torch.inference_mode():
        y = _torch_seq(cpu.g_a, x)
        z = _torch_seq(cpu.h_a, y.abs())
        z_want = oeb.compress(z)
        z_hat = oeb.decompress(z_want, z.shape[-2:])
        idx_want = ogc.build_indexes(_torch_seq(cpu.h_s, z_hat))
    # z symbols: decode this implementation's z bytes with the oracle coder
    zc, zl, zo = oeb._quantized_cdf.numpy(), oeb._cdf_length.numpy(), oeb._offset.numpy()
    zidx = np.repeat(np.arange(16, dtype=np.int32), int(np.prod(z.shape[-2:])))
    z_got = _decode_symbols(enc['strings'][1], zc, zl, zo, zidx)
    z_ref = _decode_symbols(z_want, zc, zl, zo, zidx)
    mism_z = int((z_got != z_ref).sum())
    assert mism_z <= symbol_budget(z_ref.size), 'z symbol mismatches: %d of %d' % (mism_z, z_ref.size)
    for b in range(2):
        assert enc['strings'][1][b] == cref.encode_with_indexes(z_got[b], zidx, zc, zl, zo)
    # y symbols: the oracle's indexes (from the ORACLE's z_hat: valid when z agrees, which the budget above allows to fail for
    # at most one symbol -- then the y comparison is meaningless and the z assertion message says why)
    assert mism_z == 0, 'a z symbol flipped (fp32 tie, within budget): rerun with another seed to compare y'
    gc_cdf, gc_len, gc_off = ogc._quantized_cdf.numpy(), ogc._cdf_length.numpy(), ogc._offset.numpy()
    y_want = torch.round(y).int().numpy()
    mism_y = 0
    for b in range(2):
        yi = idx_want[b].reshape(-1).int().numpy()
        y_got = cref.decode_with_indexes(enc['strings'][0][b], yi, gc_cdf, gc_len, gc_off)
        mism_y += int((y_got != y_want[b].reshape(-1)).sum())
        assert enc['strings'][0][b] == cref.encode_with_indexes(y_got, yi, gc_cdf, gc_len, gc_off)
    assert mism_y <= symbol_budget(y_want.size), 'y symbol mismatches: %d of %d' % (mism_y, y_want.size)
    with torch.inference_mode():
        want_dec = _torch_seq(cpu.g_s, torch.from_numpy(y_want).float())
    if mism_y == 0:
        assert rel_err(dec.cpu(), want_dec) < FEATURE_TOL
    else:  # one flipped y symbol moves a 3x3 neighbourhood of the features: compare everything else
        assert float(((dec.cpu() - want_dec).abs() > FEATURE_TOL * want_dec.abs().max()).float().mean()) < 1e-4


def test_entropic_classifier_wrapper(s2, oracle_compressai):
    """EntropicClassifier (wrapper.py:196-264) built the way the fine-tuning configs build it
    (encoder = the classifier's modules up to layer1, EntropyBottleneckLayer(256), decoder = the rest, classifier = fc).
    The analysed strings must equal the oracle coder's on the encoder features; logits must equal running the tail on the
    dequantised features."""
    import torchvision
    from compressai.entropy_models import EntropyBottleneck as OracleEB
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    net = torchvision.models.resnet50(weights=None)
    model = s2.wrapper.EntropicClassifier(
        net, encoder_config={'sequential': ['conv1', 'bn1', 'relu', 'maxpool', 'layer1']},
        compression_model_kwargs={'entropy_bottleneck_channels': 256},
        decoder_config={'sequential': ['layer2', 'layer3', 'layer4', 'avgpool']}, classifier_config={'sequential': ['fc']},
        analysis_config={'analyzes_after_compress': True, 'analyzer_configs': [{'key': 'FileSizeAnalyzer', 'kwargs': {'unit': 'KB'}}]})
    assert list(dict(model.encoder.named_children())) == ['conv1', 'bn1', 'relu', 'maxpool', 'layer1']
    assert s2.backbone.check_if_updatable(model) and model.get_aux_module() is model.entropy_bottleneck
    sd = model.state_dict()
    model.load_state_dict(sd)  # the reference's split load path
    model.update()
    assert model.bottleneck_updated
    model.eval().to(dev)
    captured = []
    model.analyzers.append(type('Grab', (), {'analyze': lambda self, o: captured.append(o), 'summarize': lambda self: None,
                                             'clear': lambda self: None})())
    model.activate_analysis()
    torch.manual_seed(1)
    x = torch.randn(2, 3, 128, 96).to(dev)
    with torch.inference_mode():
        logits = model(x)
        feats = model.encoder(x)
    assert logits.shape == (2, 1000) and len(captured) == 1 and len(model.analyzers[0].file_size_list) == 1
    obj = captured[0]
    assert tuple(obj['shape']) == tuple(feats.shape[-2:])
    eb = model.entropy_bottleneck.entropy_bottleneck
    oeb = OracleEB(256)
    oeb.load_state_dict({k: v.cpu() for k, v in eb.state_dict().items() if not k.startswith('_')}, strict=False)
    oeb.update()
    want = oeb.compress(feats.cpu())
    assert obj['strings'][0] == want
    with torch.inference_mode():
        f_hat = oeb.decompress(want, feats.shape[-2:]).to(dev)
        want_logits = model.classifier(torch.flatten(model.decoder(f_hat), 1))
    assert torch.equal(logits, want_logits)


def test_transform_stream_pipeline_matches_serial(s2):
    """Throughput mode (FPBasedResNetBottleneck.use_transform_stream): several batches in flight, transforms on one stream,
    lane-per-stream coders on per-batch streams.  Bitstreams and features must equal the serial path's, batch by batch."""
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    layer = s2.get_layer('FPBasedResNetBottleneck').eval()
    layer.update()
    layer.to(dev)
    torch.manual_seed(1)
    xs = [torch.randn(8, 3, 224, 224, device=dev) * (1 + i) for i in range(5)]
    with torch.inference_mode():
        want = []
        for x in xs:
            st, shape = layer.encode_packed(x)
            want.append((st.tolist(), layer.decode_packed(st, shape).clone()))
        assert layer.entropy_bottleneck.coder_layout is None if hasattr(layer.entropy_bottleneck, 'coder_layout') else True
        ts = layer.use_transform_stream(True)
        assert ts is not None and layer.entropy_bottleneck.coder_layout == 'throughput'
        side = [torch.cuda.Stream(device=dev) for _ in range(3)]
        depth, enc, got = 2, {}, {}
        for i in range(len(xs) + depth):
            if i < len(xs):
                with torch.cuda.stream(side[i % 3]):
                    enc[i] = layer.encode_packed(xs[i])
            if i >= depth:
                j = i - depth
                with torch.cuda.stream(side[j % 3]):
                    st, shape = enc.pop(j)
                    got[j] = (st, layer.decode_packed(st, shape))
        torch.cuda.synchronize()
        for j, (strings, feat) in enumerate(want):
            assert got[j][0].tolist() == strings
            assert torch.equal(got[j][1], feat)
        # the reference-facing calls work in this mode too (one host thread per batch would use host_wait=True)
        layer.use_transform_stream(ts, host_wait=True)
        obj = layer.encode(xs[0])
        assert obj['strings'][0] == want[0][0]
        assert torch.equal(layer.decode(**obj), want[0][1])
        layer.use_transform_stream(None)
        assert layer.entropy_bottleneck.coder_layout is None


def _ref_compute_accuracy(outputs, targets, topk=(1,)):
    """script/task/image_classification.py:91-103, restated for the test (the GPU box has no /root/reference)."""
    maxk = max(topk)
    batch_size = targets.size(0)
    _, preds = outputs.topk(maxk, 1, True, True)
    preds = preds.t()
    corrects = preds.eq(targets[None])
    return [corrects[:k].flatten().sum(dtype=torch.float32) * (100.0 / batch_size) for k in topk]


def test_evaluate_loop_and_device_counters_vs_reference_meters(s2):
    """a15: parallel.evaluate / compute_accuracy / EvalCounters on the GPU against the reference's per-batch compute_accuracy +
    SmoothedValue.global_avg arithmetic (image_classification.py:91-145), on the splittable ResNet-50 with the bottleneck running
    its compress -> decompress branch, ragged last batch included."""
    from sc2bench_b200 import parallel
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    model = s2.splittable_resnet(bottleneck_config={'key': 'FPBasedResNetBottleneck', 'kwargs': {'num_bottleneck_channels': 24, 'num_target_channels': 256}},
                                 resnet_name='resnet50', skips_avgpool=False, skips_fc=False, weights=None,
                                 analysis_config={'analyzes_after_compress': True, 'analyzer_configs': [{'key': 'FileSizeAnalyzer', 'kwargs': {'unit': 'KB'}}]})
    model.eval()
    model.update()
    model.activate_analysis()
    torch.manual_seed(1)
    batches = [(torch.randn(n, 3, 224, 224), torch.randint(0, 1000, (n,))) for n in (4, 4, 3)]
    # make some predictions right so that the counters are not trivially zero: reuse the model's own top-1 / top-3 as targets
    model.to(dev)
    with torch.inference_mode():
        logits = [model(x.to(dev)).cpu() for x, _ in batches]
    batches[0] = (batches[0][0], logits[0].argmax(1))
    batches[1] = (batches[1][0], logits[1].topk(3, 1).indices[:, 2])
    model.clear_analysis()
    # reference arithmetic: per-batch accuracies averaged with weights n (SmoothedValue total / count)
    tot1 = tot5 = cnt = 0.0
    for (x, t), lg in zip(batches, logits):
        a1, a5 = _ref_compute_accuracy(lg, t, topk=(1, 5))
        tot1 += a1.item() * len(x)
        tot5 += a5.item() * len(x)
        cnt += len(x)
        g1, g5 = parallel.compute_accuracy(lg.to(dev), t.to(dev), topk=(1, 5))
        assert g1.is_cuda and abs(g1.item() - a1.item()) < 1e-4 and abs(g5.item() - a5.item()) < 1e-4
    got = parallel.evaluate(model, batches, dev)
    assert abs(float(got) - tot1 / cnt) < 1e-4 and abs(got.top5 - tot5 / cnt) < 1e-4 and got.images == 11
    assert float(got) > 30.0 and got.top5 > 60.0           # the planted targets were found
    assert len(model.analyzers[0].file_size_list) == 3      # one analysed object per batch; summarize() ran at the end


@pytest.mark.parametrize('dtype', ['f32', 'u8'])
def test_native_codec_pipeline_matches_per_kernel_route(s2, dtype):
    """CodecPipeline with the one-call-per-batch codec (csrc/fp_codec.cu: preallocated slots, no allocation per batch) gives the
    SAME bytes and features as the per-kernel route, batch by batch, for fp32 and for uint8 (device-side normalisation) input."""
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    layer = s2.get_layer('FPBasedResNetBottleneck').eval()
    layer.update()
    layer.set_input_normalization((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
    layer.to(dev)
    torch.manual_seed(1)
    if dtype == 'u8':
        xs = [torch.randint(0, 256, (8, 3, 224, 224), dtype=torch.uint8, device=dev) for _ in range(5)]
    else:
        xs = [torch.randn(8, 3, 224, 224, device=dev) * (1 + i) for i in range(5)]
    with torch.inference_mode():
        want = []
        for x in xs:
            st, shape = layer.encode_packed(x)
            want.append((st.tolist(), layer.decode_packed(st, shape).clone()))
        pipe = s2.CodecPipeline(layer, depth=2, max_ahead=1)
        got = []
        for x in xs:
            r = pipe.submit(x)
            if r is not None:
                r.wait()
                got.append((r.streams.tolist(), r.features.clone()))
        for r in pipe.drain():
            r.wait()
            got.append((r.streams.tolist(), r.features.clone()))
        assert pipe._native is not None, 'the native codec was not used'
        pipe.close()
    assert len(got) == len(want)
    for (gs, gf), (ws, wf) in zip(got, want):
        assert gs == ws
        assert torch.equal(gf, wf)
