"""CPU: host-side logic of the product and the C-ABI surface (no GPU compute)."""
import ctypes
import os
import re
import struct

import numpy as np
import pytest
import torch

import cref
from helpers import load_golden, state_dict_from_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def s2():
    import sc2bench_b200
    return sc2bench_b200


def test_library_exports_every_declared_symbol(s2):
    header = open(os.path.join(ROOT, 'include', 'sc2b200.h')).read()
    declared = set(re.findall(r'SC2_API\s+[\w\s\*]+?\b(sc2_\w+)\s*\(', header))
    assert len(declared) >= 15
    lib = ctypes.CDLL(s2._native.lib_path())
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(s2._native.SIGNATURES), declared ^ set(s2._native.SIGNATURES)
    assert lib.sc2_abi_version() == int(re.search(r'#define SC2_ABI_VERSION (\d+)', header).group(1))


def test_pmf_to_quantized_cdf_matches_oracle(s2):
    g = load_golden('rans_cases.npz')
    for i in range(int(g['n_pmfs'])):
        got = s2.ops.pmf_to_quantized_cdf(g['pmf%d' % i]).numpy().astype(np.int64)
        assert (got == g['pmf%d_cdf' % i]).all()
    rng = np.random.RandomState(0)
    for _ in range(200):
        n = rng.randint(2, 400)
        pmf = (np.abs(rng.randn(n)) ** rng.randint(1, 8)).astype(np.float32)
        pmf /= pmf.sum()
        assert (s2.ops.pmf_to_quantized_cdf(pmf).numpy().astype(np.uint32) == cref.pmf_to_quantized_cdf(pmf)).all()
    with pytest.raises(ValueError):
        s2.ops.pmf_to_quantized_cdf(np.array([0.2, -1.0], np.float32))
    with pytest.raises(ValueError):
        s2.ops.pmf_to_quantized_cdf(np.zeros(3, np.float32))


def test_entropy_bottleneck_tables_match_oracle(s2, oracle_compressai):
    from compressai.entropy_models import EntropyBottleneck as OracleEB
    g = load_golden('rans_cases.npz')
    torch.manual_seed(0)
    eb = s2.EntropyBottleneck(24)
    assert eb.update() is True and eb.update() is False
    assert (eb._quantized_cdf.numpy() == g['eb24_cdf']).all()
    assert (eb._cdf_length.numpy() == g['eb24_len']).all() and (eb._offset.numpy() == g['eb24_off']).all()
    # a "trained-looking" model: ragged table lengths, non-zero medians
    small = load_golden('fp_bottleneck_small.npz')
    sd = {k[len('entropy_bottleneck.'):]: v for k, v in state_dict_from_golden(small).items() if k.startswith('entropy_bottleneck.')}
    mine, ref = s2.EntropyBottleneck(8), OracleEB(8)
    for m in (mine, ref):
        m.load_state_dict({k: v for k, v in sd.items() if not k.startswith('_')}, strict=False)
        assert m.update(force=True)
    for name in ('_quantized_cdf', '_cdf_length', '_offset'):
        assert torch.equal(getattr(mine, name), getattr(ref, name)), name
        assert torch.equal(getattr(mine, name), sd[name]), name
    assert float(mine.loss().detach()) == pytest.approx(float(ref.loss().detach()), rel=1e-6)


def test_gaussian_conditional_tables_match_oracle(s2):
    g = load_golden('rans_cases.npz')
    gc = s2.GaussianConditional(None)
    assert gc.update_scale_table(s2.get_scale_table()) is True
    assert gc.update_scale_table(s2.get_scale_table()) is False
    assert tuple(gc._quantized_cdf.shape) == tuple(g['gc_cdf_shape'])
    assert (gc._cdf_length.numpy() == g['gc_len']).all() and (gc._offset.numpy() == g['gc_off']).all()
    import hashlib
    assert hashlib.sha256(np.ascontiguousarray(gc._quantized_cdf.numpy()).tobytes()).hexdigest() == str(g['gc_cdf_sha256'])


def _parse_blob(blob):
    magic, n_rows, cdf_stride, dec_stride, meta_off, enc_off, dec_off, total = struct.unpack_from('<8i', blob, 0)
    assert magic == 0x54523253 and total == len(blob)
    meta = np.frombuffer(blob, np.int32, 2 * n_rows, meta_off)
    enc = np.frombuffer(blob, np.uint32, n_rows * cdf_stride * 4, enc_off).reshape(n_rows, cdf_stride, 4)
    dec = np.frombuffer(blob, np.int32, n_rows * dec_stride, dec_off).reshape(n_rows, dec_stride)
    return n_rows, cdf_stride, meta[:n_rows], meta[n_rows:], enc, dec


def test_encoder_reciprocals_divide_exactly(s2):
    """q = mulhi64(x, rcp) >> shift must equal x // freq for every reachable state x < 2^63 (checked with Python ints)."""
    g = load_golden('rans_cases.npz')
    gc = s2.GaussianConditional(None)
    gc.update_scale_table(s2.get_scale_table())
    rng = np.random.RandomState(3)
    for cdf, ln, off in ((g['eb24_cdf'], g['eb24_len'], g['eb24_off']),
                         (gc._quantized_cdf.numpy(), gc._cdf_length.numpy(), gc._offset.numpy())):
        t = s2.ops.CoderTables(torch.from_numpy(cdf), torch.from_numpy(ln), torch.from_numpy(off))
        n_rows, stride, sizes, offsets, enc, dec = _parse_blob(t._blob.numpy().tobytes())
        assert (sizes == ln).all() and (offsets == off).all()
        rows = range(n_rows) if n_rows <= 24 else (0, 17, 40, 63)
        for r in rows:
            assert (dec[r, :sizes[r]] == cdf[r, :sizes[r]]).all() and (dec[r, sizes[r]:] == 0x7fffffff).all()
            vs = range(sizes[r] - 1) if sizes[r] < 64 else rng.randint(0, sizes[r] - 1, size=64)
            for v in vs:
                rcp = int(enc[r, v, 0]) | (int(enc[r, v, 1]) << 32)
                bias, shift, freq = int(enc[r, v, 2]) & 0x1ffff, (int(enc[r, v, 2]) >> 24) & 15, int(enc[r, v, 3])
                start = int(cdf[r, v])
                assert freq == int(cdf[r, v + 1]) - start
                # the division is applied after renormalisation: x in [2^31, freq << 47) (or x >> 32 of such a state)
                x_max = freq << 47
                xs = [1 << 31, x_max - 1, x_max - freq, x_max // 2 + 1, freq, max(freq - 1, 1), (1 << 32) - 1, 1 << 32, (1 << 31) - 1]
                xs += [int(x) % x_max for x in rng.randint(0, 2 ** 62, size=20, dtype=np.int64) * 2 + 1]
                for x in xs:
                    if x <= 0 or x >= x_max:
                        continue
                    q = ((x * rcp) >> 64) >> shift
                    new = (x + bias + q * (65536 - freq)) & ((1 << 64) - 1)
                    assert new == ((x // freq) << 16) + (x % freq) + start, (r, v, x)


def test_table_builder_rejects_bad_tables(s2):
    bad = torch.tensor([[0, 5, 5, 65536]], dtype=torch.int32)  # zero-width bin
    with pytest.raises(ValueError):
        s2.ops.CoderTables(bad, torch.tensor([4], dtype=torch.int32), torch.tensor([0], dtype=torch.int32))
    bad2 = torch.tensor([[1, 5, 9, 65536]], dtype=torch.int32)  # does not start at 0
    with pytest.raises(ValueError):
        s2.ops.CoderTables(bad2, torch.tensor([4], dtype=torch.int32), torch.tensor([0], dtype=torch.int32))


def test_state_dict_is_interchangeable_with_the_reference_layout(s2, oracle_compressai):
    """A checkpoint written by the reference stack (oracle-backed here) loads into the product and back."""
    small = load_golden('fp_bottleneck_small.npz')
    sd = state_dict_from_golden(small)
    layer = s2.get_layer('FPBasedResNetBottleneck', num_input_channels=3, num_bottleneck_channels=8, num_target_channels=32)
    assert layer.updated is False
    assert set(layer.state_dict().keys()) | {'entropy_bottleneck._quantized_cdf', 'entropy_bottleneck._cdf_length', 'entropy_bottleneck._offset'} == set(sd.keys())
    layer.load_state_dict(sd)  # strict; CDF buffers are resized on the fly
    assert tuple(layer.entropy_bottleneck._quantized_cdf.shape) == tuple(sd['entropy_bottleneck._quantized_cdf'].shape)
    assert layer.update() is False  # tables came from the checkpoint -> early return, but the flag flips
    assert layer.updated is True
    out = layer.state_dict()
    for k, v in sd.items():
        assert torch.equal(out[k], v), k
    # legacy CompressAI <= 1.1 parameter names
    legacy = {}
    for k, v in sd.items():
        m = re.match(r'entropy_bottleneck\.(matrices|biases|factors)\.(\d)$', k)
        legacy['entropy_bottleneck._%s%s' % ({'matrices': 'matrix', 'biases': 'bias', 'factors': 'factor'}[m.group(1)], m.group(2)) if m else k] = v
    layer2 = s2.get_layer('FPBasedResNetBottleneck', num_input_channels=3, num_bottleneck_channels=8, num_target_channels=32)
    layer2.load_state_dict(legacy)
    assert torch.equal(layer2.entropy_bottleneck.matrices[2], sd['entropy_bottleneck.matrices.2'])


def test_registry_and_contract_surface(s2):
    assert s2.get_layer('no_such_layer') is None
    assert 'FPBasedResNetBottleneck' in s2.LAYER_CLASS_DICT
    layer = s2.get_layer('FPBasedResNetBottleneck')
    assert isinstance(layer, s2.CompressionModel) and isinstance(layer, s2.BaseBottleneck)
    assert [type(m).__name__ for m in layer.encoder] == ['Conv2d', 'GDN1', 'Conv2d', 'GDN1', 'Conv2d']
    assert [type(m).__name__ for m in layer.decoder] == ['Conv2d', 'GDN1', 'Conv2d', 'GDN1', 'Conv2d']
    assert [m.weight.shape for m in layer.encoder if hasattr(m, 'weight')] == [(96, 3, 5, 5), (48, 96, 5, 5), (24, 48, 2, 2)]
    assert [m.weight.shape for m in layer.decoder if hasattr(m, 'weight')] == [(512, 24, 2, 2), (256, 512, 2, 2), (256, 256, 2, 2)]
    assert sum(p.numel() for n, p in layer.named_parameters() if not n.startswith('entropy_bottleneck')) == 1302704
    assert sum(p.numel() for p in layer.entropy_bottleneck.parameters()) == 24 * 61
    with pytest.raises(ValueError, match='Uninitialized CDFs'):
        layer.entropy_bottleneck.coder_tables()
    m = s2.splittable_resnet(bottleneck_config={'key': 'FPBasedResNetBottleneck', 'kwargs': {}}, resnet_name='resnet18',
                             skips_avgpool=False, skips_fc=False, weights=None,
                             analysis_config={'analyzes_after_compress': True, 'analyzer_configs': [{'key': 'FileSizeAnalyzer', 'kwargs': {'unit': 'KB'}}]})
    assert s2.check_if_updatable(m) and m.get_aux_module() is m.bottleneck_layer
    assert not m.bottleneck_updated
    m.update()
    assert m.bottleneck_updated and m.bottleneck_layer.updated
    m.activate_analysis()
    m.analyze({'strings': [[b'12345678']], 'shape': (1, 1)})
    assert len(m.analyzers[0].file_size_list) == 1


def test_training_branches_run_on_cpu_and_hot_path_refuses_cpu(s2):
    torch.manual_seed(0)
    layer = s2.get_layer('FPBasedResNetBottleneck', num_bottleneck_channels=4, num_target_channels=8)
    x = torch.randn(2, 3, 32, 32, requires_grad=True)
    layer.train()
    y = layer(x)  # noise / likelihood branch
    assert y.shape == (2, 8, 8, 8)
    (y.sum() + layer.aux_loss()).backward()
    y_hat, lik = layer.entropy_bottleneck(layer.encoder(x))
    assert lik.shape == y_hat.shape and float(lik.min()) > 0
    layer.update()
    y2 = layer(x)  # fine-tuning branch: rounding, detached
    assert y2.shape == y.shape
    layer.eval()
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        layer(x.detach())
    with pytest.raises(RuntimeError, match='CUDA'):
        layer.entropy_bottleneck.decompress([b'\x00' * 8], (7, 7))


def test_forward_matches_oracle_in_training_mode(s2, oracle_compressai):
    """The differentiable (torch) branches agree with the reference stack bit for bit on CPU."""
    from compressai.layers import GDN1 as OracleGDN1
    torch.manual_seed(3)
    a, b = s2.GDN1(6), OracleGDN1(6)
    b.load_state_dict(a.state_dict())
    x = torch.randn(2, 6, 5, 5, requires_grad=True)
    assert torch.equal(a(x), b(x))
    ai, bi = s2.GDN(6, inverse=True), oracle_compressai.layers.GDN(6, inverse=True)
    bi.load_state_dict(ai.state_dict())
    assert torch.equal(ai(x), bi(x))
    eb_a, eb_b = s2.EntropyBottleneck(6), oracle_compressai.entropy_models.EntropyBottleneck(6)
    eb_b.load_state_dict(eb_a.state_dict())
    eb_a.eval(), eb_b.eval()
    ya, la = eb_a(x)
    yb, lb = eb_b(x)
    assert torch.equal(ya, yb) and torch.allclose(la, lb, rtol=0, atol=0)


def test_hostbytes_split_and_join_roundtrip(s2):
    """csrc/hostbytes.c: the contract's list[bytes] <-> one staging buffer (copies run with the GIL released)."""
    hb = s2._native.hostbytes()
    rng = np.random.RandomState(3)
    lens = [8, 12, 0, 49240, 8, 4096]
    buf = rng.randint(0, 256, size=sum(lens), dtype=np.uint8)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    parts = hb.split(buf, offs)
    assert [len(p) for p in parts] == lens and all(isinstance(p, bytes) for p in parts)
    assert b''.join(parts) == buf.tobytes()
    dst = np.zeros(buf.size + 7, dtype=np.uint8)
    offs2 = np.zeros(len(lens) + 1, dtype=np.int64)
    assert hb.join(parts, dst, offs2) == buf.size
    assert (offs2 == offs).all() and (dst[:buf.size] == buf).all() and not dst[buf.size:].any()
    assert hb.join(tuple(parts), dst[:10], offs2) == -1           # destination too small: nothing is written past it
    assert hb.split(buf[:0], np.zeros(1, dtype=np.int64)) == []
    with pytest.raises(TypeError):
        hb.join([b'12345678', 'not bytes'], dst, offs2)
    with pytest.raises(ValueError):
        hb.split(buf, np.array([0, buf.size + 1], dtype=np.int64))  # offsets beyond the buffer
    with pytest.raises(ValueError):
        hb.join(parts, dst, np.zeros(2, dtype=np.int64))            # offsets buffer too small


def test_coder_layouts_and_throughput_mode_surface(s2):
    """The coder layout constants of the header match the Python table; the throughput-mode entry points refuse a CPU layer."""
    header = open(os.path.join(ROOT, 'include', 'sc2b200.h')).read()
    consts = dict(re.findall(r'#define (SC2_RANS_\w+) (\d+)', header))
    assert consts == {'SC2_RANS_AUTO': '0', 'SC2_RANS_WARP_PER_STREAM': '1', 'SC2_RANS_LANE_PER_STREAM': '2'}
    assert s2._native.RANS_LAYOUTS == {None: 0, 'auto': 0, 'warp': 1, 'lanes': 2}
    assert os.environ.get('CUDA_DEVICE_MAX_CONNECTIONS')  # set at import (one hardware queue per stream), unless the user chose
    layer = s2.get_layer('FPBasedResNetBottleneck', num_bottleneck_channels=8, num_target_channels=64).eval()
    layer.update()
    with pytest.raises(RuntimeError, match='CUDA only'):
        s2.CodecPipeline(layer, depth=2)
    with pytest.raises(ValueError):
        s2.CodecPipeline(layer, depth=0)
    assert layer.use_transform_stream(None) is None and layer.entropy_bottleneck.coder_layout is None


def test_header_is_valid_c_and_struct_layouts_match_ctypes(s2, tmp_path):
    """include/sc2b200.h compiles as plain C (gcc) and its descriptor structs have the size / field offsets the ctypes mirror
    in _native.py assumes (the descriptors are passed by pointer across the ABI)."""
    import subprocess
    src = tmp_path / 'abi.c'
    src.write_text('''#include <stdio.h>
#include <stddef.h>
#include "sc2b200.h"
int main(void) {
    printf("%zu %zu %zu\\n", sizeof(sc2_conv_desc), offsetof(sc2_conv_desc, in_transform), offsetof(sc2_conv_desc, epi_param));
    printf("%zu %zu\\n", sizeof(sc2_tc_conv_desc), offsetof(sc2_tc_conv_desc, mode));
    printf("%zu %zu\\n", sizeof(sc2_tc_split_desc), offsetof(sc2_tc_split_desc, out_c));
    printf("%d\\n", SC2_ABI_VERSION);
    return 0;
}
''')
    exe = tmp_path / 'abi'
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    lines = subprocess.check_output([str(exe)]).decode().split('\n')
    n = s2._native
    assert lines[0].split() == [str(ctypes.sizeof(n.ConvDesc)), str(n.ConvDesc.in_transform.offset), str(n.ConvDesc.epi_param.offset)]
    assert lines[1].split() == [str(ctypes.sizeof(n.TcConvDesc)), str(n.TcConvDesc.mode.offset)]
    assert lines[2].split() == [str(ctypes.sizeof(n.TcSplitDesc)), str(n.TcSplitDesc.out_c.offset)]
    assert int(lines[3]) == n.load().sc2_abi_version()


def test_c_program_links_the_library_and_builds_tables(s2, tmp_path):
    """A plain C caller (gcc, no Python in between) links libsc2b200.so and runs the host-side entry points the reference's
    update() needs: sc2_pmf_to_quantized_cdf (vs the oracle) and sc2_rans_build_tables on the result."""
    import subprocess
    pmf = np.array([0.02, 0.08, 0.2, 0.4, 0.2, 0.08, 0.02], dtype=np.float32)
    want = cref.pmf_to_quantized_cdf(pmf, 16)
    src = tmp_path / 'caller.c'
    src.write_text('''#include <stdio.h>
#include <stdlib.h>
#include "sc2b200.h"
int main(void) {
    const float pmf[7] = {0.02f, 0.08f, 0.2f, 0.4f, 0.2f, 0.08f, 0.02f};
    uint32_t cdf[8];
    int rc = sc2_pmf_to_quantized_cdf(pmf, 7, 16, cdf);
    if (rc != SC2_OK) { printf("error %s\\n", sc2_error_string(rc)); return 1; }
    for (int i = 0; i < 8; ++i) printf("%u ", cdf[i]);
    printf("\\n");
    int32_t row[8], size = 8, offset = -3;
    for (int i = 0; i < 8; ++i) row[i] = (int32_t)cdf[i];
    size_t bytes = sc2_rans_table_bytes(1, 8);
    void *blob = malloc(bytes);
    rc = sc2_rans_build_tables(row, &size, &offset, 1, 8, blob);
    printf("%d %zu %d\\n", rc, bytes, sc2_abi_version());
    row[3] = row[2];  /* not strictly increasing: must be refused, not crash */
    printf("%d\\n", sc2_rans_build_tables(row, &size, &offset, 1, 8, blob));
    free(blob);
    return 0;
}
''')
    exe = tmp_path / 'caller'
    lib_dir = os.path.dirname(s2._native.lib_path())
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe),
                           '-L', lib_dir, '-lsc2b200', '-Wl,-rpath,' + lib_dir])
    lines = subprocess.check_output([str(exe)]).decode().strip().split('\n')
    assert [int(v) for v in lines[0].split()] == want.tolist()
    rc, nbytes, abi = lines[1].split()
    assert int(rc) == 0 and int(nbytes) >= 8 * 16 and int(abi) == s2._native.load().sc2_abi_version()
    assert int(lines[2]) < 0


def _header_prototypes():
    """{name: [parameter type strings]} for every SC2_API function include/sc2b200.h declares."""
    import re
    text = open(os.path.join(ROOT, 'include', 'sc2b200.h')).read()
    text = re.sub(r'/\*.*?\*/', ' ', text, flags=re.S)
    protos = {}
    for m in re.finditer(r'SC2_API\s+[\w\s\*]+?\b(sc2_\w+)\s*\(([^;]*?)\)\s*;', text, flags=re.S):
        params = [p.strip() for p in m.group(2).replace('\n', ' ').split(',')]
        protos[m.group(1)] = [] if params == ['void'] else params
    return protos


def test_ctypes_signatures_match_the_header(s2):
    """Every function the header declares is bound in _native.SIGNATURES with the same NUMBER of parameters and the same
    pointer / integer / float class per parameter (a mismatch works by accident under cdecl until it does not), and the
    library exports every one of them."""
    protos = _header_prototypes()
    sigs = s2._native.SIGNATURES
    assert set(protos) == set(sigs), set(protos) ^ set(sigs)
    lib = s2._native.load()
    for name, params in protos.items():
        assert hasattr(lib, name), name
        argtypes = sigs[name][1]
        assert len(argtypes) == len(params), '%s: header has %d parameters, ctypes %d' % (name, len(params), len(argtypes))
        for p, a in zip(params, argtypes):
            is_ptr = '*' in p or 'sc2_stream_t' in p
            if is_ptr:
                assert a is ctypes.c_void_p or a is ctypes.c_char_p or hasattr(a, 'contents'), (name, p, a)
            elif 'float' in p:
                assert a is ctypes.c_float, (name, p, a)
            elif 'int64_t' in p:
                assert a is ctypes.c_int64, (name, p, a)
            elif 'size_t' in p:
                assert a is ctypes.c_size_t, (name, p, a)
            else:
                assert a is ctypes.c_int, (name, p, a)


def test_coder_layout_resolution_and_channel_tiles():
    """Host logic behind the round-2 entry points: the 'throughput' coder layout resolves per batch; output-channel tiles of the
    extended split convolution cover c_out exactly with tiles the kernel has (sc2_tc_split_n_tile)."""
    import sc2bench_b200 as s2
    N = s2._native
    assert N.rans_layout('throughput', 256) == N.RANS_LAYOUTS['lanes'] and N.rans_layout('throughput', 32) == N.RANS_LAYOUTS['lanes']
    assert N.rans_layout('throughput', 31) == N.RANS_LAYOUTS['warp'] and N.rans_layout('throughput', 1) == N.RANS_LAYOUTS['warp']
    assert N.rans_layout(None, 5) == 0 and N.rans_layout('lanes', 1) == 2 and N.rans_layout('warp', 999) == 1
    lib = N.load()
    for c_out in (8, 24, 96, 128, 144, 192, 200, 256, 320, 384, 512):
        tiles = s2.ops.split_n_tiles(c_out)
        assert tiles[0][0] == 0 and sum(n for _, n in tiles) == c_out
        assert all(a + n == b for (a, n), (b, _) in zip(tiles, tiles[1:]))
        for c0, n in tiles[:-1]:
            assert lib.sc2_tc_split_n_tile(n) == n and c0 % 8 == 0  # inner tiles are full tiles of a size the kernel has
        assert 0 < lib.sc2_tc_split_n_tile(tiles[-1][1]) <= 128


def test_split_analysis_plan_coverage_rules():
    """SplitAnalysisPlan.why_not: which transforms of the zoo codecs / hyperprior bottlenecks the tensor-core plan takes (no GPU needed)."""
    import sc2bench_b200 as s2
    P = s2.models.SplitAnalysisPlan
    m = s2.models.bmshj2018_hyperprior(8)
    assert P.why_not(m.g_a, (3, 256, 256)) is None
    assert P.why_not(m.h_a, (320, 16, 16), planes_in=True) is None
    assert P.why_not(m.h_s, (192, 4, 4)) is None
    assert P.why_not(m.g_a, (3, 250, 250)) is None           # 125 x 125 after the first layer: odd sizes are zero-padded
    f = s2.models.bmshj2018_factorized(1)
    assert P.why_not(f.g_a, (3, 64, 64)) is None and P.why_not(f.g_s, (192, 4, 4)) is not None  # g_s has inverse GDNs: ZooSynthesisPlan's job
    assert s2.models.ZooSynthesisPlan.why_not(f.g_s) is None
    shp = s2.get_layer('SHPBasedResNetBottleneck', num_latent_channels=16, num_bottleneck_channels=24, num_target_channels=256)
    assert P.why_not(shp.h_s, (16, 7, 7)) is None and P.why_not(shp.h_a, (24, 55, 55)) is None  # k5 s2 p1 transposed convs, 24 channels
    grouped = torch.nn.Sequential(torch.nn.Conv2d(16, 16, 3, groups=2))
    assert P.why_not(grouped, (16, 8, 8)) is not None
    assert s2.bottleneck.TensorCoreAnalysis.why_not(shp.g_a, (2, 3, 224, 224)) is None


@pytest.mark.parametrize('padding,output_padding', [(2, 1), (1, 0)])
def test_deconv5_parity_taps_rebuild_the_transposed_convolution(padding, output_padding):
    """ops.deconv5_parity_taps: ConvTranspose2d(k5, s2, p) as four stride-1 sub-convolutions, checked with torch CPU convolutions
    (the tap lists and paddings are what the tensor-core plans launch; no GPU needed)."""
    import torch.nn.functional as F
    import sc2bench_b200 as s2
    torch.manual_seed(padding)
    x = torch.randn(2, 3, 5, 7, dtype=torch.float64)
    w = torch.randn(3, 4, 5, 5, dtype=torch.float64)
    ref = F.conv_transpose2d(x, w, stride=2, padding=padding, output_padding=output_padding)
    taps = s2.ops.deconv5_parity_taps(padding)
    out = torch.full_like(ref, float('nan'))
    Ho, Wo = ref.shape[-2:]
    for py in (0, 1):
        for px in (0, 1):
            (ky, pad_y), (kx, pad_x) = taps[py], taps[px]
            sub = w[:, :, ky][:, :, :, kx].permute(1, 0, 2, 3)                      # conv weight [c_out, c_in, Ty, Tx]
            hs, ws = (Ho - py + 1) // 2, (Wo - px + 1) // 2                         # pixels of this parity
            # out[Y] = sum_j x[Y + j - pad] * w[taps[j]] for Y in [0, hs): pad the input so that every tap position exists
            xp = F.pad(x, (pad_x, ws + len(kx) - 1 - pad_x - x.shape[3], pad_y, hs + len(ky) - 1 - pad_y - x.shape[2]))
            y = F.conv2d(xp, sub)
            assert y.shape[-2:] == (hs, ws)
            out[:, :, py::2, px::2] = y
    assert torch.isfinite(out).all() and float((out - ref).abs().max()) < 1e-12
