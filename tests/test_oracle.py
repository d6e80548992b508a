"""CPU: the oracle against the committed golden vectors, and its two independent restatements
(oracle/rans_oracle.c vs oracle/pyrans.py) against each other.  PARITY UNPINNED vs real CompressAI."""
import numpy as np
import pytest

import cref
import pyrans
from helpers import load_golden


@pytest.fixture(scope='module')
def g():
    return load_golden('rans_cases.npz')


def _table(g, name):
    if name.startswith('gc'):
        pytest.skip('gc table is checked in test_gc_tables')
    return g['eb24_cdf'], g['eb24_len'], g['eb24_off']


def test_eb24_default_tables_match_survey_probe(g):
    # default-init EntropyBottleneck(24): length 23, offset -10, row 0 = [0, 1184, 2429, ..., 31365, 65536]
    assert g['eb24_cdf'].shape == (24, 23)
    assert (g['eb24_len'] == 23).all() and (g['eb24_off'] == -10).all()
    row = g['eb24_cdf'][0]
    assert list(row[:3]) == [0, 1184, 2429] and list(row[-2:]) == [31365, 65536]
    assert (np.diff(g['eb24_cdf'], axis=1) > 0).all()


@pytest.mark.parametrize('name', ['eb24_sigma1', 'eb24_sigma3', 'eb24_sigma8', 'empty', 'single', 'single_escape',
                                  'edge_values', 'huge_escapes'])
def test_c_and_python_restatements_reproduce_golden_streams(g, name):
    cdf, ln, off = g['eb24_cdf'], g['eb24_len'], g['eb24_off']
    sym, idx, stream = g[name + '_symbols'], g[name + '_indexes'], g[name + '_stream'].tobytes()
    assert cref.encode_with_indexes(sym, idx, cdf, ln, off) == stream
    assert pyrans.encode_with_indexes(sym.tolist(), idx.tolist(), cdf.tolist(), ln.tolist(), off.tolist()) == stream
    assert (cref.decode_with_indexes(stream, idx, cdf, ln, off) == sym).all()
    assert pyrans.decode_with_indexes(stream, idx.tolist(), cdf.tolist(), ln.tolist(), off.tolist()) == sym.tolist()
    assert len(stream) % 4 == 0 and len(stream) >= 8


def test_empty_stream_is_the_initial_state(g):
    # x = RANS64_L = 1 << 31 flushed as two little-endian u32 words
    assert g['empty_stream'].tobytes() == (1 << 31).to_bytes(8, 'little')


def test_gc_tables(g, oracle_compressai):
    import torch
    from compressai.entropy_models import GaussianConditional
    from compressai.models import get_scale_table
    gc = GaussianConditional(None)
    gc.update_scale_table(get_scale_table())
    cdf = gc._quantized_cdf.numpy()
    assert tuple(g['gc_cdf_shape']) == cdf.shape == (64, 3133)
    assert (gc._cdf_length.numpy() == g['gc_len']).all() and (gc._offset.numpy() == g['gc_off']).all()
    assert (cdf[0, :g['gc_len'][0]] == g['gc_cdf_row0']).all()
    assert (cdf[31, :g['gc_len'][31]] == g['gc_cdf_row31']).all()
    assert (cdf[63, :64] == g['gc_cdf_row63_head']).all()
    import hashlib
    assert hashlib.sha256(np.ascontiguousarray(cdf).tobytes()).hexdigest() == str(g['gc_cdf_sha256'])
    sym, idx, stream = g['gc_mixed_symbols'], g['gc_mixed_indexes'], g['gc_mixed_stream'].tobytes()
    ln, off = gc._cdf_length.numpy(), gc._offset.numpy()
    assert cref.encode_with_indexes(sym, idx, cdf, ln, off) == stream
    assert (cref.decode_with_indexes(stream, idx, cdf, ln, off) == sym).all()
    assert torch.allclose(gc.scale_table, torch.from_numpy(g['gc_scale_table']))


def test_pmf_to_quantized_cdf_golden(g):
    for i in range(int(g['n_pmfs'])):
        cdf = cref.pmf_to_quantized_cdf(g['pmf%d' % i], 16).astype(np.int64)
        assert (cdf == g['pmf%d_cdf' % i]).all()
        assert cdf[0] == 0 and cdf[-1] == 65536 and (np.diff(cdf) > 0).all()


def test_pmf_to_quantized_cdf_rejects_bad_input():
    with pytest.raises(ValueError):
        cref.pmf_to_quantized_cdf(np.array([0.5, -0.1, 0.6], np.float32))
    with pytest.raises(ValueError):
        cref.pmf_to_quantized_cdf(np.array([0.5, np.nan], np.float32))
    with pytest.raises(ValueError):
        cref.pmf_to_quantized_cdf(np.zeros(4, np.float32))


def test_random_roundtrips_c_vs_python(g):
    cdf, ln, off = g['eb24_cdf'], g['eb24_len'], g['eb24_off']
    rng = np.random.RandomState(7)
    for n in (1, 2, 31, 32, 33, 257):
        for sigma in (0.5, 4.0, 40.0):
            sym = np.round(rng.randn(n) * sigma).astype(np.int32)
            idx = rng.randint(0, 24, size=n).astype(np.int32)
            s = cref.encode_with_indexes(sym, idx, cdf, ln, off)
            assert s == pyrans.encode_with_indexes(sym.tolist(), idx.tolist(), cdf.tolist(), ln.tolist(), off.tolist())
            assert (cref.decode_with_indexes(s, idx, cdf, ln, off) == sym).all()


def test_reference_bottleneck_goldens_are_self_consistent(oracle_compressai):
    """The golden bitstreams of the reference's FPBasedResNetBottleneck decode (with the oracle coder) to the golden
    symbols, and the stored tables in the state dict are what the oracle's update() yields."""
    g = load_golden('fp_bottleneck_small.npz')
    cdf, ln, off = g['sd/entropy_bottleneck._quantized_cdf'], g['sd/entropy_bottleneck._cdf_length'], g['sd/entropy_bottleneck._offset']
    sym = g['symbols']
    B, C, H, W = sym.shape
    idx = np.repeat(np.arange(C, dtype=np.int32), H * W)
    offs = g['stream_offsets']
    for b in range(B):
        stream = g['streams'][offs[b]:offs[b + 1]].tobytes()
        assert cref.encode_with_indexes(sym[b].reshape(-1), idx, cdf, ln, off) == stream
        assert (cref.decode_with_indexes(stream, idx, cdf, ln, off) == sym[b].reshape(-1)).all()
    med = g['sd/entropy_bottleneck.quantiles'][:, 0, 1].reshape(1, C, 1, 1)
    assert (g['latent_hat'] == sym.astype(np.float32) + med).all()
    assert (np.rint(g['latent'] - med).astype(np.int32) == sym).all()


def test_config1_golden_stream_decodes(g):
    c1 = load_golden('config1_entropic_student_resnet50.npz')
    sym = c1['symbols'].astype(np.int32)
    assert sym.shape == (1, 24, 55, 55) and tuple(c1['shape']) == (55, 55)
    idx = np.repeat(np.arange(24, dtype=np.int32), 55 * 55)
    stream = c1['stream'].tobytes()
    assert len(stream) == 49240
    assert (cref.decode_with_indexes(stream, idx, g['eb24_cdf'], g['eb24_len'], g['eb24_off']) == sym.reshape(-1)).all()
    assert cref.encode_with_indexes(sym.reshape(-1), idx, g['eb24_cdf'], g['eb24_len'], g['eb24_off']) == stream


def test_property_random_tables_c_vs_python_and_roundtrip():
    """hypothesis: random probability tables (through the oracle's own pmf_to_quantized_cdf), random offsets, symbols far
    into both escape tails -- the two independent restatements produce the same bytes and decode them back."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 4), st.integers(2, 40), st.integers(0, 2 ** 31 - 1), st.integers(1, 200))
    def check(n_rows, n_sym, seed, n):
        rng = np.random.RandomState(seed)
        stride = n_sym + 2
        cdf = np.zeros((n_rows, stride), dtype=np.int32)
        ln = np.zeros(n_rows, dtype=np.int32)
        off = rng.randint(-30, 5, size=n_rows).astype(np.int32)
        for r in range(n_rows):
            k = rng.randint(1, n_sym + 1)                    # symbols of this row (ragged rows)
            pmf = rng.rand(k + 1).astype(np.float32) ** 3     # + the tail-mass entry CompressAI appends; skewed
            row = cref.pmf_to_quantized_cdf(pmf / pmf.sum(), 16)
            cdf[r, :row.size] = row
            ln[r] = row.size
        scale = float(rng.choice([0.5, 3.0, 50.0, 5000.0]))
        sym = np.round(rng.randn(n) * scale).astype(np.int32)
        idx = rng.randint(0, n_rows, size=n).astype(np.int32)
        s = cref.encode_with_indexes(sym, idx, cdf, ln, off)
        assert len(s) % 4 == 0 and len(s) >= 8
        assert s == pyrans.encode_with_indexes(sym.tolist(), idx.tolist(), cdf.tolist(), ln.tolist(), off.tolist())
        assert (cref.decode_with_indexes(s, idx, cdf, ln, off) == sym).all()
        assert pyrans.decode_with_indexes(s, idx.tolist(), cdf.tolist(), ln.tolist(), off.tolist()) == sym.tolist()

    check()
