"""oracle/cref.py -- ctypes front-end of oracle/rans_oracle.c (builds it on first use).

TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (see oracle/README.md).  numpy in, numpy/bytes out.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, '_build', 'liboracle.so')
_lib = None


def build(force=False):
    src = os.path.join(_HERE, 'rans_oracle.c')
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-s', '-B'] if force else ['make', '-C', _HERE, '-s'])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_long
        L.orc_rans_max_bytes.restype = i64
        L.orc_rans_max_bytes.argtypes = [i64]
        L.orc_rans_encode.restype = i64
        L.orc_rans_encode.argtypes = [vp, vp, i64, vp, i32, vp, vp, vp, i64]
        L.orc_rans_decode.restype = i32
        L.orc_rans_decode.argtypes = [vp, i64, vp, i64, vp, i32, vp, vp, vp]
        L.orc_pmf_to_quantized_cdf.restype = i32
        L.orc_pmf_to_quantized_cdf.argtypes = [vp, i32, i32, vp]
        _lib = L
    return _lib


def _i32(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.int32)


def _tables(cdfs, cdf_sizes, offsets):
    cdfs = _i32(cdfs)
    assert cdfs.ndim == 2
    return cdfs, _i32(cdf_sizes).reshape(-1), _i32(offsets).reshape(-1)


def encode_with_indexes(symbols, indexes, cdfs, cdf_sizes, offsets):
    symbols, indexes = _i32(symbols).reshape(-1), _i32(indexes).reshape(-1)
    assert symbols.shape == indexes.shape
    cdfs, cdf_sizes, offsets = _tables(cdfs, cdf_sizes, offsets)
    L = lib()
    cap = L.orc_rans_max_bytes(symbols.size)
    out = np.empty(cap, dtype=np.uint8)
    n = L.orc_rans_encode(symbols.ctypes.data, indexes.ctypes.data, symbols.size, cdfs.ctypes.data,
                          cdfs.shape[1], cdf_sizes.ctypes.data, offsets.ctypes.data, out.ctypes.data, cap)
    if n < 0:
        raise RuntimeError('orc_rans_encode failed: %d' % n)
    return out[:n].tobytes()


def decode_with_indexes(stream, indexes, cdfs, cdf_sizes, offsets):
    indexes = _i32(indexes).reshape(-1)
    cdfs, cdf_sizes, offsets = _tables(cdfs, cdf_sizes, offsets)
    buf = np.frombuffer(stream, dtype=np.uint8)
    out = np.empty(indexes.size, dtype=np.int32)
    rc = lib().orc_rans_decode(buf.ctypes.data, buf.size, indexes.ctypes.data, indexes.size, cdfs.ctypes.data,
                               cdfs.shape[1], cdf_sizes.ctypes.data, offsets.ctypes.data, out.ctypes.data)
    if rc != 0:
        raise RuntimeError('orc_rans_decode failed: %d' % rc)
    return out


def pmf_to_quantized_cdf(pmf, precision=16):
    pmf = np.ascontiguousarray(np.asarray(pmf), dtype=np.float32).reshape(-1)
    out = np.empty(pmf.size + 1, dtype=np.uint32)
    rc = lib().orc_pmf_to_quantized_cdf(pmf.ctypes.data, pmf.size, precision, out.ctypes.data)
    if rc == -1:
        raise ValueError('Invalid `pmf`, non-finite or negative element found.')
    if rc != 0:
        raise ValueError('Invalid `pmf`: at least one element must have a non-zero probability.')
    return out
