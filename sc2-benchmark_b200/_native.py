"""ctypes binding of libsc2b200.so (declared in include/sc2b200.h).

The library is the product: there is NO Python/CPU fallback for the hot path.  If the shared object is
missing it is built in-tree with nvcc; if that is impossible, importing this module raises.
"""
import ctypes
import os

from . import build as _build

_c = ctypes
vp, i32, i64, f32 = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float


class ConvDesc(_c.Structure):
    """struct sc2_conv_desc"""
    _fields_ = [(n, i32) for n in ('batch', 'c_in', 'h_in', 'w_in', 'c_out', 'kh', 'kw', 'stride', 'pad',
                                    'transposed', 'output_padding', 'epilogue', 'in_transform')] + [('epi_param', _c.c_float)]


class TcConvDesc(_c.Structure):
    """struct sc2_tc_conv_desc"""
    _fields_ = [(n, i32) for n in ('batch', 'h_in', 'w_in', 'c_in_pad', 'c_out', 'kh', 'kw', 'pad', 'mode')]


class TcSplitDesc(_c.Structure):
    """struct sc2_tc_split_desc"""
    _fields_ = [(n, i32) for n in ('images', 'h_in', 'w_in', 'c_in', 'c_out', 'kh', 'kw', 'stride', 'pad', 'mode',
                                    'h_out', 'w_out', 'out_c')]


class TcSplitExDesc(_c.Structure):
    """struct sc2_tc_split_ex_desc"""
    _fields_ = [(n, i32) for n in ('images', 'h_in', 'w_in', 'c_in', 'c_out', 'kh', 'kw', 'stride', 'pad', 'mode', 'h_out', 'w_out',
                                    'out_pitch', 'n_off', 'c_total', 'in_nhwc', 'act')] + [('slope', _c.c_float)] + \
               [(n, i32) for n in ('pad_x', 'out_stride', 'out_py', 'out_px', 'out_h', 'out_w')]


class TcConvExDesc(_c.Structure):
    """struct sc2_tc_conv_ex_desc"""
    _fields_ = [(n, i32) for n in ('batch', 'h_in', 'w_in', 'c_in_pad', 'c_out', 'kh', 'kw', 'pad_y', 'pad_x', 'mode', 'h_out', 'w_out',
                                    'out_h', 'out_w', 'out_stride', 'out_py', 'out_px')]


class FpPlan(_c.Structure):
    """struct sc2_fp_plan"""
    _fields_ = [(n, i32) for n in ('batch', 'h_in', 'w_in', 'c1', 'c2', 'c3', 'k1', 'k2', 'k3', 'p3', 'd1', 'd2', 'd3', 'kd1', 'pd1',
                                    'kd2', 'pd2', 'kd3', 'pd3', 'n_rows', 'cdf_stride')] + \
               [(n, vp) for n in ('w1_stack', 'g1_stack', 'beta1', 'w2_stack', 'g2_stack', 'beta2', 'w3_hi', 'w3_lo', 'medians', 'lut',
                                   'tables', 'wd1', 'gd1', 'wd2', 'gd2', 'wd3', 'betad1', 'betad2')]


class GaHaloDesc(_c.Structure):
    """struct sc2_ga_halo_desc"""
    _fields_ = [(n, i32) for n in ('images', 'h_in', 'w_in', 'c_in', 'c_out', 'kh', 'kw', 'pad', 'h_out', 'w_out', 'out_c')]


# name -> (restype, argtypes): every symbol include/sc2b200.h declares
SIGNATURES = {
    'sc2_abi_version': (i32, []),
    'sc2_set_persistent_ctas': (i32, [i32]),
    'sc2_trace_start': (i32, [vp, i64]),
    'sc2_trace_stop': (i32, []),
    'sc2_error_string': (_c.c_char_p, [i32]),
    'sc2_last_cuda_error': (_c.c_char_p, []),
    'sc2_pmf_to_quantized_cdf': (i32, [vp, i32, i32, vp]),
    'sc2_rans_table_bytes': (_c.c_size_t, [i32, i32]),
    'sc2_rans_build_tables': (i32, [vp, vp, vp, i32, i32, vp]),
    'sc2_rans_max_stream_bytes': (i64, [i64]),
    'sc2_rans_encode_batch': (i32, [vp, vp, i32, i64, i64, vp, i32, i32, vp, i64, vp, vp, i32, vp]),
    'sc2_rans_pack': (i32, [vp, i64, vp, i32, vp, vp, vp]),
    'sc2_rans_decode_batch': (i32, [vp, vp, i32, i64, vp, i64, vp, i32, i32, vp, vp, vp, vp, i32, vp]),
    'sc2_quantize_symbols': (i32, [vp, vp, vp, i32, i32, i64, vp]),
    'sc2_dequantize': (i32, [vp, vp, vp, i64, vp]),
    'sc2_gc_build_indexes': (i32, [vp, i64, vp, i32, f32, vp, vp]),
    'sc2_conv_out_size': (i32, [_c.POINTER(ConvDesc), _c.POINTER(i32), _c.POINTER(i32)]),
    'sc2_conv2d_f32': (i32, [_c.POINTER(ConvDesc), vp, vp, vp, vp, vp, vp]),
    'sc2_gdn_f32': (i32, [vp, vp, vp, vp, i32, i32, i64, i32, i32, vp]),
    'sc2_tc_conv_nhwc': (i32, [_c.POINTER(TcConvDesc), vp, vp, vp, vp, vp, vp, vp, vp]),
    'sc2_tc_conv_ex': (i32, [_c.POINTER(TcConvExDesc), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    'sc2_nchw_f32_to_nhwc_f16': (i32, [vp, vp, i32, i32, i64, i32, vp]),
    'sc2_tc_split_n_tile': (i32, [i32]),
    'sc2_tc_split_conv': (i32, [_c.POINTER(TcSplitDesc), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    'sc2_tc_split_conv_ex': (i32, [_c.POINTER(TcSplitExDesc), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    'sc2_patchify_split_nhwc': (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
    'sc2_ga_halo_n': (i32, [i32]),
    'sc2_ga_halo_conv_gdn': (i32, [_c.POINTER(GaHaloDesc), vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    'sc2_ga_first_conv_gdn': (i32, [vp, i32, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, i32, vp, vp]),
    'sc2_fp_workspace_bytes': (i32, [_c.POINTER(FpPlan), _c.POINTER(i64), _c.POINTER(i64), _c.POINTER(i64), _c.POINTER(i32), _c.POINTER(i32),
                                     _c.POINTER(i32), _c.POINTER(i32)]),
    'sc2_fp_encode_batch': (i32, [_c.POINTER(FpPlan), vp, i32, vp, vp, vp, i64, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp]),
    'sc2_fp_decode_batch': (i32, [_c.POINTER(FpPlan), vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp]),
    'sc2_patchify_split': (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
    'sc2_tc_first_layer': (i32, [vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, i32, vp, vp]),
}

SC2_OK = 0
ABI_VERSION = 10  # include/sc2b200.h SC2_ABI_VERSION
FAULT_ARENA_OVERFLOW, FAULT_STREAM_TRUNCATED, FAULT_BAD_STREAM, FAULT_BAD_INDEX = 1, 2, 4, 8
EPI_NONE, EPI_RELU, EPI_CLAMP01, EPI_QUANTIZE, EPI_ABS, EPI_LEAKY_RELU = 0, 1, 2, 3, 4, 5
IN_NONE, IN_ABS = 0, 1
TC_STORE_F16, TC_STORE_F32, TC_IGDN1_F16, TC_GDN1_F16, TC_STORE_ABS_F16, TC_IGDN1_ABS_F16 = 0, 1, 2, 3, 4, 5
TC_STORE_SQ_F16, TC_IGDN_SQ_F16, TC_NCHW_F32_CLAMP = 6, 7, 8
TCS_STORE, TCS_GDN1, TCS_QUANT, TCS_GDN = 0, 1, 2, 3
TCS_ACT_NONE, TCS_ACT_RELU, TCS_ACT_LEAKY = 0, 1, 2
RANS_LAYOUTS = {None: 0, 'auto': 0, 'warp': 1, 'lanes': 2}


def rans_layout(layout, n_streams):
    """`layout` argument of sc2_rans_*_batch.  'throughput' (what batches in flight ask for) resolves per call: a lane per stream when
    the batch fills at least one warp, else a warp per stream (one lane of 32 working would be the slowest choice)."""
    if layout == 'throughput':
        layout = 'lanes' if n_streams >= 32 else 'warp'
    return RANS_LAYOUTS[layout]

_lib = None


def lib_path():
    return _build.LIB_PATH


def load():
    """Loads (building first if needed) the native library; raises loudly when that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path) or os.environ.get('SC2B200_REBUILD'):
        path = _build.build_native(force=bool(os.environ.get('SC2B200_REBUILD')))
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.sc2_abi_version() != ABI_VERSION:
        raise ImportError('libsc2b200.so ABI version mismatch: rebuild with sc2-benchmark_b200/build.py --force')
    _lib = lib
    return lib


_hostbytes = None


def hostbytes():
    """The CPython helper module (csrc/hostbytes.c): list[bytes] <-> staging buffer with the GIL released."""
    global _hostbytes
    if _hostbytes is None:
        import importlib.util
        load()  # builds everything if needed
        path = _build.hostbytes_path()
        if not os.path.exists(path):
            _build.build_hostbytes()
        spec = importlib.util.spec_from_file_location('_sc2_hostbytes', path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _hostbytes = mod
    return _hostbytes


class NativeError(RuntimeError):
    pass


def check(rc, what):
    if rc != SC2_OK:
        lib = load()
        msg = lib.sc2_error_string(rc).decode()
        if rc == -3:
            msg += ': ' + lib.sc2_last_cuda_error().decode()
        raise NativeError('%s failed: %s' % (what, msg))
