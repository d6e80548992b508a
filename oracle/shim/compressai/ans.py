"""Restatement of the `compressai.ans` pybind module (SURVEY.md A.5, 8b "FFI actually underneath").
Call sites in CompressAI: EntropyModel.compress / decompress <- sc2bench/models/layer.py:506,520."""
import os
import sys

_ORACLE_DIR = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ORACLE_DIR not in sys.path:
    sys.path.insert(0, _ORACLE_DIR)
import cref  # noqa: E402  (oracle/cref.py)


class RansEncoder:
    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes, offsets):
        return cref.encode_with_indexes(symbols, indexes, _pad(cdfs), cdfs_sizes, offsets)


class RansDecoder:
    def decode_with_indexes(self, encoded, indexes, cdfs, cdfs_sizes, offsets):
        return cref.decode_with_indexes(encoded, indexes, _pad(cdfs), cdfs_sizes, offsets).tolist()


def _pad(cdfs):
    import numpy as np
    if isinstance(cdfs, np.ndarray):
        return cdfs
    width = max(len(r) for r in cdfs)
    out = np.zeros((len(cdfs), width), dtype=np.int32)
    for i, r in enumerate(cdfs):
        out[i, :len(r)] = r
    return out
