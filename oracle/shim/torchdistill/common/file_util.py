import pickle
import sys


def get_binary_object_size(x, unit_size=1024):
    """pickled size / unit (consumed by FileSizeAnalyzer, sc2bench/analysis.py:133)."""
    return sys.getsizeof(pickle.dumps(x)) / unit_size
