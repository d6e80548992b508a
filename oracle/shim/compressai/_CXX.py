"""Restatement of `compressai._CXX.pmf_to_quantized_cdf`."""
from . import ans as _ans  # sets sys.path for cref

import cref


def pmf_to_quantized_cdf(pmf, precision):
    return cref.pmf_to_quantized_cdf(pmf, precision).tolist()
