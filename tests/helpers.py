"""Shared test helpers (TEST side: may use the oracle)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def state_dict_from_golden(g):
    return {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith('sd/')}


def unpack_streams(data, offsets):
    raw = data.tobytes()
    return [raw[offsets[i]:offsets[i + 1]] for i in range(len(offsets) - 1)]


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
