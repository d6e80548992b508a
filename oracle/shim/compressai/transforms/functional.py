"""Import-only stand-ins (sc2bench/transforms/codec.py:6; off the hot path)."""


def rgb2ycbcr(rgb):
    raise NotImplementedError('off the bottleneck path')


def ycbcr2rgb(ycbcr):
    raise NotImplementedError('off the bottleneck path')
