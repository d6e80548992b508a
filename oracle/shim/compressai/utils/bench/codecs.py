"""Import-only stand-in (sc2bench/transforms/codec.py:7; off the hot path)."""


def run_command(cmd, ignore_returncodes=None):
    raise NotImplementedError('off the bottleneck path')
