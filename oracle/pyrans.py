"""oracle/pyrans.py -- pure-Python (arbitrary precision int) restatement of the CompressAI rANS coder.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  PARITY UNPINNED: restated from the published
algorithm of compressai/cpp_exts/rans/rans_interface.cpp + ryg_rans rans64.h;
reached in the reference from sc2bench/models/layer.py:506,520.  Deliberately written as
plain loops, independent of oracle/rans_oracle.c, so the two restatements check each other.
Use only for small cases (it is ~1 us/bit slow).
"""
import struct

PRECISION = 16
BYPASS_PRECISION = 4
MAX_BYPASS_VAL = (1 << BYPASS_PRECISION) - 1
RANS64_L = 1 << 31
_MASK64 = (1 << 64) - 1


def _entries(symbols, indexes, cdfs, cdf_sizes, offsets):
    """Forward pass of encode_with_indexes: (start, range, is_bypass) triples."""
    out = []
    for sym, row in zip(symbols, indexes):
        cdf = cdfs[row]
        max_value = cdf_sizes[row] - 2
        value = sym - offsets[row]
        raw = 0
        if value < 0:
            raw = -2 * value - 1
            value = max_value
        elif value >= max_value:
            raw = 2 * (value - max_value)
            value = max_value
        out.append((cdf[value], cdf[value + 1] - cdf[value], False))
        if value == max_value:
            n_bypass = 0
            while (raw >> (n_bypass * BYPASS_PRECISION)) != 0:
                n_bypass += 1
            val = n_bypass
            while val >= MAX_BYPASS_VAL:
                out.append((MAX_BYPASS_VAL, MAX_BYPASS_VAL + 1, True))
                val -= MAX_BYPASS_VAL
            out.append((val, val + 1, True))
            for j in range(n_bypass):
                v = (raw >> (j * BYPASS_PRECISION)) & MAX_BYPASS_VAL
                out.append((v, v + 1, True))
    return out


def encode_with_indexes(symbols, indexes, cdfs, cdf_sizes, offsets):
    """list[int] x5 -> bytes, like compressai.ans.RansEncoder().encode_with_indexes."""
    entries = _entries(symbols, indexes, cdfs, cdf_sizes, offsets)
    words = []  # emitted back to front
    x = RANS64_L
    for start, rng, bypass in reversed(entries):
        if not bypass:
            x_max = ((RANS64_L >> PRECISION) << 32) * rng
            if x >= x_max:
                words.append(x & 0xFFFFFFFF)
                x >>= 32
            x = ((x // rng) << PRECISION) + (x % rng) + start
        else:
            freq = 1 << (PRECISION - BYPASS_PRECISION)
            x_max = ((RANS64_L >> PRECISION) << 32) * freq
            if x >= x_max:
                words.append(x & 0xFFFFFFFF)
                x >>= 32
            x = (x << BYPASS_PRECISION) | start
        assert x <= _MASK64
    words.append((x >> 32) & 0xFFFFFFFF)
    words.append(x & 0xFFFFFFFF)
    words.reverse()
    return struct.pack('<%dI' % len(words), *words)


def decode_with_indexes(stream, indexes, cdfs, cdf_sizes, offsets):
    """bytes -> list[int], like compressai.ans.RansDecoder().decode_with_indexes."""
    words = struct.unpack('<%dI' % (len(stream) // 4), stream)
    pos = 2
    x = words[0] | (words[1] << 32)

    def getbits(nbits):
        nonlocal x, pos
        val = x & ((1 << nbits) - 1)
        x >>= nbits
        if x < RANS64_L:
            x = (x << 32) | words[pos]
            pos += 1
        return val

    out = []
    for row in indexes:
        cdf = cdfs[row]
        size = cdf_sizes[row]
        max_value = size - 2
        cum = x & 0xFFFF
        k = 0
        while k < size and not cdf[k] > cum:
            k += 1
        s = k - 1
        x = (cdf[s + 1] - cdf[s]) * (x >> PRECISION) + (x & 0xFFFF) - cdf[s]
        if x < RANS64_L:
            x = (x << 32) | words[pos]
            pos += 1
        value = s
        if value == max_value:
            val = getbits(BYPASS_PRECISION)
            n_bypass = val
            while val == MAX_BYPASS_VAL:
                val = getbits(BYPASS_PRECISION)
                n_bypass += val
            raw = 0
            for j in range(n_bypass):
                raw |= getbits(BYPASS_PRECISION) << (j * BYPASS_PRECISION)
            value = raw >> 1
            value = -value - 1 if raw & 1 else value + max_value
        out.append(value + offsets[row])
    return out
