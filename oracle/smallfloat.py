"""Lucene SmallFloat.intToByte4 / byte4ToInt restated (test infrastructure).

Third-party (Lucene 9.7, org.apache.lucene.util.SmallFloat), not vendored under
/root/reference -> UNPINNED; restated from the published algorithm quoted in
SURVEY.md section 8c.  BM25Similarity stores the per-document field length as
this byte ("norm") and scores with the DECODED length, so doc lengths above 39
are quantised (4 significant bits, rounding down).
"""
from __future__ import annotations

import numpy as np


def _long_to_int4(v: int) -> int:
    if v < 0:
        raise ValueError("only non-negative values are supported")
    nb = v.bit_length()  # 64 - numberOfLeadingZeros
    if nb < 4:
        return v
    shift = nb - 4
    enc = (v >> shift) & 0x07  # drop the implicit leading 1
    enc |= (shift + 1) << 3
    return enc


def _int4_to_long(e: int) -> int:
    bits = e & 0x07
    shift = (e >> 3) - 1
    if shift == -1:
        return bits
    return (bits | 0x08) << shift


_NUM_FREE = 255 - _long_to_int4(2 ** 31 - 1)  # = 24


def int_to_byte4(i: int) -> int:
    if i < 0:
        raise ValueError("only non-negative values are supported")
    if i < _NUM_FREE:
        return i
    return _NUM_FREE + _long_to_int4(i - _NUM_FREE)


def byte4_to_int(b: int) -> int:
    b &= 0xFF
    if b < _NUM_FREE:
        return b
    return _NUM_FREE + _int4_to_long(b - _NUM_FREE)


_DECODED = np.array([byte4_to_int(b) for b in range(256)], dtype=np.int64)
LENGTH_TABLE = _DECODED.astype(np.float32)       # what BM25Similarity multiplies with (float)


def encode_lengths(lengths: np.ndarray) -> np.ndarray:
    """Vectorised int_to_byte4 for an int array of token counts -> uint8 norms."""
    lengths = np.asarray(lengths, dtype=np.int64)
    if lengths.size and lengths.min() < 0:
        raise ValueError("negative length")
    # the decoded values are strictly increasing: largest byte whose decoded value <= length (the integer table, not
    # its float image, which rounds above 2^24)
    return (np.searchsorted(_DECODED, lengths, side="right") - 1).astype(np.uint8)
