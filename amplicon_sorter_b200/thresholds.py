"""Integer cut-off tables for the three-way rule of ``similarity`` (amplicon_sorter.py:776-807).

The reference never compares distances; it compares ``iden = round(1 - d/len(longer), 3)``
(amplicon_sorter.py:233) with ``similarg = args.similar_genes/100`` (:783, :791, :796) and with the
literal ``0.5`` (:794).  ``iden`` is non-increasing in ``d`` for a fixed length, so both tests are
equivalent to integer tests against per-length cut-offs.  The tables are built HERE, with the
reference's own Python float expressions, and shipped to the GPU as integers -- the device never
re-derives a threshold (rounding moves them: L=1024, sg=0.80 gives dpass=205 although
205/1024 > 0.2).
"""
from __future__ import annotations

import numpy as np


def iden(d: int, len_long: int) -> float:
    """amplicon_sorter.py:233 verbatim arithmetic."""
    return round(1 - d / len_long, 3)


def dpass_for(L: int, cut: float) -> int:
    """max{d in [0, L] : round(1 - d/L, 3) >= cut}, or -1 if no d qualifies."""
    if L <= 0:
        return -1
    if not iden(0, L) >= cut:
        return -1
    lo, hi = 0, L  # iden(lo) >= cut ; find last d with iden(d) >= cut
    if iden(L, L) >= cut:
        return L
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if iden(mid, L) >= cut:
            lo = mid
        else:
            hi = mid
    return lo


def drev_for(L: int, below: float = 0.5) -> int:
    """min{d in [0, L] : round(1 - d/L, 3) < below}; L + 1 if no d qualifies."""
    if L <= 0:
        return 1
    if iden(0, L) < below:
        return 0
    if not iden(L, L) < below:
        return L + 1
    lo, hi = 0, L  # iden(lo) >= below, iden(hi) < below
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if iden(mid, L) < below:
            hi = mid
        else:
            lo = mid
    return hi


_CACHE: dict = {}


def tables(cut: float, max_len: int, below: float = 0.5):
    """(dpass, drev) as uint32 arrays indexed by the length of the LONGER read, 0..max_len.

    dpass[L] is stored +1 biased?  No: it is stored as int32-compatible uint32 where a length with no
    passing distance gets 0xFFFFFFFF (never happens for cut <= 1.0).
    """
    key = (float(cut), int(max_len), float(below))
    hit = _CACHE.get(key)
    if hit is not None:
        return hit
    dp = np.empty(max_len + 1, dtype=np.uint32)
    dr = np.empty(max_len + 1, dtype=np.uint32)
    for L in range(max_len + 1):
        p = dpass_for(L, cut)
        dp[L] = 0xFFFFFFFF if p < 0 else p
        dr[L] = drev_for(L, below)
    if len(_CACHE) > 64:
        _CACHE.clear()
    _CACHE[key] = (dp, dr)
    return dp, dr
