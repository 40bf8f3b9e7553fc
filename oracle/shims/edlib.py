"""Minimal stand-in for the PyPI package `edlib` (not installed here).

TEST INFRASTRUCTURE ONLY.  Implements exactly the surface amplicon_sorter.py uses:
``align(query, target, mode='NW'|'HW', task='distance'|'path', k=-1, additionalEqualities=None)``
returning ``{'editDistance', 'alphabetLength', 'locations', 'cigar'}``.  Distances come from
oracle/asref.c (mathematical definitions).  The CIGAR tie-breaking for task='path' is this
shim's own; real edlib's is not derivable from the reference (parity unpinned for consensus text).
"""
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as _o  # noqa: E402


def _as_bytes(x):
    if isinstance(x, (bytes, bytearray)):
        return bytes(x)
    if isinstance(x, str):
        return x.encode("latin-1")
    return "".join(x).encode("latin-1")


def align(query, target, mode="NW", task="distance", k=-1, additionalEqualities=None):
    q, t = _as_bytes(query), _as_bytes(target)
    res = {"editDistance": -1, "alphabetLength": len(set(q) | set(t)), "locations": [(None, None)], "cigar": None}
    if task == "path" or additionalEqualities:
        if mode != "NW":
            raise NotImplementedError("shim: path/equalities only for NW")
        eq = bytearray(256 * 256)
        for c in range(256):
            eq[c * 256 + c] = 1
        for a, b in additionalEqualities or []:
            eq[ord(a) * 256 + ord(b)] = 1
            eq[ord(b) * 256 + ord(a)] = 1
        ops, d = _o.nw_path(q, t, bytes(eq))
        res["editDistance"] = d
        res["locations"] = [(0, len(t) - 1)]
        if task == "path":
            res["cigar"] = "".join(f"{len(m.group(0))}{m.group(0)[0]}" for m in re.finditer(r"M+|I+|D+", ops))
    elif mode == "NW":
        res["editDistance"] = _o.lib().asref_myers_nw(q, len(q), t, len(t)) if q and t else max(len(q), len(t))
        res["locations"] = [(0, len(t) - 1)]
    elif mode == "HW":
        res["editDistance"] = _o.lib().asref_dp_hw(q, len(q), t, len(t))
    else:
        raise NotImplementedError(mode)
    if k >= 0 and res["editDistance"] > k:
        res["editDistance"] = -1
    return res
