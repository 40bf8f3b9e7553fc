"""Host-side mirror of the consumers of ``<stem>_compare.tmp``.

The reference re-reads the text file once in ``SSG`` (amplicon_sorter.py:809-835), once in ``update_list``
(:986-1012) and once PER GROUP in ``read_indexes`` (:1364-1398), running a sort per line.  Here the file's lines
live on the GPU in integer form (``asb_lines_upload``) and the three scans become kernels:

    ssg_estimate     SSG                 histogram kernel + the reference's own float expressions on 1001 bins
    best_hits        the best-hit filter  key sort + one thread per key replaying the append/sort/drop loop
    make_groups      greedy grouping + merge_groups = connected components, numbered in first-seen order

``iden`` values have at most 3 decimals, so ``milli = iden * 1000`` is an exact integer stand-in:
``float(text) == milli / 1000`` and ``str(milli / 1000) == text`` for every value ``round(x, 3)`` can print
(checked in tests/test_groups.py).  Everything order-dependent in the reference (dictionary insertion order,
stable sorts, the leftovers of lower scores in the per-key lists) is reproduced, see lines.cuh.
"""
from __future__ import annotations

import numpy as np

IDEN_STR = [str(m / 1000) for m in range(1001)]  # the text of iden = m/1000 as the reference writes it


class Lines:
    """The lines of one ``_compare.tmp`` in file order: a = e[0], b = e[1] (idx values), milli = e[2] * 1000,
    rev = the line carries the ':reverse' tag (:798; read_indexes keeps it as a fourth field)."""

    def __init__(self, a, b, milli, rev=None):
        self.a = np.ascontiguousarray(a, dtype=np.uint32)
        self.b = np.ascontiguousarray(b, dtype=np.uint32)
        self.milli = np.ascontiguousarray(milli, dtype=np.uint32)
        self.rev = np.zeros(self.a.shape[0], dtype=bool) if rev is None else np.ascontiguousarray(rev, dtype=bool)
        assert self.a.shape == self.b.shape == self.milli.shape == self.rev.shape

    def __len__(self):
        return int(self.a.shape[0])

    @classmethod
    def from_text(cls, text: str):
        """Parse ``idxA:idxB:iden[:reverse]`` lines (used when the stage that wrote the file did not leave its arrays)."""
        rows = text.split("\n")
        if rows and rows[-1] == "":
            rows.pop()
        if not rows:
            z = np.zeros(0, dtype=np.uint32)
            return cls(z, z, z)
        rev = np.fromiter((r.endswith("e") for r in rows), dtype=bool, count=len(rows))  # '...:reverse'
        flat = text.replace(":reverse", "").replace("\n", ":").split(":")
        a = np.array(flat[0:3 * len(rows):3]).astype(np.uint32)
        b = np.array(flat[1:3 * len(rows):3]).astype(np.uint32)
        milli = np.rint(np.array(flat[2:3 * len(rows):3]).astype(np.float64) * 1000.0).astype(np.uint32)
        return cls(a, b, milli, rev)

    @classmethod
    def concat(cls, parts):
        parts = [p for p in parts if len(p)]
        if not parts:
            z = np.zeros(0, dtype=np.uint32)
            return cls(z, z, z)
        return cls(np.concatenate([p.a for p in parts]), np.concatenate([p.b for p in parts]),
                   np.concatenate([p.milli for p in parts]), np.concatenate([p.rev for p in parts]))


# lines left behind by host.process_list, keyed by the absolute path of the tempfile it wrote
CACHE: dict = {}


def lines_for(path: str) -> Lines:
    """The integer lines of the tempfile at `path`: from the stage that wrote it, else parsed from the text.
    Raises FileNotFoundError like the reference's open() when the file does not exist (:1014, :1392)."""
    import os

    key = os.path.abspath(path)
    if not os.path.exists(path):
        CACHE.pop(key, None)
        raise FileNotFoundError(path)
    got = CACHE.get(key)
    if got is not None and got[1] == os.path.getsize(path):
        return got[0]
    with open(path, "r") as f:
        lines = Lines.from_text(f.read())
    CACHE[key] = (lines, os.path.getsize(path))
    return lines


def upload(engine, lines: Lines):
    """Make `lines` the engine's resident line set (skipped when they already are)."""
    if getattr(engine, "_lines_token", None) is not lines:
        engine.lines_upload(lines.a, lines.b, lines.milli)
        engine._lines_token = lines


def ssg_estimate(engine, lines: Lines, stats: dict | None = None):
    """``SSG(tempfile)`` (:809-835) -> the estimated similar_species_groups value, or None when the walk ends
    without reaching the 6 % mark (the reference then falls off the end of the function)."""
    upload(engine, lines)
    hist, ms = engine.lines_hist()
    if stats is not None:
        stats["hist_ms"] = ms
    # totalsimil is a running float sum in file order (:826): add.accumulate adds in exactly that order
    totalsimil = float(np.add.accumulate(lines.milli.astype(np.float64) / 1000.0)[-1]) if len(lines) else 0
    b = int(totalsimil * 0.06)  # :829
    N6 = 0
    for m in range(1000, -1, -1):  # :827-828 keys sorted descending
        c = int(hist[m])
        if c == 0:
            continue
        x = m / 1000
        N6 += c * x  # :832
        if N6 >= b:
            return int(x * 100)  # :834-835
    return None


def member_bitmap(indexes, n_idx: int) -> np.ndarray:
    """Bitmap over idx values of the group's members (read_indexes' `indexes` set of strings, :1346-1349)."""
    words = np.zeros((n_idx + 32) // 32, dtype=np.uint32)
    ids = np.fromiter((int(x) for x in indexes if x.isdigit() and int(x) < n_idx), dtype=np.int64)
    if ids.size:
        np.bitwise_or.at(words, ids >> 5, (np.uint32(1) << (ids & 31).astype(np.uint32)))
    return words


def best_hits(engine, lines: Lines, ssg=None, indexes=None, stats: dict | None = None):
    """The best-hit filter and the sort that follows it.

    ssg is None : update_list (:986-1012) -- every line; result sorted by (score, int(idx of the longer read))
                  descending, stable.
    otherwise   : read_indexes (:1364-1398) -- lines with float(iden) >= ssg (a fraction) touching `indexes`;
                  result sorted by score descending, stable over dictionary order.
    Returns (templist as the reference builds it: [idxA, idxB, iden] string triples, a, b, milli arrays of it)."""
    upload(engine, lines)
    if ssg is None:
        min_milli, member = 0, None
    else:
        # float(e[2]) >= ssg  <=>  milli/1000 >= ssg: smallest milli whose float passes, found with the same compare
        min_milli = int(np.searchsorted(np.arange(1001, dtype=np.float64) / 1000.0 >= ssg, True))
        n_idx = int(max(lines.a.max(initial=0), lines.b.max(initial=0))) + 1 if len(lines) else 1
        member = member_bitmap(indexes, n_idx)
    line, first, ms = engine.lines_besthit(min_milli, member)
    if stats is not None:
        stats["besthit_ms"] = ms
        stats["survivors"] = int(line.shape[0])
    # flatten order: dictionary insertion order of the keys (first admitted line), then the key's list order
    line = line[np.lexsort((np.arange(line.shape[0]), first))]
    mi, bi = lines.milli[line].astype(np.int64), lines.b[line].astype(np.int64)
    line = line[np.lexsort((-bi, -mi)) if ssg is None else np.argsort(-mi, kind="stable")]  # reverse=True keeps ties in order
    a, b, m = lines.a[line], lines.b[line], lines.milli[line]
    templist = [[str(x), str(y), IDEN_STR[z]] for x, y, z in zip(a.tolist(), b.tolist(), m.tolist())]
    if ssg is not None:  # read_indexes does not cut the line to three fields (:1366)
        for e, r in zip(templist, lines.rev[line].tolist()):
            if r:
                e.append("reverse")
    return templist, a, b, m


def make_groups(engine, a, b, stats: dict | None = None):
    """Greedy grouping (:1022-1031 / :1403-1409) followed by merge_groups (:1057-1086) on the templist edges
    (a[i], b[i]) in templist order.  Their fixed point is the set of connected components; merge_groups keeps the
    union at the smaller position, so components are numbered by their first appearance in the templist.
    Returns (number of groups the greedy pass creates = the 'before merge' count, list of sets of idx strings)."""
    a = np.ascontiguousarray(a, dtype=np.uint32)
    b = np.ascontiguousarray(b, dtype=np.uint32)
    if a.shape[0] == 0:
        return 0, []
    n_nodes = int(max(a.max(), b.max())) + 1
    label, ms = engine.components(a, b, n_nodes)
    if stats is not None:
        stats["components_ms"] = ms
    # greedy creates a group exactly when neither end of an edge has been seen before
    seq = np.stack([a, b], axis=1).reshape(-1).astype(np.int64)
    first_seen = np.full(n_nodes, seq.shape[0], dtype=np.int64)
    np.minimum.at(first_seen, seq, np.arange(seq.shape[0]))
    e = np.arange(a.shape[0], dtype=np.int64)
    n_greedy = int(np.count_nonzero((first_seen[a] >= 2 * e) & (first_seen[b] >= 2 * e)))
    # components in order of first appearance; members = the nodes that occur in an edge
    lab = label[a].astype(np.int64)
    roots, first_edge = np.unique(lab, return_index=True)
    roots = roots[np.argsort(first_edge, kind="stable")]
    nodes = np.unique(seq)
    node_lab = label[nodes].astype(np.int64)
    by = np.argsort(node_lab, kind="stable")
    nodes, node_lab = nodes[by], node_lab[by]
    starts = np.searchsorted(node_lab, roots, side="left")
    ends = np.searchsorted(node_lab, roots, side="right")
    groups = [set(str(v) for v in nodes[s:t].tolist()) for s, t in zip(starts.tolist(), ends.tolist())]
    return n_greedy, groups
