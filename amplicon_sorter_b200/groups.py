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

import os

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

    def discard(self):
        """The file these lines belong to is gone: nobody will ask for them again."""

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


class DeviceLines(Lines):
    """The line set host.process_list left RESIDENT on `engine`'s GPU (csrc/text.cuh appends the integer form of every
    line it prints).  Host copies are taken only when somebody asks for them -- or when the engine is about to replace
    its resident set (Engine._evict_lines)."""

    def __init__(self, engine):
        self._engine = engine
        self._n = int(engine.lines_count())
        self._host = None
        engine._lines_token = self

    def materialize(self):
        if self._host is None:
            self._host = self._engine.lines_fetch()
            self._engine = None
        return self._host

    a = property(lambda self: self.materialize()[0])
    b = property(lambda self: self.materialize()[1])
    milli = property(lambda self: self.materialize()[2])
    rev = property(lambda self: self.materialize()[3])

    def __len__(self):
        return self._n

    def discard(self):
        eng, self._engine = self._engine, None
        if eng is not None and getattr(eng, "_lines_token", None) is self:
            eng._lines_token = None  # no host copy needed when the engine replaces its resident set


# lines left behind by host.process_list, keyed by the absolute path of the tempfile it wrote (one file at a time)
CACHE: dict = {}


def lines_for(path: str) -> Lines:
    """The integer lines of the tempfile at `path`: from the stage that wrote it, else parsed from the text.
    Raises FileNotFoundError like the reference's open() when the file does not exist (:1014, :1392)."""
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


def make_groups(engine, templist, update_with_list: bool = False, stats: dict | None = None):
    """Greedy grouping (:1022-1031 / :1403-1409) followed by merge_groups (:1057-1086) on the templist.

    The fixed point of the two is the set of connected components of the templist edges, numbered by first
    appearance; the reference gets there with O(groups^2) set intersections, almost all of them between groups of
    DIFFERENT components, which can never intersect.  The GPU labels the components (asb_components); the host then
    replays the reference's own set operations -- the same update()/union()/clear() calls on real Python sets, in
    the same order per set -- only inside each component.  The returned sets therefore have the reference's
    partition, numbering AND internal layout (iteration order under the same PYTHONHASHSEED), which the later
    stages depend on (list(set(i)) + random.sample at :1412/:1436).
    update_with_list: read_indexes grows a group with s.update([x0, x1]) (:1406), update_list with s.update({x0, x1}).
    Returns (number of groups before merge, merged groups)."""
    if not templist:
        return 0, []
    a = np.fromiter((int(x[0]) for x in templist), dtype=np.uint32, count=len(templist))
    b = np.fromiter((int(x[1]) for x in templist), dtype=np.uint32, count=len(templist))
    n_nodes = int(max(a.max(), b.max())) + 1
    label, ms = engine.components(a, b, n_nodes)
    if stats is not None:
        stats["components_ms"] = ms
    # greedy pass: "the first group that holds x0 or x1" = the lower of the two nodes' first groups
    grouplist, first = [], {}
    big = len(templist) + 1
    for x in templist:
        x0, x1 = x[0], x[1]
        g0, g1 = first.get(x0, big), first.get(x1, big)
        g = g0 if g0 < g1 else g1
        if g == big:
            g = len(grouplist)
            grouplist.append({x0, x1})
        elif update_with_list:
            grouplist[g].update([x0, x1])
        else:
            grouplist[g].update({x0, x1})
        if g < g0:
            first[x0] = g
        if g < g1:
            first[x1] = g
    n_greedy = len(grouplist)
    if n_greedy <= 1:  # merge_groups leaves a single group alone (:1060)
        return n_greedy, grouplist
    grouplist = [i for i in grouplist if len(i) > 1]  # :1063
    grouplist = [set(i) for i in grouplist]           # :1064
    buckets: dict = {}
    for p, g in enumerate(grouplist):
        buckets.setdefault(int(label[int(next(iter(g)))]), []).append(p)
    for positions in buckets.values():
        merged_any = len(positions) > 1
        while merged_any:  # :1065 `while a1 > a2`: repeat while a pass merged something
            merged_any = False
            for i in range(len(positions) - 1):
                for j in range(i + 1, len(positions)):
                    pi, pj = positions[i], positions[j]
                    if not grouplist[pi].isdisjoint(grouplist[pj]):   # :1076
                        grouplist[pi] = grouplist[pi].union(grouplist[pj])  # :1077
                        grouplist[pj].clear()                                # :1078
                        merged_any = True
            positions[:] = [p for p in positions if len(grouplist[p]) > 0]
    grouplist = [i for i in grouplist if len(i) > 0]  # :1081
    return n_greedy, grouplist
