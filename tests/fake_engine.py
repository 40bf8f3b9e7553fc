"""The CPU oracle behind the Engine interface -- lets CPU tests drive the product's HOST code
(host.process_list, launcher) end to end without a GPU.  Test infrastructure, never shipped."""
import numpy as np

from amplicon_sorter_b200._ffi import RECORD
from amplicon_sorter_b200.engine import EngineBase, TextChunk
from oracle import oracle


class OracleEngine(EngineBase):
    def __init__(self):
        self.closed = False
        self._lines_token = None
        self._res = [np.zeros(0, np.uint32)] * 3 + [np.zeros(0, bool)]

    # batch_begin / batch_step / text_begin / text_step: the stepping interface of Engine, one slab per batch
    def batch_begin(self, order, hi, dpass, drev, rank=0, world=1):
        self._batch = (np.asarray(order, np.uint32), np.asarray(hi, np.uint32), np.asarray(dpass, np.uint32), np.asarray(drev, np.uint32), rank, world)
        self._stepped = False

    def batch_step(self):
        if self._stepped:
            return None
        self._stepped = True
        self._recs, tot = self.compare_batch(*self._batch)
        return dict(tot, launches=0, screen_word_updates=0)

    def batch_records(self, n):
        return self._recs

    def step_records_tensor(self, n_records, dev):
        import torch

        return torch.from_numpy(self._recs.view(np.uint32).reshape(-1, 4).view(np.int32).copy()).to(dev)

    def text_begin(self, idx_sorted, lbase, soff, milli, sbuf):
        self._tabs = (np.asarray(idx_sorted), np.asarray(lbase), np.asarray(soff), np.asarray(milli), bytes(sbuf))
        self._lens_sorted = (self.offs[1:] - self.offs[:-1]).astype(np.int64)[self._batch[0]]
        self._res = [np.zeros(0, np.uint32)] * 3 + [np.zeros(0, bool)]
        self._lines_token = None

    def text_load(self, dev_ptr, n_records, sort=True):
        assert dev_ptr is None
        self._staged = self._recs

    def text_load_tensor(self, recs, sort=True):
        r = recs.cpu().numpy().view(np.uint32).reshape(-1).view(RECORD)
        if sort:
            key = r["i_pos"].astype(np.uint64) << np.uint64(32) | r["j_pos"].astype(np.uint64)
            r = r[np.argsort(key, kind="stable")]
        self._staged = r
        return int(r.shape[0])

    def text_chunks(self, n_records, append_lines=True):
        return [self._format(self._staged, append_lines)] if n_records else []

    def text_measure(self):
        return len(self._format(self._staged, False).data)

    def lines_append_tensor(self, recs):
        r = recs.cpu().numpy().view(np.uint32).reshape(-1).view(RECORD)
        if r.shape[0]:
            self._format(r, True)

    def _format(self, recs, append_lines=True):
        idx, lbase, soff, milli, sbuf = self._tabs
        out, a, b, m, rv = [], [], [], [], []
        for i, j, d, rev in recs.tolist():
            e = int(lbase[self._lens_sorted[j]]) + d
            assert lbase[self._lens_sorted[j]] != 0xFFFFFFFF and e < milli.shape[0]
            out.append(b"%d:%d:%s%s\n" % (idx[i], idx[j], sbuf[soff[e]:soff[e + 1]], b":reverse" if rev else b""))
            a.append(idx[i]); b.append(idx[j]); m.append(milli[e]); rv.append(bool(rev))
        if append_lines:
            self._res = [np.concatenate([x, np.asarray(y, dtype=x.dtype)]) for x, y in zip(self._res, (a, b, m, rv))]
            self._lines = tuple(self._res[:3])  # the printed lines are the resident line set
        return TextChunk(b"".join(out))

    def lines_count(self):
        return int(self._res[0].shape[0])

    def lines_fetch(self):
        return tuple(self._res)

    def upload_reads(self, buf, offs):
        self.buf = np.array(buf, dtype=np.uint8)    # "buffers are caller-owned and copied during the call" (asb200.h)
        self.offs = np.array(offs, dtype=np.uint64)

    # the scattered upload of the product engine: lets the CPU tests drive host.process_list's record walker path
    # (pyhost.collect -> pointers into the records' own str objects) end to end
    scattered_ok = True

    def upload_reads_scattered(self, ptrs, lens):
        import ctypes

        lens = np.asarray(lens, dtype=np.uint64)
        offs = np.zeros(lens.shape[0] + 1, dtype=np.uint64)
        np.cumsum(lens, out=offs[1:])
        blob = b"".join(ctypes.string_at(int(p), int(n)) for p, n in zip(np.asarray(ptrs).tolist(), lens.tolist()))
        self.upload_reads(np.frombuffer(blob, dtype=np.uint8), offs)
        self.scattered_uploads = getattr(self, "scattered_uploads", 0) + 1

    def set_param(self, *a):
        pass

    def compare_batch(self, order, hi, dpass, drev, rank=0, world=1, fetch=True):
        """Same contract as Engine.compare_batch, decisions taken with the integer tables so the
        host-built thresholds are what is being exercised."""
        order = np.asarray(order, dtype=np.uint32)
        n = order.shape[0]
        lens = (self.offs[1:] - self.offs[:-1]).astype(np.int64)[order]
        rows = []
        pairs = 0
        g = 0
        for i in range(n):
            cnt = int(hi[i]) - i
            if cnt <= 0:
                continue
            if (i * world) // max(n, 1) != rank:  # contiguous pieces in rank order, like the engine's (any such split will do)
                continue
            js = np.arange(i + 1, int(hi[i]) + 1)
            pairs += js.size
            a = np.full(js.size, order[i], dtype=np.uint32)
            b = order[js]
            d = oracle.distance_pairs(self.buf, self.offs, a, b)
            L = lens[js]
            fwd = d <= dpass[L].astype(np.int64)
            need = (~fwd) & (d >= drev[L].astype(np.int64))
            for t in np.nonzero(fwd)[0]:
                rows.append((i, int(js[t]), int(d[t]), 0))
            for t in np.nonzero(need)[0]:
                x = self.buf[self.offs[order[i]]:self.offs[order[i] + 1]].tobytes()
                y = self.buf[self.offs[b[t]]:self.offs[b[t] + 1]].tobytes()
                dr = oracle.nw(x, oracle.compl_reverse(y), "myers")
                if dr <= int(dpass[L[t]]):
                    rows.append((i, int(js[t]), dr, 1))
        rows.sort()
        recs = np.array(rows, dtype=RECORD) if rows else np.empty(0, dtype=RECORD)
        tot = {"pairs": pairs, "n_records": len(rows), "fwd_survivors": 0, "rc_survivors": 0, "zone_checks": 0,
               "word_updates": 0, "screen_ms": 0.0, "total_ms": 0.0, "steps": 1}
        return recs, tot

    def threeway_pairs(self, q, t, dpass, drev):
        q = np.asarray(q, dtype=np.uint32)
        t = np.asarray(t, dtype=np.uint32)
        lens = (self.offs[1:] - self.offs[:-1]).astype(np.int64)
        d = oracle.distance_pairs(self.buf, self.offs, q, t)
        L = np.maximum(lens[q], lens[t])
        rows = []
        for p in range(q.shape[0]):
            if d[p] <= int(dpass[L[p]]):
                rows.append((int(q[p]), int(t[p]), int(d[p]), 0))
            elif d[p] >= int(drev[L[p]]):
                x = self.buf[self.offs[q[p]]:self.offs[q[p] + 1]].tobytes()
                y = self.buf[self.offs[t[p]]:self.offs[t[p] + 1]].tobytes()
                if len(x) > len(y):
                    x, y = y, x
                dr = oracle.nw(x, oracle.compl_reverse(y), "myers")  # compl_reverse of either side gives the same distance
                if dr <= int(dpass[L[p]]):
                    rows.append((int(q[p]), int(t[p]), dr, 1))
        rows.sort()
        recs = np.array(rows, dtype=RECORD) if rows else np.empty(0, dtype=RECORD)
        return recs, {"pairs": int(q.shape[0]), "n_records": len(rows)}

    def distance_pairs(self, a, b, strand=None, mode="NW"):
        out = np.empty(len(a), dtype=np.int32)
        for p in range(len(a)):
            x = self.buf[self.offs[a[p]]:self.offs[a[p] + 1]].tobytes()
            y = self.buf[self.offs[b[p]]:self.offs[b[p] + 1]].tobytes()
            if len(x) > len(y):
                x, y = y, x
            if strand is not None and strand[p]:
                y = oracle.compl_reverse(y)
            out[p] = oracle.hw(x, y) if mode == "HW" else oracle.nw(x, y, "myers")
        return out

    # consumers of <stem>_compare.tmp: the C restatements of oracle/asref.c behind the Engine methods
    def lines_upload(self, a, b, milli):
        tok, self._lines_token = self._lines_token, None
        if tok is not None and hasattr(tok, "materialize"):
            tok.materialize()
        self._lines = (np.asarray(a, np.uint32), np.asarray(b, np.uint32), np.asarray(milli, np.uint32))

    def lines_hist(self):
        return np.bincount(self._lines[2], minlength=1001).astype(np.uint64), 0.0

    def lines_besthit(self, min_milli=0, member_bits=None):
        line, first = oracle.besthit(*self._lines, min_milli=min_milli, member_bits=member_bits)
        return line, first, 0.0

    def components(self, a, b, n_nodes):
        return oracle.components(a, b, n_nodes), 0.0

    def close(self):
        self.closed = True
