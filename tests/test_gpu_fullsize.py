"""GPU parity at BASELINE.json's FULL sizes.  The oracle cannot decide 5e9 pairs (36 core-hours), so
full-size runs are checked through size-independent properties plus exact oracle parity on sampled rows:

  * every record is inside the reference's pair set (i < j <= hi[i]) and in the reference's line order;
  * every record's distance clears the integer cut-off of its longer read;
  * for a strided sample of ROWS the records are exactly the oracle's (all partners of those rows, both strands);
  * the run is idempotent, and a 2-way sharded run partitions the pair set and merges to the same records.
"""
import zlib

import numpy as np
import pytest

from amplicon_sorter_b200 import synth
from oracle import oracle
from tests import util

pytestmark = pytest.mark.gpu


def checksum(recs):
    return zlib.crc32(np.ascontiguousarray(recs).view(np.uint8).tobytes())


def rows_of(recs, rows):
    return recs[np.isin(recs["i_pos"], rows)]


@pytest.mark.parametrize("cfg,row_stride", [(5, 4999), (4, 2503), (2, 251)])
def test_full_size_config(engine, cfg, row_stride):
    reads, _, _ = synth.make_config(cfg)  # full size: 100,000 / 50,000 / 10,000 reads
    buf, offs, order, lens_sorted, hi, dpass, drev = util.batch_inputs(reads)
    n = len(reads)
    tl = int((hi.astype(np.int64) - np.arange(n)).sum())
    engine.upload_reads(buf, offs)
    recs, tot = engine.compare_batch(order, hi, dpass, drev)
    assert tot["pairs"] == tl
    # structure: inside the pair set, strictly increasing (i, j) = the reference's -np 1 line order
    i, j = recs["i_pos"].astype(np.int64), recs["j_pos"].astype(np.int64)
    assert np.all(i < j) and np.all(j <= hi[i])
    key = (i << 32) | j
    assert np.all(np.diff(key) > 0)
    assert np.all(recs["d"] <= dpass[lens_sorted[j]])
    assert set(np.unique(recs["reverse"]).tolist()) <= {0, 1} and (recs["reverse"] == 1).any() and (recs["reverse"] == 0).any()
    # exact parity on sampled rows (the oracle decides every partner of these rows)
    sample_rows = np.arange(row_stride // 2, n, row_stride)
    want, st = oracle.process_batch(buf, offs, order, 80.0, rows=(row_stride // 2, n, row_stride))
    util.assert_same_records(rows_of(recs, sample_rows), want)
    assert st["pairs"] > 10000 and len(want) > 0
    # idempotence + sharding invariants (cheap on the smaller configs, one extra job on config 5)
    crc = checksum(recs)
    if cfg != 5:
        again, _ = engine.compare_batch(order, hi, dpass, drev)
        assert checksum(again) == crc
    parts, pairs = [], 0
    for rank in range(2):
        r, t = engine.compare_batch(order, hi, dpass, drev, rank, 2)
        parts.append(r)
        pairs += t["pairs"]
    assert pairs == tl
    merged = np.concatenate(parts)
    mk = merged["i_pos"].astype(np.uint64) << np.uint64(32) | merged["j_pos"].astype(np.uint64)
    merged = merged[np.argsort(mk, kind="stable")]
    assert checksum(merged) == crc


@pytest.mark.parametrize("cfg", [1, 3])
def test_full_size_batched_configs(engine, cfg):
    """BASELINE configs 1 and 3 at full size: default mode (one batch of 1,000 + nine empty ones) and -ra (20 overlapping
    random batches of 1,000 out of 10,000 reads), every batch laid end to end as ONE engine batch
    (host.AllPairs.plan) -- against the oracle run batch by batch, every pair."""
    import bench
    from amplicon_sorter_b200 import host, thresholds

    reads, _, _ = synth.make_config(cfg)
    buf, offs = synth.pack_reads(reads)
    lens = (offs[1:] - offs[:-1]).astype(np.int64)
    batches = [np.asarray(b, dtype=np.int64) for b in bench.batches_of(cfg, len(reads)) if len(b)]
    ap = host.AllPairs(engine)
    ap.lens = lens
    perms, order, lens_sorted, hi, tl = ap.plan(batches)
    dpass, drev = thresholds.tables(0.80, int(lens.max()) + 1)
    engine.upload_reads(buf, offs)
    got, tot = engine.compare_batch(order, hi, dpass, drev)
    assert tot["pairs"] == tl
    want, base, pairs = [], 0, 0
    for b in batches:
        o = b[np.argsort(lens[b], kind="stable")].astype(np.uint32)
        recs, st = oracle.process_batch(buf, offs, o, 80.0)
        recs = recs.copy()
        recs["i_pos"] += base
        recs["j_pos"] += base
        want.append(recs)
        pairs += st["pairs"]
        base += len(b)
    assert pairs == tl and len(batches) == (1 if cfg == 1 else 20)
    util.assert_same_records(got, np.concatenate(want))
    # the file host.process_list writes for these batches has the CRC the bench checks
    text = host.format_records(got, order.astype(np.int64), lens_sorted, dpass)
    assert zlib.crc32(text.encode()) == bench.KNOWN_TEXT_CRC[cfg]


def test_full_size_text_crc_matches_host_formatter(engine):
    """cfg5 at full size: the device-assembled tempfile (CRC pinned in bench.KNOWN_TEXT_CRC, checked by every bench run
    at any GPU count) equals the host-side formatter applied to the records -- whose sampled rows the test above
    compares with the oracle."""
    import bench
    from amplicon_sorter_b200 import host

    reads, _, _ = synth.make_config(5)
    buf, offs, order, lens_sorted, hi, dpass, drev = util.batch_inputs(reads)
    engine.upload_reads(buf, offs)
    recs, tot = engine.compare_batch(order, hi, dpass, drev)
    crc = 0
    for a in range(0, len(recs), 1 << 21):  # formatted in pieces: 540 MB of text
        crc = zlib.crc32(host.format_records(recs[a:a + (1 << 21)], order.astype(np.int64), lens_sorted, dpass).encode(), crc)
    assert crc == bench.KNOWN_TEXT_CRC[5] and len(recs) == 24887125
