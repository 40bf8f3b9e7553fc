"""Shared helpers for the parity tests (oracle = checker, engine = product)."""
import numpy as np

from amplicon_sorter_b200 import host, synth, thresholds
from oracle import oracle


def batch_inputs(reads, similar_genes=80.0):
    buf, offs = synth.pack_reads(reads)
    lens = (offs[1:] - offs[:-1]).astype(np.int64)
    order = oracle.stable_length_order(lens)
    lens_sorted = lens[order]
    hi = host.batch_geometry(lens_sorted)
    dpass, drev = thresholds.tables(similar_genes / 100, int(lens.max()) + 1 if len(reads) else 1)
    return buf, offs, order, lens_sorted, hi, dpass, drev


def gpu_batch(engine, reads, similar_genes=80.0, world=1, **params):
    buf, offs, order, lens_sorted, hi, dpass, drev = batch_inputs(reads, similar_genes)
    engine.upload_reads(buf, offs)
    for k, v in params.items():
        engine.set_param(k, v)
    parts, tot = [], None
    for rank in range(world):
        recs, t = engine.compare_batch(order, hi, dpass, drev, rank, world)
        parts.append(recs)
        tot = t if tot is None else {k: tot[k] + t[k] for k in tot}
    recs = np.concatenate(parts)
    if world > 1:
        key = recs["i_pos"].astype(np.uint64) << np.uint64(32) | recs["j_pos"].astype(np.uint64)
        recs = recs[np.argsort(key, kind="stable")]
    return recs, tot


def oracle_batch(reads, similar_genes=80.0, algo="myers"):
    buf, offs, order, *_ = batch_inputs(reads, similar_genes)
    return oracle.process_batch(buf, offs, order, similar_genes, algo=algo)


def assert_same_records(a, b):
    assert a.shape == b.shape, (a.shape, b.shape)
    for f in ("i_pos", "j_pos", "d", "reverse"):
        bad = np.nonzero(a[f] != b[f])[0]
        assert bad.size == 0, (f, bad[:5], a[bad[:5]], b[bad[:5]])


def random_reads(rng, n, lo, hi, alphabet=b"ACGT", families=0, err=0.06):
    """n reads; with families > 0, reads are noisy copies (either strand) of `families` templates."""
    al = np.frombuffer(alphabet, dtype=np.uint8)
    reads = []
    if families:
        T = [rng.integers(0, 4, size=int(rng.integers(lo, hi + 1)), dtype=np.uint8) for _ in range(families)]
        for _ in range(n):
            t = T[int(rng.integers(0, families))]
            r = synth.mutate(rng, t, sub=err / 2, ins=err / 4, dele=err / 4)
            if rng.random() < 0.5:
                r = synth.revcomp_codes(r)
            reads.append(np.frombuffer(b"ACGT", dtype=np.uint8)[r].tobytes())
    else:
        for _ in range(n):
            reads.append(al[rng.integers(0, al.size, size=int(rng.integers(lo, hi + 1)))].tobytes())
    return reads
