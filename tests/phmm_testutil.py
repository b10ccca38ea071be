"""Shared helpers for the PairHMM tests: golden-fixture loaders and small input builders."""
import gzip
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fastq(s, floor=0):
    q = np.frombuffer(s.encode(), dtype=np.uint8).astype(np.int16) - 33
    return np.maximum(q, floor).astype(np.uint8)


def _rows(name):
    with gzip.open(os.path.join(GOLDEN, name), "rt") as f:
        header = f.readline().rstrip("\n").split("\t")
        for line in f:
            yield dict(zip(header, line.rstrip("\n").split("\t")))


def load_testdata():
    """pairhmm-testdata.txt records with the test's own normalisation: q-33, base quals floored at 6
    (VectorPairHMMUnitTest.java:65-68,112-118)."""
    out = []
    for r in _rows("pairhmm_testdata.tsv.gz"):
        out.append(dict(hap=r["hap"].encode(), read=r["read"].encode(), base_q=_fastq(r["baseQ"], 6),
                        ins_q=_fastq(r["insQ"]), del_q=_fastq(r["delQ"]), gcp=_fastq(r["gcp"]),
                        expected=float(r["expected"])))
    return out


def load_hmmresults():
    """expected.{Java,AVX,Exact,Original}.hmmresults.txt merged; quals are final (q-33, no floor)."""
    out = []
    for r in _rows("hmmresults.tsv.gz"):
        out.append(dict(hap=r["hap"].encode(), read=r["read"].encode(), base_q=_fastq(r["baseQ"]),
                        ins_q=_fastq(r["insQ"]), del_q=_fastq(r["delQ"]), gcp=_fastq(r["gcp"]),
                        java=float(r["java"]), avx=float(r["avx"]), exact=float(r["exact"]),
                        original=float(r["original"]), java_text=r["java"]))
    return out


def const_quals(n, q):
    return np.full(n, q, dtype=np.uint8)


def oracle_batch(batch, tristate_off=False, threads=0):
    """Double-precision oracle for every unit of a gatk_b200 Batch, in the batch's output layout."""
    from oracle import oracle
    out = np.full(batch.n_out, np.nan, dtype=np.float64)
    threads = threads or oracle.max_threads()
    for u in batch.units:
        r0, r1, h0, h1, o = (int(u[k]) for k in ("read_begin", "read_end", "hap_begin", "hap_end", "out_off"))
        if r1 == r0 or h1 == h0:
            continue
        b0, b1 = int(batch.read_off[r0]), int(batch.read_off[r1])
        c0, c1 = int(batch.hap_off[h0]), int(batch.hap_off[h1])
        res = oracle.unit(batch.read_bases[b0:b1], batch.base_q[b0:b1], batch.ins_q[b0:b1], batch.del_q[b0:b1],
                          batch.gcp[b0:b1], batch.read_off[r0:r1 + 1] - b0, batch.hap_bases[c0:c1],
                          batch.hap_off[h0:h1 + 1] - c0, tristate_off=tristate_off, threads=threads)
        out[o:o + len(res)] = res
    return out


def records_to_batch(recs):
    """Each fixture record becomes its own 1-read x 1-haplotype unit (as VectorPairHMMUnitTest.java:95-98 does)."""
    from gatk_b200.native import UNIT_DTYPE, Batch
    as_u8 = lambda x: np.frombuffer(x, dtype=np.uint8) if isinstance(x, bytes) else x
    rb = np.concatenate([as_u8(r["read"]) for r in recs])
    cols = [np.concatenate([r[k] for r in recs]) for k in ("base_q", "ins_q", "del_q", "gcp")]
    hb = np.concatenate([as_u8(r["hap"]) for r in recs])
    read_off = np.concatenate([[0], np.cumsum([len(r["read"]) for r in recs])])
    hap_off = np.concatenate([[0], np.cumsum([len(r["hap"]) for r in recs])])
    units = np.array([(i, i + 1, i, i + 1, i) for i in range(len(recs))], dtype=UNIT_DTYPE)
    return Batch(rb, cols[0], cols[1], cols[2], cols[3], read_off, hb, hap_off, units)
