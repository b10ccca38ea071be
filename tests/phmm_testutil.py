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
