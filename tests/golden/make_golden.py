#!/usr/bin/env python
"""Convert the reference's own PairHMM fixtures into the golden vectors committed next to this script.

Run once in the build container (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

Sources (read-only, all under /root/reference/src/test/resources/):
  pairhmm-testdata.txt
      104 (hap, read, quals, expected log10 lk) records; pinned at 1e-5 for the native (float) path by
      src/test/java/org/broadinstitute/hellbender/utils/pairhmm/VectorPairHMMUnitTest.java:24,100.
      Quals are FASTQ+33; the test floors BASE quals at 6 before use (:65,112-118) -- the floor is
      NOT applied here, loaders apply it (tests/phmm_testutil.py::load_testdata).
  org/broadinstitute/hellbender/tools/haplotypecaller/expected.{Java,AVX,Exact,Original}.hmmresults.txt
      284 records each, identical inputs, one result column per implementation
      (HaplotypeCallerIntegrationTest.java:2193-2244; format written by PairHMM.java:364-385).
      Quals are FASTQ+33 and already final (no floor).

Outputs (gzip'd TSV, one header line):
  pairhmm_testdata.tsv.gz   hap read baseQ insQ delQ gcp expected
  hmmresults.tsv.gz         hap read baseQ insQ delQ gcp java avx exact original
Values are kept as the exact text tokens of the source files.
"""
import gzip
import os

REF = "/root/reference/src/test/resources/"
HERE = os.path.dirname(os.path.abspath(__file__))


def records(path):
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            tok = line.split()
            assert len(tok) == 7, (path, len(tok))
            yield tok


def write_gz(name, header, rows):
    # mtime=0 keeps the file byte-stable across regenerations
    with open(os.path.join(HERE, name), "wb") as raw:
        with gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as gz:
            gz.write(("\t".join(header) + "\n").encode())
            for r in rows:
                gz.write(("\t".join(r) + "\n").encode())


def main():
    rows = list(records(REF + "pairhmm-testdata.txt"))
    assert len(rows) == 104, len(rows)
    write_gz("pairhmm_testdata.tsv.gz", ["hap", "read", "baseQ", "insQ", "delQ", "gcp", "expected"], rows)

    d = REF + "org/broadinstitute/hellbender/tools/haplotypecaller/"
    per_impl = {k: list(records(d + "expected.%s.hmmresults.txt" % k)) for k in ("Java", "AVX", "Exact", "Original")}
    n = len(per_impl["Java"])
    assert n == 284, n
    merged = []
    for i in range(n):
        base = per_impl["Java"][i][:6]
        for k in ("AVX", "Exact", "Original"):
            assert per_impl[k][i][:6] == base, (k, i)
        merged.append(base + [per_impl[k][i][6] for k in ("Java", "AVX", "Exact", "Original")])
    write_gz("hmmresults.tsv.gz", ["hap", "read", "baseQ", "insQ", "delQ", "gcp", "java", "avx", "exact", "original"], merged)
    print("wrote", len(rows), "+", n, "records")


if __name__ == "__main__":
    main()
