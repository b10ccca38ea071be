"""PD-HMM (LoglessPDPairHMM, DRAGEN-GATK mode).  The reference's only fixture for it is a git-lfs
stub, so the oracle is anchored by: all-zero PD bytes == the pinned LoglessPairHMM oracle bit for bit, an independent
pure-Python restatement with full matrices, and hand-checkable SNP / deletion cases.  GPU part: the CUDA path against
the oracle through gphmm_pd_compute."""
import numpy as np
import pytest

from oracle import oracle, pyoracle

L = np.frombuffer(b"ACGT", dtype=np.uint8)
SNP, DEL_START, DEL_END, A, C, G, T = 1, 2, 4, 8, 16, 32, 64


def random_pd(rng, H, mode):
    """mode 0: no flags; 1: SNP sites; 2: SNPs + well-formed deletions; 3: arbitrary flag bytes"""
    pd = np.zeros(H, np.uint8)
    if mode == 3:
        return rng.integers(0, 128, H).astype(np.uint8)
    if mode >= 1:
        sites = rng.random(H) < 0.12
        pd[sites] = SNP | rng.choice([A, C, G, T], int(sites.sum())) | rng.choice([0, A, C, G, T], int(sites.sum()))
    if mode >= 2:
        j = 0
        while j < H:
            if rng.random() < 0.08:
                e = min(H - 1, j + int(rng.integers(1, 7)) - 1)
                pd[j] |= DEL_START
                pd[e] |= DEL_END
                j = e + 1
            else:
                j += 1
    return pd


def random_pair(rng, max_h=40, max_r=30):
    H, R = int(rng.integers(1, max_h + 1)), int(rng.integers(1, max_r + 1))
    hap = L[rng.integers(0, 4, H)]
    read = L[rng.integers(0, 4, R)]
    if rng.random() < 0.7:
        n = min(R, H)
        off = int(rng.integers(0, H - n + 1))
        read[:n] = hap[off:off + n]
    quals = [rng.integers(6, 41, R).astype(np.uint8), rng.integers(10, 60, R).astype(np.uint8),
             rng.integers(10, 60, R).astype(np.uint8), rng.integers(5, 30, R).astype(np.uint8)]
    return hap, read, quals


def test_zero_pd_bytes_equal_the_plain_pairhmm_bit_for_bit():
    rng = np.random.default_rng(3)
    for _ in range(200):
        hap, read, q = random_pair(rng, 120, 90)
        assert oracle.pd_logless(hap, np.zeros(len(hap), np.uint8), read, *q) == oracle.logless(hap, read, *q)


def test_c_oracle_equals_python_restatement():
    rng = np.random.default_rng(4)
    for t in range(400):
        hap, read, q = random_pair(rng)
        pd = random_pd(rng, len(hap), t % 4)
        a = oracle.pd_logless(hap, pd, read, *q)
        b = pyoracle.pd_logless(bytes(hap), bytes(pd), bytes(read), *q)
        assert abs(a - b) < 1e-9, (t, a, b)


def test_snp_and_deletion_semantics():
    hap = np.frombuffer(b"ACGTACGTAC", dtype=np.uint8)
    q, i, g = np.full(8, 30, np.uint8), np.full(8, 45, np.uint8), np.full(8, 10, np.uint8)
    # a SNP column T|A: a read carrying the A scores (almost exactly) like against the haplotype with the A
    pd = np.zeros(10, np.uint8)
    pd[3] = SNP | A
    read = np.frombuffer(b"ACGAACGT", dtype=np.uint8)
    with_alt = oracle.logless(np.frombuffer(b"ACGAACGTAC", dtype=np.uint8), read, q, i, i, g)
    assert abs(oracle.pd_logless(hap, pd, read, q, i, i, g) - with_alt) < 1e-6
    assert oracle.logless(hap, read, q, i, i, g) < with_alt - 3           # the undetermined haplotype alone is a mismatch
    # a deletion of columns 5-6: a read without those bases scores like against the deleted haplotype (whose initial
    # condition is 1/8 instead of 1/10)
    pd = np.zeros(10, np.uint8)
    pd[4], pd[5] = DEL_START, DEL_END
    read = np.frombuffer(b"ACGTGTAC", dtype=np.uint8)
    deleted = oracle.logless(np.frombuffer(b"ACGTGTAC", dtype=np.uint8), read, q, i, i, g)
    assert abs(oracle.pd_logless(hap, pd, read, q, i, i, g) - (deleted + np.log10(8 / 10))) < 1e-4
    # and a read WITH those bases still scores like against the full haplotype
    full = oracle.logless(hap, hap[:8], q, i, i, g)
    assert abs(oracle.pd_logless(hap, pd, hap[:8], q, i, i, g) - full) < 1e-3
    # a read base that is not ACGT at a SNP column is an error in the reference (LoglessPDPairHMM.java:202)
    pd = np.zeros(10, np.uint8)
    pd[3] = SNP | A
    with pytest.raises(ValueError):
        oracle.pd_logless(hap, pd, np.frombuffer(b"ACGRACGT", dtype=np.uint8), q, i, i, g)
