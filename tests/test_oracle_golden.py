"""Pins the CPU oracle (oracle/) to the reference's own fixtures and known-answer tests.  CPU only."""
import math

import numpy as np
import pytest

from oracle import oracle, pyoracle
from phmm_testutil import const_quals, load_hmmresults, load_testdata


def test_java_hmmresults_exact_text():
    # HaplotypeCallerIntegrationTest.java:2197,2241 compares the LOGLESS_CACHING dump as exact text;
    # PairHMM.java:380 formats with %e.  The oracle must reproduce every token.
    recs = load_hmmresults()
    assert len(recs) == 284
    bad = 0
    for r in recs:
        v = oracle.logless(r["hap"], r["read"], r["base_q"], r["ins_q"], r["del_q"], r["gcp"])
        assert abs(v - r["java"]) <= 1e-6
        bad += ("%e" % v) != r["java_text"]
    assert bad == 0


def test_exact_and_original_hmmresults():
    recs = load_hmmresults()
    for r in recs:
        args = (r["hap"], r["read"], r["base_q"], r["ins_q"], r["del_q"], r["gcp"])
        assert abs(oracle.log10hmm(*args, exact=True) - r["exact"]) <= 1e-6
        assert abs(oracle.log10hmm(*args, exact=False) - r["original"]) <= 1e-6


def test_avx_hmmresults_within_float_noise():
    # The AVX file is only shape-checked by the reference (HaplotypeCallerIntegrationTest.java:2226-2237);
    # it still has to sit within float noise of the double oracle.
    for r in load_hmmresults():
        v = oracle.logless(r["hap"], r["read"], r["base_q"], r["ins_q"], r["del_q"], r["gcp"])
        assert abs(v - r["avx"]) <= 1e-5


def test_pairhmm_testdata_1e5():
    # VectorPairHMMUnitTest.java:100 tolerance
    recs = load_testdata()
    assert len(recs) == 104
    for r in recs:
        v = oracle.logless(r["hap"], r["read"], r["base_q"], r["ins_q"], r["del_q"], r["gcp"])
        assert abs(v - r["expected"]) <= 1e-5


def test_c_oracle_matches_pure_python():
    for r in load_testdata()[:20] + load_hmmresults()[:40]:
        args = (r["hap"], r["read"], r["base_q"], r["ins_q"], r["del_q"], r["gcp"])
        assert abs(oracle.logless(*args) - pyoracle.logless(*args)) <= 1e-9


def test_match_to_match_prob_table():
    # PairHMMModelUnitTest.java:237-245,249-286 : same qual grid, same 1e-9 tolerance
    grid = [0, 1, 2, 5, 10, 13, 17, 20, 23, 27, 30, 43, 57, 70, 100, 200, 254]
    for a, i in enumerate(grid):
        for d in grid[a:]:
            want = min(1.0, max(0.0, 1.0 - (10 ** (-0.1 * i) + 10 ** (-0.1 * d))))
            assert abs(oracle.match_to_match_prob(i, d) - want) <= 1e-9
            assert oracle.match_to_match_prob(i, d) == oracle.match_to_match_prob(d, i)


def test_match_to_match_prob_jacobian_cutoff():
    # MathUtils.java:406-410,479 : a term more than 8.0 log10 units (80 phred) below the other is dropped,
    # so the table is NOT 1 - eps_i - eps_d there (off by up to 1e-8).  The oracle keeps the Java value.
    assert oracle.match_to_match_prob(1, 81) == pytest.approx(1.0 - 10 ** -0.1, abs=1e-15)
    assert abs(oracle.match_to_match_prob(1, 80) - (1.0 - 10 ** -0.1 - 1e-8)) <= 1e-12


def test_qual_to_trans_probs():
    # PairHMMModelUnitTest.java:21-108
    for ins in (1, 10, 45, 93):
        for dele in (1, 30, 45):
            for gcp in (5, 10, 40):
                t = oracle.qual_to_trans_probs(ins, dele, gcp)
                ei, ed, eg = (10 ** (-q / 10.0) for q in (ins, dele, gcp))
                for got, want in zip(t, (max(0.0, 1 - ei - ed), 1 - eg, ei, eg, ed, eg)):
                    assert abs(got - want) <= 1e-9
    with pytest.raises(ValueError):
        oracle.qual_to_trans_probs(200, 10, 10)  # negative as a Java byte: PairHMMModel.java:109


def _ctx(bases, left, right):
    CONTEXT, LEFT, RIGHT = b"ACGTAATGACGATTGCA", b"GATTTATCATCGAGTCTGC", b"CATGGATCGTTATCAGCTATCTCGAGGGATTCACTTAACAGTTTTA"
    return (LEFT if left else b"") + CONTEXT + bases + CONTEXT + (RIGHT if right else b"")


def test_basic_likelihoods_vs_exact_and_theory():
    # PairHMMUnitTest.java:148-199,239-256 (tristate correction off, anchored indel quals)
    MASSIVE = 100
    CTXLEN = 17
    n = 0
    for base_q in (10, 30, 50):
        for indel_q in (20, 40):
            for gcp in (8, 10, 20):
                cases = [(b"A", b"A", 0), (b"A", b"C", base_q)]
                for size in (2, 3, 5, 10, 35):
                    exp = indel_q + (size - 2) * gcp
                    cases += [(b"G", b"G" * size, exp), (b"G" * size, b"G", exp)]
                for ref, read, expected_q in cases:
                    for left, right in ((False, False), (True, True)):
                        hap = _ctx(ref, left, right)
                        rd = _ctx(read, False, False)
                        L = len(rd)
                        bq = const_quals(L, MASSIVE); bq[CTXLEN:CTXLEN + len(read)] = base_q
                        iq = const_quals(L, MASSIVE); iq[CTXLEN] = indel_q
                        dq = const_quals(L, MASSIVE); dq[CTXLEN] = indel_q
                        gq = const_quals(L, MASSIVE); gq[CTXLEN:CTXLEN + len(read)] = gcp
                        ll = oracle.logless(hap, rd, bq, iq, dq, gq, tristate_off=True)
                        ex = oracle.log10hmm(hap, rd, bq, iq, dq, gq, exact=True, tristate_off=True)
                        theory = expected_q / -10.0 + 0.03 + math.log10(1.0 / len(hap))
                        assert abs(ll - theory) <= 0.2
                        assert abs(ll - ex) <= 0.2
                        assert ll <= 0.0
                        n += 1
    assert n > 100


def test_mismatch_in_every_position():
    # PairHMMUnitTest.java:272-326
    hap = b"TTCTCTTCTGTTGTGGCTGGTT"
    for offset_end in (2, 0):
        L = len(hap) - 2 - offset_end
        for k in range(L):
            quals = const_quals(L, 90); quals[k] = 20
            rd = bytearray(hap[2:2 + L]); rd[k] = ord("T") if rd[k] == ord("C") else ord("C")
            gop = const_quals(L, 80)
            v = oracle.logless(hap, bytes(rd), quals, gop, gop, gop, tristate_off=True)
            want = math.log10(1.0 / len(hap) * (1 - 1e-9) ** (L - 1) * 1e-2)
            assert abs(v - want) <= 1e-2


def test_all_matching_read():
    # PairHMMUnitTest.java:390-418
    for read_size in (1, 2, 5, 10):
        for ref_size in (1, 2, 5, 10):
            if ref_size <= read_size:
                continue
            rd, hap = b"A" * read_size, b"A" * ref_size
            v = oracle.logless(hap, rd, const_quals(read_size, 20), const_quals(read_size, 100),
                               const_quals(read_size, 100), const_quals(read_size, 100), tristate_off=True)
            want = math.log10((abs(ref_size - read_size + 1) / ref_size) * 0.99 ** read_size)
            assert abs(v - want) <= 1e-3


def test_really_big_reads_do_not_underflow_in_double():
    # PairHMMUnitTest.java:420-457
    read1, ref1 = b"ACCAAGTAGTCACCGT", b"ACCAAGTAGTCACCGTAACG"
    rd, hap = read1 * 50, ref1 * 100
    n = len(rd)
    v = oracle.logless(hap, rd, const_quals(n, 30), const_quals(n, 40), const_quals(n, 40), const_quals(n, 10))
    assert np.isfinite(v) and v <= 0.0


def test_zero_length_read_and_read_longer_than_hap():
    e = np.zeros(0, dtype=np.uint8)
    assert oracle.logless(b"ACGT", b"", e, e, e, e) == -math.inf  # LoglessPairHMM.java:47 loop never runs
    n = 30
    v = oracle.logless(b"ACGTACGT", b"ACGTACGTAC" * 3, const_quals(n, 30), const_quals(n, 40), const_quals(n, 40), const_quals(n, 10))
    assert np.isfinite(v) and v < 0.0  # PairHMMUnitTest.java:24 allows reads longer than the haplotype


def test_unit_layout_read_major():
    # PairHMM.java:236 / VectorLoglessPairHMM.java:148 : out[r*nHaps + h]
    recs = load_hmmresults()[:6]
    reads = [recs[0], recs[3]]
    haps = [recs[0]["hap"], recs[1]["hap"], recs[5]["hap"]]
    cat = lambda k: np.concatenate([np.frombuffer(r[k], dtype=np.uint8) if isinstance(r[k], bytes) else r[k] for r in reads])
    ro = np.cumsum([0] + [len(r["read"]) for r in reads])
    ho = np.cumsum([0] + [len(h) for h in haps])
    hb = np.frombuffer(b"".join(haps), dtype=np.uint8)
    for threads in (1, 2):
        out = oracle.unit(cat("read"), cat("base_q"), cat("ins_q"), cat("del_q"), cat("gcp"), ro, hb, ho, threads=threads)
        for ri, r in enumerate(reads):
            for hi, h in enumerate(haps):
                assert out[ri * 3 + hi] == oracle.logless(h, r["read"], r["base_q"], r["ins_q"], r["del_q"], r["gcp"])


def test_simd_baseline_matches_oracle():
    # the vectorised fp32 CPU baseline that bench.py reports (not the oracle) agrees with the oracle to float noise,
    # including reads of different lengths in one vector and pairs that need the double redo
    from gatk_b200 import synth
    from phmm_testutil import oracle_batch
    for b in (synth.config2(6), synth.random_batch(11, n_units=4, wild_quals=True), synth.config5(hap_len=400, n_regions=2, reads_per_region=20, n_haps=3, bad_fraction=0.5)):
        u = np.stack([b.units[k] for k in ("read_begin", "read_end", "hap_begin", "hap_end", "out_off")], axis=1)
        got, _ = oracle.simd_batch(b.read_bases, b.base_q, b.ins_q, b.del_q, b.gcp, b.read_off, b.hap_bases, b.hap_off, u, b.n_out, threads=2)
        want = oracle_batch(b)
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), fin)
        assert np.abs(got[fin] - want[fin]).max() <= 2e-5
    assert oracle.simd_isa() in (0, 256, 512)
