"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the reference's golden vectors.
All tests here need a B200 (`-m gpu`)."""
import math

import numpy as np
import pytest

from gatk_b200 import native, synth
from gatk_b200.native import Batch, GpuPhmm
from phmm_testutil import const_quals, load_hmmresults, load_testdata, oracle_batch, records_to_batch

pytestmark = pytest.mark.gpu

TOL = 1e-4  # BASELINE.json north_star: |log10 lk - LoglessPairHMM(double)| <= 1e-4


@pytest.fixture(scope="module")
def hmm():
    h = GpuPhmm()
    yield h
    h.close()


@pytest.fixture(scope="module")
def hmm64():
    h = GpuPhmm(force_fp64=True)
    yield h
    h.close()


def _check(got, want, tol):
    assert got.shape == want.shape
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin)
    assert np.array_equal(got[~fin], want[~fin])
    err = np.abs(got[fin] - want[fin])
    assert err.size == 0 or err.max() <= tol, "max err %.3g at %d" % (err.max(), int(np.argmax(err)))
    assert np.all(got[fin] <= 1e-9)  # PairHMM.java:305: a log10 probability is never > 0


def test_golden_pairhmm_testdata(hmm):
    # VectorPairHMMUnitTest.java:100 : 1e-5 against the file's expected column
    recs = load_testdata()
    got = hmm.compute(records_to_batch(recs))
    want = np.array([r["expected"] for r in recs])
    assert np.abs(got - want).max() <= 1e-5


def test_golden_hmmresults(hmm, hmm64):
    recs = load_hmmresults()
    b = records_to_batch(recs)
    java = np.array([r["java"] for r in recs])
    got = hmm.compute(b)
    assert np.abs(got - java).max() <= 1e-5
    got64 = hmm64.compute(b)
    assert np.abs(got64 - java).max() <= 1e-6  # print precision of the fixture
    assert all(("%e" % v) == r["java_text"] for v, r in zip(got64, recs))  # fp64 mode reproduces the text dump


def test_golden_hmmresults_as_one_region(hmm):
    # the same 284 pairs come from one region: 20 reads x ... ; group by haplotype set instead of 1x1 units
    recs = load_hmmresults()
    haps = sorted({r["hap"] for r in recs})
    reads = {}
    for r in recs:
        reads.setdefault((r["read"], r["base_q"].tobytes(), r["ins_q"].tobytes(), r["del_q"].tobytes(), r["gcp"].tobytes()), r)
    rl = list(reads.values())
    b = Batch.single_unit([(r["read"], r["base_q"], r["ins_q"], r["del_q"], r["gcp"]) for r in rl], haps)
    got = hmm.compute(b)
    _check(got, oracle_batch(b), 1e-5)
    # and every fixture row is reproduced at its (read, hap) slot
    idx = {k: i for i, k in enumerate(reads.keys())}
    for r in recs:
        ri = idx[(r["read"], r["base_q"].tobytes(), r["ins_q"].tobytes(), r["del_q"].tobytes(), r["gcp"].tobytes())]
        assert abs(got[ri * len(haps) + haps.index(r["hap"])] - r["java"]) <= 1e-5


@pytest.mark.parametrize("seed", range(6))
def test_random_ragged_batches(hmm, seed):
    b = synth.random_batch(100 + seed, n_units=5, wild_quals=bool(seed % 2))
    _check(hmm.compute(b), oracle_batch(b), TOL)


@pytest.mark.parametrize("seed", range(3))
def test_random_ragged_batches_fp64(hmm64, seed):
    b = synth.random_batch(200 + seed, n_units=4, wild_quals=True)
    _check(hmm64.compute(b), oracle_batch(b), 1e-9)


def test_every_read_length_bucket(hmm):
    # one read per length 1..300 exercises every rows-per-lane kernel, the bucket edges (31/32, 254/255/256)
    # and the striped kernel for reads of 255+ bases
    rng = np.random.default_rng(5)
    hap = rng.integers(0, 4, 320, dtype=np.uint8)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    reads = []
    for R in list(range(1, 301)):
        off = int(rng.integers(0, 320 - R + 1))
        rd = letters[hap[off:off + R]].copy()
        if R > 3:
            rd[R // 2] = ord("A") if rd[R // 2] != ord("A") else ord("C")
        reads.append((rd, const_quals(R, 30), const_quals(R, 45), const_quals(R, 45), const_quals(R, 10)))
    b = Batch.single_unit(reads, [letters[hap], letters[hap[:200]], letters[hap[7:]]])
    _check(hmm.compute(b), oracle_batch(b), TOL)


def test_config1_matches_oracle(hmm):
    b = synth.config1()
    got = hmm.compute(b)
    _check(got, oracle_batch(b), TOL)
    s = hmm.stats()
    assert s["cells"] >= b.cells()


def test_config2_sample_matches_oracle(hmm):
    b = synth.config2(40)
    _check(hmm.compute(b), oracle_batch(b), TOL)


def test_fp64_rescue_path(hmm):
    # indel-heavy reads under-flow fp32 and must come back from the fp64 redo with double accuracy
    b = synth.config5(hap_len=600, n_regions=4, reads_per_region=24, n_haps=4, bad_fraction=0.5)
    hmm.reset_stats()
    got = hmm.compute(b)
    want = oracle_batch(b)
    s = hmm.stats()
    assert s["rescued_pairs"] > 0
    assert want.min() < -70  # the workload really leaves the fp32 range
    _check(got, want, TOL)
    low = want < -70
    assert np.abs(got[low] - want[low]).max() <= 1e-9  # rescued pairs carry fp64 accuracy


def test_config4_deep_coverage_every_pair(hmm):
    # BASELINE.json configs[3] at the PairHMM boundary (Mutect2 500x: 4 000 reads x 32 haplotypes per region, 150 bp
    # reads): EVERY pair of two full-size regions against the double-precision oracle (VectorPairHMMUnitTest.java:100
    # compares every pair of its fixture the same way), through the synchronous call and through the queue
    b = synth.config4(n_regions=2)
    assert b.pairs() == 2 * 4000 * 32
    want = oracle_batch(b)
    hmm.reset_stats()
    got = hmm.compute(b)
    _check(got, want, TOL)
    assert hmm.stats()["pairs"] == b.pairs()
    t = hmm.submit(b)
    assert np.array_equal(hmm.wait(t), got)
    with GpuPhmm(no_prefix_sharing=True) as plain:
        assert np.array_equal(plain.compute(b), got)


@pytest.mark.parametrize("bad_fraction", [0.0, 0.1, 0.5])
def test_config5_long_haplotypes_h1000(hmm, bad_fraction):
    # BASELINE.json configs[4] at its largest haplotype length (250 bp reads x 1 000 bp haplotypes; reads of
    # PairHMMUnitTest.java:420-457 "really big" shapes) with 0 / 10 / 50 % indel-heavy reads: every pair against the
    # oracle, and the pairs that left the fp32 range come back from the fp64 redo with double accuracy
    b = synth.config5(hap_len=1000, n_regions=6, reads_per_region=32, n_haps=8, bad_fraction=bad_fraction)
    want = oracle_batch(b)
    hmm.reset_stats()
    got = hmm.compute(b)
    s = hmm.stats()
    _check(got, want, TOL)
    # a pair is redone when its raw fp32 sum is below 1e-28; the sum is likelihood * 2^(116 - 10) * H, i.e. likelihoods
    # below ~1e-63 (GKL's AVX float path switches to double on the same kind of threshold)
    low = want < -64
    if bad_fraction == 0.0:
        assert not low.any()
    else:
        assert low.sum() >= 0.5 * bad_fraction * low.size * 0.5 and s["rescued_pairs"] >= low.sum()
        assert np.abs(got[low] - want[low]).max() <= 1e-9
    with GpuPhmm(chunk_cells=200_000_000) as small:   # several chunks: the redo list is per chunk
        assert np.array_equal(small.compute(b), got)


def test_really_big_reads(hmm, hmm64):
    # PairHMMUnitTest.java:420-457 : reads up to 800 bp x haplotypes up to 2000 bp
    read1, ref1 = b"ACCAAGTAGTCACCGT", b"ACCAAGTAGTCACCGTAACG"
    reads, haps = [], []
    for n in (1, 2, 10, 20, 50):
        rd = read1 * n
        reads.append((rd, const_quals(len(rd), 30), const_quals(len(rd), 40), const_quals(len(rd), 40), const_quals(len(rd), 10)))
    for n in (2, 10, 20, 100):
        haps.append(ref1 * n)
    b = Batch.single_unit(reads, haps)
    want = oracle_batch(b)
    _check(hmm.compute(b), want, TOL)
    _check(hmm64.compute(b), want, 1e-9)


def test_basic_likelihoods_tristate_off():
    # PairHMMUnitTest.java:148-199 : substitution / insertion / deletion micro-reads, tristate correction off
    CONTEXT, LEFT = b"ACGTAATGACGATTGCA", b"GATTTATCATCGAGTCTGC"
    reads, haps, expect = [], [], []
    h = GpuPhmm(tristate_off=True)
    try:
        units_reads, units_haps = [], []
        for base_q in (10, 30, 50):
            for indel_q in (20, 40):
                for gcp in (8, 10, 20):
                    for ref, read, eq in ((b"A", b"A", 0), (b"A", b"C", base_q), (b"G", b"GGG", indel_q + gcp), (b"GGGGG", b"G", indel_q + 3 * gcp)):
                        hap = LEFT + CONTEXT + ref + CONTEXT
                        rd = CONTEXT + read + CONTEXT
                        L = len(rd)
                        bq = const_quals(L, 100); bq[17:17 + len(read)] = base_q
                        iq = const_quals(L, 100); iq[17] = indel_q
                        dq = const_quals(L, 100); dq[17] = indel_q
                        gq = const_quals(L, 100); gq[17:17 + len(read)] = gcp
                        units_reads.append((rd, bq, iq, dq, gq))
                        units_haps.append(hap)
                        expect.append(eq / -10.0 + 0.03 + math.log10(1.0 / len(hap)))
        recs = [dict(read=r[0], base_q=r[1], ins_q=r[2], del_q=r[3], gcp=r[4], hap=hp) for r, hp in zip(units_reads, units_haps)]
        b = records_to_batch(recs)
        got = h.compute(b)
        assert np.abs(got - np.array(expect)).max() <= 0.2
        _check(got, oracle_batch(b, tristate_off=True), TOL)
    finally:
        h.close()


def test_n_bases_and_exotic_bytes(hmm):
    # LoglessPairHMM.java:89 : N in the read or the haplotype matches anything; other bytes compare exactly
    hap = b"ACGTNACGTRRACGTacgtACGT"
    reads = []
    for rd in (b"ACGTAACGT", b"NNNNNNNNN", b"ACGTRRACG", b"acgtACGT", b"ACGTYACGT", b"ACGTNACGTRRACGTacgtACGT"):
        n = len(rd)
        reads.append((rd, const_quals(n, 30), const_quals(n, 45), const_quals(n, 45), const_quals(n, 10)))
    b = Batch.single_unit(reads, [hap, b"ACGTACGTACGT", b"NNNNNNNNNNNNNNNN", b"RRRRYYYYKKKKMMMMSSSSWWWW"])
    _check(hmm.compute(b), oracle_batch(b), TOL)


def test_edge_shapes(hmm):
    # reads longer than the haplotype (PairHMMUnitTest.java:24), 1-base reads/haps, empty units, zero-length read
    e = np.zeros(0, np.uint8)
    reads = [(b"A", const_quals(1, 30), const_quals(1, 45), const_quals(1, 45), const_quals(1, 10)),
             (b"ACGTACGTACGTACGTACGTACGTACGTACGTACGT", const_quals(36, 25), const_quals(36, 40), const_quals(36, 40), const_quals(36, 10)),
             (b"", e, e, e, e)]
    b = Batch.single_unit(reads, [b"A", b"ACGT", b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT"])
    got = hmm.compute(b)
    want = oracle_batch(b)
    assert np.all(np.isneginf(want[6:9]))  # zero-length read: log10(0) (LoglessPairHMM.java:47 never runs)
    _check(got, want, TOL)
    # a unit without reads and a unit without haplotypes produce no output and no error
    u = np.array([(0, 0, 0, 3, 0), (0, 3, 0, 0, 0), (0, 2, 1, 3, 0)], dtype=native.UNIT_DTYPE)
    b2 = Batch(b.read_bases, b.base_q, b.ins_q, b.del_q, b.gcp, b.read_off, b.hap_bases, b.hap_off, u)
    got2 = hmm.compute(b2)
    _check(got2, oracle_batch(b2), TOL)
    empty = Batch(e, e, e, e, e, [0], e, [0], np.zeros(0, dtype=native.UNIT_DTYPE))
    assert hmm.compute(empty).size == 0  # VectorLoglessPairHMM.java:110-112 early return


def test_error_conventions(hmm):
    b = synth.config1()
    bad = Batch(b.read_bases, b.base_q, b.ins_q.copy(), b.del_q, b.gcp, b.read_off, b.hap_bases, b.hap_off, b.units)
    bad.ins_q[5] = 200  # negative as a Java byte: PairHMMModel.java:109 IllegalArgumentException
    with pytest.raises(native.GpuPhmmError) as e:
        hmm.compute(bad)
    assert e.value.code == native.ERR_BAD_QUAL
    bad2 = Batch(b.read_bases, b.base_q.copy(), b.ins_q, b.del_q, b.gcp, b.read_off, b.hap_bases, b.hap_off, b.units)
    bad2.base_q[0] = 255  # QualityUtils.java:157 cache index out of bounds
    with pytest.raises(native.GpuPhmmError):
        hmm.compute(bad2)
    # zero-length haplotype: PairHMM.java:139
    hoff = b.hap_off.copy(); hoff[1] = hoff[0]
    with pytest.raises(native.GpuPhmmError) as e:
        hmm.compute(Batch(b.read_bases, b.base_q, b.ins_q, b.del_q, b.gcp, b.read_off, b.hap_bases, hoff, b.units))
    assert e.value.code == native.ERR_INVALID_ARG
    # the handle stays usable after an error
    _check(hmm.compute(b), oracle_batch(b), TOL)
    # same conventions without an fp32 pass (--native-pair-hmm-use-double-precision)
    with GpuPhmm(force_fp64=True) as h64:
        for broken in (bad, bad2):
            with pytest.raises(native.GpuPhmmError) as e:
                h64.compute(broken)
            assert e.value.code == native.ERR_BAD_QUAL


def test_async_queue_and_prepared(hmm):
    batches = [synth.random_batch(300 + k, n_units=3) for k in range(4)]
    tickets = [hmm.submit(b) for b in batches]
    for t, b in zip(tickets, batches):
        _check(hmm.wait(t), oracle_batch(b), TOL)
    with pytest.raises(native.GpuPhmmError):
        hmm.wait(tickets[0])  # already collected
    b = synth.config2(16, pinned=True)
    p = hmm.prepare(b)
    out = np.full(b.n_out, np.nan)
    hmm.run_prepared(p, out)
    out2 = np.full(b.n_out, np.nan)
    hmm.run_prepared(p, out2)
    hmm.release_prepared(p)
    assert np.array_equal(out, out2)  # idempotent, deterministic
    _check(out, oracle_batch(b), TOL)
    assert np.array_equal(out, hmm.compute(b))  # staged path == prepared path, bit for bit


def test_chunking_is_invisible():
    b = synth.config2(24)
    with GpuPhmm() as big, GpuPhmm(chunk_cells=30_000_000, chunk_bytes=200_000) as small:
        a, c = big.compute(b), small.compute(b)
    assert np.array_equal(a, c)


def test_lane_layout_is_invisible():
    # a chunk with enough reads runs two reads of a unit per warp (half-warp kernels, 10-16 rows per lane) where their last
    # rows fall on the same register slot; small chunks keep one read per warp (5-8 rows per lane) and split the haplotypes
    # into groups.  A row's arithmetic does not depend on the slot it lands on, so both layouts give the same bits --
    # for 250-base reads (configs[1]), 150-base reads (configs[0] shape) and PCR-indel-model-like per-base gap qualities
    rng = np.random.default_rng(9)
    for b in (synth.config2(200), synth.config1_many(48)):
        assert b.n_reads >= 148 * 16 * 2
        sym = b.ins_q.copy()
        sym[rng.random(len(sym)) < 0.1] = 38
        variants = (b, Batch(b.read_bases, b.base_q, sym, sym.copy(), b.gcp, b.read_off, b.hap_bases, b.hap_off, b.units))
        with GpuPhmm() as paired, GpuPhmm(chunk_cells=400_000_000) as single:
            for v in variants:
                p = paired.prepare(v)   # one chunk holding every read (the staged path ramps its chunk sizes up from small ones)
                a = np.full(v.n_out, np.nan)
                paired.reset_stats()
                paired.run_prepared(p, a)
                paired.release_prepared(p)
                assert paired.stats()["rescued_pairs"] == 0   # (a task no kernel ran would come back from the fp64 redo)
                c = single.compute(v)
                assert np.array_equal(a, c)
        sub = Batch(b.read_bases, b.base_q, sym, sym.copy(), b.gcp, b.read_off, b.hap_bases, b.hap_off, b.units[:6])
        want = oracle_batch(sub)
        n = sub.n_out
        assert np.abs(a[:n] - want[:n]).max() <= TOL


def test_half_warp_general_kernel():
    # per-base gap-open AND gap-continuation qualities (DRAGstr: DragstrPairHMMInputScoreImputator.java:57-66; BI/BD tags): in a
    # chunk large enough to pair reads, reads of 128-191 bases run the half-warp form of the general kernel (the flat-kernel
    # family with four coefficient registers per row), longer ones the full-warp kernel.  Against the oracle on a sample of
    # units, and against the one-read-per-warp layout (small chunks) on every pair.
    rng = np.random.default_rng(17)
    for b0, n_units in ((synth.config1_many(48), 3), (synth.config2(200), 4)):
        n = len(b0.read_bases)
        r = rng.random(n)
        gop = np.full(n, 40, np.uint8)
        gop[r < 0.25] = rng.integers(30, 40, int((r < 0.25).sum()))
        gop[r < 0.05] = rng.integers(15, 30, int((r < 0.05).sum()))
        gcp = np.full(n, 10, np.uint8)
        gcp[r < 0.25] = rng.integers(6, 12, int((r < 0.25).sum()))
        for ins, dele in ((gop, gop.copy()), (gop, np.minimum(gop + rng.integers(0, 6, n), 60).astype(np.uint8))):
            b = Batch(b0.read_bases, b0.base_q, ins, dele, gcp, b0.read_off, b0.hap_bases, b0.hap_off, b0.units)
            assert b.n_reads >= 148 * 16 * 2
            with GpuPhmm() as paired, GpuPhmm(chunk_cells=400_000_000) as single:
                p = paired.prepare(b)   # one chunk holding every read
                a = np.full(b.n_out, np.nan)
                paired.run_prepared(p, a)
                paired.release_prepared(p)
                c = single.compute(b)
            assert np.all(np.isfinite(a)) and np.abs(a - c).max() <= 1e-5
            sub = Batch(b.read_bases, b.base_q, ins, dele, gcp, b.read_off, b.hap_bases, b.hap_off, b.units[:n_units])
            want = oracle_batch(sub)
            assert np.abs(a[:sub.n_out] - want).max() <= TOL


def test_quarter_warp_tasks():
    # three or four reads of EQUAL length (up to 159 bases) of a unit share a warp, 8 lanes each (quarter-warp buckets); the
    # other reads of the unit keep their half-warp / full-warp tasks.  Every quality class, mixed classes inside one task, a
    # remainder of three, a haplotype that is a prefix of another (snapshot next to an END column, 8-step windows).  The fp64
    # redo would hide a task that no kernel ran (its sums stay zero), so the redo count is checked too.
    rng = np.random.default_rng(5)
    L4 = np.frombuffer(b"ACGT", dtype=np.uint8)
    hap = L4[rng.integers(0, 4, 300)]
    haps = [hap.tobytes(), hap[:250].tobytes(), hap[10:].tobytes()]
    with GpuPhmm() as hmm:
        for N, L, mixed in ((4, 150, False), (3, 151, False), (7, 150, True), (4, 119, False), (3, 100, True), (5, 76, False), (4, 64, True), (4, 159, True)):
            reads = []
            for k in range(N):
                off = int(rng.integers(0, 300 - L + 1))
                rd = hap[off:off + L].copy()
                rd[L // 3] = ord("T") if rd[L // 3] != ord("T") else ord("A")
                q = np.clip(rng.normal(30, 8, L), 6, 41).astype(np.uint8)
                kind = k % 4 if mixed else 0
                if kind == 3:    # per-base qualities: the general kernel takes the task's reads two by two
                    iq, dq, gq = (rng.integers(20, 50, L).astype(np.uint8) for _ in range(3))
                elif kind == 2:  # ins == del per base, flat gcp
                    iq = rng.integers(30, 45, L).astype(np.uint8)
                    dq, gq = iq.copy(), const_quals(L, 10)
                else:
                    t = [(45, 45, 10), (40, 42, 10)][kind]
                    iq, dq, gq = const_quals(L, t[0]), const_quals(L, t[1]), const_quals(L, t[2])
                reads.append((rd, q, iq, dq, gq))
            for k in range(5):   # reads of other lengths in the same unit
                R = 200 + k
                off = int(rng.integers(0, 300 - R + 1))
                reads.append((hap[off:off + R].copy(), np.full(R, 30, np.uint8), const_quals(R, 45), const_quals(R, 45), const_quals(R, 10)))
            one = Batch.single_unit(reads, haps)
            want = oracle_batch(one)
            copies = 6000 // len(reads) + 1
            big, n = _replicate(one, copies)
            p = hmm.prepare(big)
            out = np.full(big.n_out, np.nan)
            hmm.reset_stats()
            hmm.run_prepared(p, out)
            hmm.release_prepared(p)
            assert hmm.stats()["rescued_pairs"] == 0
            assert np.abs(out.reshape(copies, n) - want).max() <= TOL


def test_invariants_at_scale(hmm):
    # size-independent properties on a larger slice of configs[1]:
    #  - permutation invariance: shuffling the unit order permutes the outputs, bit for bit
    #  - a read's likelihood against a haplotype does not depend on the other haplotypes streamed with it
    b = synth.config2(300)
    got = hmm.compute(b)
    assert np.all(np.isfinite(got)) and np.all(got <= 1e-9)
    perm = np.random.default_rng(3).permutation(len(b.units))
    b2 = Batch(b.read_bases, b.base_q, b.ins_q, b.del_q, b.gcp, b.read_off, b.hap_bases, b.hap_off, b.units[perm])
    assert np.array_equal(hmm.compute(b2), got)
    u = b.units[7]
    nh = int(u["hap_end"] - u["hap_begin"]); nr = int(u["read_end"] - u["read_begin"])
    solo = np.array([(u["read_begin"], u["read_end"], u["hap_begin"] + k, u["hap_begin"] + k + 1, k * nr) for k in range(nh)], dtype=native.UNIT_DTYPE)
    b3 = Batch(b.read_bases, b.base_q, b.ins_q, b.del_q, b.gcp, b.read_off, b.hap_bases, b.hap_off, solo)
    g3 = hmm.compute(b3).reshape(nh, nr)
    ref = got[int(u["out_off"]):int(u["out_off"]) + nr * nh].reshape(nr, nh)
    assert np.abs(g3.T - ref).max() <= 2e-6  # different c0 exponent per unit => not bit-equal, but float-noise close


def test_plugin_surface_like_vector_pairhmm_unit_test():
    # VectorPairHMMUnitTest.java:31-106 through the mirrored plugin surface: one haplotype x one read per call,
    # initialize() + computeLog10Likelihoods() + getLogLikelihoodArray(), tolerance 1e-5 (:100)
    from gatk_b200 import pairhmm
    args = pairhmm.PairHMMNativeArguments(maxNumberOfThreads=1, useDoublePrecision=False)
    hmm = pairhmm.Implementation.CUDA_LOGLESS_CACHING.makeNewHMM(args)
    try:
        for r in load_testdata():
            hap = r["hap"]
            read = pairhmm.Read(r["read"], r["base_q"], r["ins_q"], r["del_q"])

            class Imputator:
                def impute(self, read_):
                    return r["ins_q"], r["del_q"], r["gcp"]

            hmm.initialize([hap], None, 0, 0)
            m = pairhmm.LikelihoodMatrix([hap], 1)
            hmm.computeLog10Likelihoods(m, [read], Imputator())
            la = hmm.getLogLikelihoodArray()
            assert abs(la[0] - r["expected"]) <= 1e-5
            assert m.values[0, 0] == la[0]
        # PairHMMUnitTest.java:510-536 testLikelihoodsFromHaplotypes: empty read list leaves the array untouched
        hmm2 = pairhmm.CudaLoglessPairHMM(args)
        hmm2.initialize([b"A" * 20])
        hmm2.computeLog10Likelihoods(pairhmm.LikelihoodMatrix([b"A" * 20], 0), [], pairhmm.StandardPairHMMInputScoreImputator(100))
        assert hmm2.getLogLikelihoodArray() is None
        rd = pairhmm.Read(b"A" * 10, [20] * 10, [100] * 10, [100] * 10)
        m = pairhmm.LikelihoodMatrix([b"A" * 20], 1)
        hmm2.computeLog10Likelihoods(m, [rd], pairhmm.StandardPairHMMInputScoreImputator(100))
        expected = math.log10((abs(20 - 10 + 1) / 20.0) * 0.99 ** 10)
        assert abs(hmm2.getLogLikelihoodArray()[0] - expected) <= 1e-3
        # haplotype order of the matrix may differ from initialize() order (VectorLoglessPairHMM.java:141-153)
        haps = [b"ACGTACGTACGTAAAC", b"ACGTACGTACGTCCCA", b"TTGTACGTACGTCCCA"]
        hmm2.initialize(haps)
        m = pairhmm.LikelihoodMatrix(haps[::-1], 1)
        rd = pairhmm.Read(b"ACGTACGTACGTA", [30] * 13)
        hmm2.computeLog10Likelihoods(m, [rd], pairhmm.StandardPairHMMInputScoreImputator(10))
        la = hmm2.getLogLikelihoodArray()
        assert [m.values[k, 0] for k in range(3)] == [la[2], la[1], la[0]]
        hmm2.close()
    finally:
        hmm.close()


def _nested_haplotype_unit(seed):
    """haplotypes built to stress prefix sharing: duplicates, nested prefixes, branches at many depths, short haps"""
    rng = np.random.default_rng(seed)
    L = np.frombuffer(b"ACGT", dtype=np.uint8)
    base = L[rng.integers(0, 4, 420)]
    haps = [base.tobytes(), base.tobytes(), base[:300].tobytes(), base[:31].tobytes(), base[:33].tobytes()]
    for d in (1, 20, 31, 32, 33, 47, 64, 65, 100, 200, 250, 399, 419):
        h = base.copy()
        h[d] = L[(np.searchsorted(L, h[d]) + 1) % 4]
        haps.append(h.tobytes())
        h2 = h.copy()
        if d + 40 < 420:
            h2[d + 40] = L[(np.searchsorted(L, h2[d + 40]) + 2) % 4]
            haps.append(h2[: 300 + d // 4].tobytes())
    order = rng.permutation(len(haps))
    haps = [haps[k] for k in order]
    reads = []
    for R in (10, 31, 62, 100, 151, 250, 254, 255, 300):
        off = int(rng.integers(0, 420 - R))
        rd = base[off:off + R].copy()
        rd[R // 2] = ord("A") if rd[R // 2] != ord("A") else ord("G")
        q = np.clip(rng.normal(32, 6, R), 6, 41).astype(np.uint8)
        reads.append((rd, q, const_quals(R, 45), const_quals(R, 45), const_quals(R, 10)))
        reads.append((rd, q, rng.integers(20, 50, R).astype(np.uint8), rng.integers(20, 50, R).astype(np.uint8), rng.integers(5, 30, R).astype(np.uint8)))
    return Batch.single_unit(reads, haps)


def _replicate(b, copies):
    """the same unit `copies` times (same read / haplotype ranges, separate output slots): enough reads per chunk for
    the planner to keep whole haplotype sets together (small chunks are split into haplotype groups instead)"""
    n = b.n_out
    u = np.zeros(copies, dtype=native.UNIT_DTYPE)
    for name in ("read_begin", "read_end", "hap_begin", "hap_end"):
        u[name] = b.units[name][0]
    u["out_off"] = np.arange(copies) * n
    return Batch(b.read_bases, b.base_q, b.ins_q, b.del_q, b.gcp, b.read_off, b.hap_bases, b.hap_off, u), n


def test_prefix_sharing_is_bit_identical_and_correct():
    # haplotype-prefix caching (PairHMMUnitTest.java:607-677 testHaplotypeIndexing pins caching == full recompute)
    with GpuPhmm() as shared, GpuPhmm(no_prefix_sharing=True) as plain:
        for seed in range(3):
            one = _nested_haplotype_unit(seed)
            b, n = _replicate(one, 600)
            shared.reset_stats()
            a = shared.compute(b)
            c = plain.compute(b)
            assert np.array_equal(a, c)
            assert shared.stats()["skipped_cells"] > 0.15 * shared.stats()["cells"]
            assert native.plan_stats(b, True)["snapshots"] > 0
            want = oracle_batch(one)
            for k in (0, 1, 311, 599):
                _check(a[k * n:(k + 1) * n], want, TOL)
            # the copies land in chunks of different size (small ones are split into haplotype groups): float noise only
            assert np.abs(a.reshape(600, n) - a[:n]).max() < 1e-5
            # a single small unit is split into haplotype groups instead (fills the GPU): same numbers to float noise
            solo = shared.compute(one)
            _check(solo, want, TOL)
        b = synth.config2(60)
        a, c = shared.compute(b), plain.compute(b)
        assert np.array_equal(a, c)
        assert plain.stats()["skipped_cells"] == 0
        for seed in range(3):
            b = synth.random_batch(500 + seed, n_units=6, max_haps=12, hap_len=(60, 400), wild_quals=bool(seed % 2))
            assert np.array_equal(shared.compute(b), plain.compute(b))


def test_async_queue_coalesces_and_isolates_errors(hmm):
    # many small regions queued at once are merged into one GPU batch; a bad region fails only its own ticket
    good = [synth.config1(seed=synth.SEED + k) for k in range(12)]
    bad = synth.config1(seed=99)
    bad = Batch(bad.read_bases, bad.base_q, bad.ins_q.copy(), bad.del_q, bad.gcp, bad.read_off, bad.hap_bases, bad.hap_off, bad.units)
    bad.ins_q[3] = 222
    order = good[:5] + [bad] + good[5:]
    tickets = [hmm.submit(b) for b in order]
    results = []
    for t, b in zip(tickets, order):
        if b is bad:
            with pytest.raises(native.GpuPhmmError) as e:
                hmm.wait(t)
            assert e.value.code == native.ERR_BAD_QUAL
        else:
            results.append((hmm.wait(t), b))
    for got, b in results:
        _check(got, oracle_batch(b), TOL)


def test_two_devices_one_process():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    b = synth.config2(200)
    with GpuPhmm(devices=[0]) as one, GpuPhmm(devices=[0, 1], chunk_cells=2_000_000_000) as two:
        assert np.array_equal(one.compute(b), two.compute(b))


def test_two_devices_region_steps_and_queue():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from test_region_steps import _raw_batch
    b, mapq = _raw_batch(77, n_units=40)
    ref = np.zeros(len(b.units), np.int32)
    with GpuPhmm(devices=[0]) as one, GpuPhmm(devices=[0, 1], chunk_cells=20_000_000) as two:
        a, c = one.compute_regions(b, mapq, ref), two.compute_regions(b, mapq, ref)
        for name in ("keep", "base_q", "ins_q", "del_q"):
            assert np.array_equal(a[name], c[name]), name
        assert np.abs(a["lk"] - c["lk"]).max() < 1e-5 and np.abs(a["raw"] - c["raw"]).max() < 1e-5
        tickets = [two.submit_regions(b, mapq, ref) for _ in range(3)]
        for t in tickets:
            assert np.array_equal(two.wait(t)["keep"], a["keep"])


def test_many_flat_quality_classes(hmm):
    # reads with flat (ins, del, gcp) triples: four classes get the constant-coefficient kernel, the rest and the
    # per-base reads fall back to the general kernel; every route must agree with the oracle
    rng = np.random.default_rng(77)
    L = np.frombuffer(b"ACGT", dtype=np.uint8)
    hap = L[rng.integers(0, 4, 300)]
    triples = [(45, 45, 10), (40, 42, 10), (30, 30, 8), (45, 40, 20), (35, 45, 12), (20, 25, 6), (50, 50, 3)]
    reads = []
    for k in range(70):
        R = int(rng.integers(20, 255))
        off = int(rng.integers(0, 300 - R + 1))
        rd = hap[off:off + R].copy()
        rd[R // 3] = ord("T") if rd[R // 3] != ord("T") else ord("A")
        q = np.clip(rng.normal(30, 8, R), 6, 41).astype(np.uint8)
        if k % 8 == 7:
            iq, dq, gq = (rng.integers(20, 50, R).astype(np.uint8) for _ in range(3))
        else:
            t = triples[k % len(triples)]
            iq, dq, gq = const_quals(R, t[0]), const_quals(R, t[1]), const_quals(R, t[2])
        reads.append((rd, q, iq, dq, gq))
    one = Batch.single_unit(reads, [hap.tobytes(), hap[:250].tobytes(), hap[10:].tobytes(), hap[::-1].copy().tobytes()])
    want = oracle_batch(one)
    _check(hmm.compute(one), want, TOL)
    big, n = _replicate(one, 200)   # enough reads for whole-unit tasks with prefix sharing
    got = hmm.compute(big)
    _check(got[:n], want, TOL)
    assert np.abs(got.reshape(200, n) - got[:n]).max() < 1e-5   # copies in differently sized chunks: float noise only


def test_symmetric_quality_reads(hmm):
    # per-base gap-open penalties with ins == del and a flat gcp (what the PCR indel model produces,
    # PairHMMLikelihoodCalculationEngine.java:394-416 / StandardPairHMMInputScoreImputator) take the symmetric-quality
    # kernel; several gcp classes, the overflow to the general kernel, extreme quals and prefix sharing on/off
    rng = np.random.default_rng(78)
    L = np.frombuffer(b"ACGT", dtype=np.uint8)
    hap = L[rng.integers(0, 4, 320)]
    gcps = [10, 10, 10, 12, 8, 20, 3, 10]   # 6 distinct values: more than the kernel has classes
    reads = []
    for k in range(90):
        R = int(rng.integers(1, 300))
        off = int(rng.integers(0, 320 - R + 1))
        rd = hap[off:off + R].copy()
        rd[R // 3] = ord("T") if rd[R // 3] != ord("T") else ord("A")
        q = np.clip(rng.normal(30, 8, R), 6, 41).astype(np.uint8)
        g = np.full(R, 40, np.uint8)
        u = rng.random(R)
        g[u < 0.3] = rng.integers(1, 60, int((u < 0.3).sum())).astype(np.uint8)
        if k % 9 == 0:
            g[:] = rng.integers(0, 128, R).astype(np.uint8)   # the whole legal range, quality 0 included
        if k % 11 == 5:
            g[0] = 2
            g[-1] = 127
        reads.append((rd, q, g, g.copy(), const_quals(R, gcps[k % len(gcps)])))
    haps = [hap.tobytes(), hap[:250].tobytes(), hap[10:].tobytes(), hap[::-1].copy().tobytes(), hap[:33].tobytes()]
    one = Batch.single_unit(reads, haps)
    want = oracle_batch(one)
    _check(hmm.compute(one), want, TOL)
    big, n = _replicate(one, 200)
    with GpuPhmm(no_prefix_sharing=True) as plain:
        got = hmm.compute(big)
        ref = plain.compute(big)
        with GpuPhmm(force_fp64=True) as h64:
            all64 = h64.compute(big)
    # bit-identical, except where a wild-quality read overflows fp32 in one mode only (the state left behind by the
    # previous haplotype differs between a restored snapshot and a full recompute) and is redone in fp64.  A pair that
    # went through the fp64 redo carries exactly the forced-fp64 value (same kernels, same tasks), an fp32 pair never
    # does: so the set of differing pairs must be EXACTLY the pairs redone in one mode and not in the other.
    redone_shared, redone_plain = got == all64, ref == all64
    differs = got != ref
    assert np.array_equal(differs, redone_shared ^ redone_plain)
    assert differs.mean() < 0.005
    assert not differs.any() or np.abs(got - ref)[differs].max() < 1e-5
    _check(got[:n], want, TOL)
    assert np.abs(got.reshape(200, n) - got[:n]).max() < 1e-5   # copies in differently sized chunks: float noise only


def test_likelihoods_near_the_double_underflow_limit(hmm, hmm64):
    # LoglessPairHMM starts from 2^1020/H and still returns finite values down to ~1e-631; the scaled fp64 kernels
    # start 2^60 lower, so pairs below ~1e-550 take the last tier (phmm_exact_f64_kernel: Java's own arithmetic) and
    # must agree with the double-precision oracle far beyond the 1e-4 bar, -inf included
    A, C = ord("A"), ord("C")
    hap = np.full(200, A, np.uint8)
    reads = []
    for n_mis in (100, 120, 128, 132, 136, 139, 141, 150, 153, 155, 156, 157, 158, 200):   # ~ -4 per base: -400 ... -inf
        R = n_mis
        reads.append((np.full(R, C, np.uint8), const_quals(R, 40), const_quals(R, 60), const_quals(R, 60), const_quals(R, 40)))
        wild = np.random.default_rng(R)
        reads.append((np.full(R, C, np.uint8), const_quals(R, 40), wild.integers(50, 70, R).astype(np.uint8),
                      wild.integers(50, 70, R).astype(np.uint8), wild.integers(35, 45, R).astype(np.uint8)))
    b = Batch.single_unit(reads, [hap.tobytes(), hap[:150].tobytes()])
    want = oracle_batch(b)
    assert np.isneginf(want).any() and (want[np.isfinite(want)] < -560).any() and (want > -500).any()
    for h in (hmm, hmm64):
        got = h.compute(b)
        assert np.array_equal(np.isfinite(got), np.isfinite(want))
        deep = np.isfinite(want) & (want < -560)
        assert np.abs(got[deep] - want[deep]).max() < 1e-9
        _check(got, want, TOL)


def test_closed_form_cases_of_the_reference_unit_test():
    # PairHMMUnitTest.java:272-326 (one low-quality mismatch in every position, read centred in / hanging off the
    # haplotype) and :390-418, 511-536 (an all-matching read: the likelihood is just the read's error probability);
    # like the reference's test class with the tristate correction off (:34-38)
    import math
    with GpuPhmm(tristate_off=True) as hmm:
        hap = b"TTCTCTTCTGTTGTGGCTGGTT"
        reads, wants = [], []
        for offset_end in (2, 0):
            L = len(hap) - 2 - offset_end
            for k in range(L):
                quals = const_quals(L, 90)
                quals[k] = 20
                rd = bytearray(hap[2:2 + L])
                rd[k] = ord("T") if rd[k] == ord("C") else ord("C")
                gop = const_quals(L, 80)
                reads.append((bytes(rd), quals, gop, gop, gop))
                wants.append(math.log10(1.0 / len(hap) * (1 - 1e-9) ** (L - 1) * 1e-2))
        got = hmm.compute(Batch.single_unit(reads, [hap]))
        assert np.abs(got - np.array(wants)).max() <= 1e-2
        for read_size in (1, 2, 5, 10):
            for ref_size in (2, 5, 10, 20):
                if ref_size <= read_size:
                    continue
                rd, ref = b"A" * read_size, b"A" * ref_size
                got = hmm.compute(Batch.single_unit([(rd, const_quals(read_size, 20), const_quals(read_size, 100), const_quals(read_size, 100),
                                                      const_quals(read_size, 100))], [ref]))
                want = math.log10((abs(ref_size - read_size + 1) / ref_size) * 0.99 ** read_size)
                assert abs(got[0] - want) <= 1e-3


def test_full_size_config2_sample_against_oracle(hmm):
    # BASELINE.json configs[1] at full size (10 000 regions, ~6 M pairs, 5.8e11 cells): every output is a valid
    # log10 probability, the staged and the device-resident paths agree bit for bit, and a random sample of whole
    # regions matches the double-precision oracle.
    b = synth.config2(10000)
    got = hmm.compute(b)
    assert got.shape == (b.pairs(),)
    assert np.all(np.isfinite(got)) and np.all(got <= 1e-9)
    p = hmm.prepare(b)
    res = np.full(b.n_out, np.nan)
    hmm.run_prepared(p, res)
    hmm.release_prepared(p)
    assert np.array_equal(res, got)
    pick = np.random.default_rng(12).choice(len(b.units), 24, replace=False)
    sub = Batch(b.read_bases, b.base_q, b.ins_q, b.del_q, b.gcp, b.read_off, b.hap_bases, b.hap_off, b.units[pick])
    want = oracle_batch(sub)
    for u in sub.units:
        lo = int(u["out_off"]); n = int((u["read_end"] - u["read_begin"]) * (u["hap_end"] - u["hap_begin"]))
        assert np.abs(got[lo:lo + n] - want[lo:lo + n]).max() <= TOL


def test_several_instances_and_threads_in_one_process():
    # SURVEY 8b "Threading": Spark creates one engine (one PairHMM instance) per task, so several handles live in one
    # process and are driven from different threads (J/tools/HaplotypeCallerSpark.java:175).  Three handles on the same
    # device, one thread each, plus a fourth thread that uses the first handle's queue while its owner calls compute():
    # every synchronous result is bit-identical with the same batch computed alone on a fresh handle.
    import threading
    batches = [synth.random_batch(900 + k, n_units=4) for k in range(6)]
    with GpuPhmm() as solo:
        alone = [solo.compute(b) for b in batches]
    for got, b in zip(alone[:2], batches[:2]):
        _check(got, oracle_batch(b), TOL)
    handles = [GpuPhmm() for _ in range(3)]
    results, errors = {}, []

    def sync_worker(k):
        try:
            for rep in range(3):
                for i, b in enumerate(batches):
                    results[(k, rep, i)] = handles[k].compute(b)
        except Exception as e:   # noqa: BLE001 -- reported by the main thread
            errors.append(e)

    def queue_worker():
        try:
            for rep in range(3):
                tickets = [handles[0].submit(b) for b in batches]
                for i, t in enumerate(tickets):
                    results[("q", rep, i)] = handles[0].wait(t)
        except Exception as e:   # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=sync_worker, args=(k,)) for k in range(3)] + [threading.Thread(target=queue_worker)]
    try:
        for t in threads:
            t.start()
        for t in threads:
            t.join(120)
        assert not any(t.is_alive() for t in threads), "a worker thread hung"
        assert not errors, errors
        assert len(results) == 4 * 3 * len(batches)
        for (who, _, i), got in results.items():
            if who == "q":   # queued batches are merged with whatever else was pending: other chunks, float noise only
                assert np.abs(got - alone[i]).max() < 1e-5
            else:
                assert np.array_equal(got, alone[i])
    finally:
        for h in handles:
            h.close()
