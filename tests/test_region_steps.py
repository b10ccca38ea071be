"""The steps either side of the kernel: modifyReadQualities before, normalizeLikelihoods and
filterPoorlyModeledEvidence after.  CPU part: the oracle against the reference's known-answer tests; GPU part: the
device kernels against the oracle through gphmm_compute_regions (bit-exact for the integer steps and for the post
steps applied to the device's own likelihoods)."""
import numpy as np
import pytest

from gatk_b200 import native, synth
from gatk_b200.native import Batch, GpuPhmm
from oracle import oracle, pyoracle
from phmm_testutil import oracle_batch

# GATKVariantContextUtilsUnitTest.java:922-944 (unit, test string, leading, expected)
REPETITION_KATS = [
    (b"AT", b"GATAT", False, 2), (b"AT", b"GATAT", True, 0), (b"A", b"ATATG", True, 1), (b"AT", b"ATATG", True, 2),
    (b"CCC", b"CCCCCCCC", True, 2), (b"CCC", b"CCCCCCCC", False, 2),
    (b"ATG", b"ATGATGATGATG", True, 4), (b"G", b"ATGATGATGATG", True, 0), (b"T", b"T", True, 1),
    (b"AT", b"ATGATGATCATG", True, 1), (b"CCCCCCCC", b"CCC", True, 0), (b"AT", b"AT", True, 1), (b"AT", b"", True, 0),
    (b"ATG", b"ATGATGATGATG", False, 4), (b"G", b"ATGATGATGATG", False, 1), (b"T", b"T", False, 1),
    (b"AT", b"ATGATGATCATG", False, 0), (b"CCCCCCCC", b"CCC", False, 0), (b"AT", b"AT", False, 1), (b"AT", b"", False, 0),
]
# GATKVariantContextUtilsUnitTest.java:949-962 (unit array, off, len, test array, off, len, leading, expected)
REPETITION_KATS_FULL = [
    (b"XXXATG", 3, 3, b"ATGATGATGATGYYY", 0, 12, True, 4), (b"GGGG", 0, 1, b"GGGGATGATGATGATG", 4, 12, True, 0),
    (b"T", 0, 1, b"TTTTT", 0, 1, True, 1), (b"AT", 0, 2, b"AT", 0, 0, True, 0), (b"AT", 0, 2, b"AT", 1, 0, True, 0),
    (b"AT", 0, 2, b"", 0, 0, True, 0), (b"XXXAT", 3, 2, b"XXXGATAT", 4, 4, False, 2), (b"AT", 0, 2, b"GATAT", 0, 5, False, 2),
    (b"ATG", 0, 3, b"ATGATGATGATG", 0, 12, False, 4), (b"ATG", 0, 3, b"ATGATGATGATGATG", 3, 12, False, 4),
    (b"G", 0, 1, b"ATGATGATGATG", 0, 12, False, 1), (b"G", 0, 1, b"ATGATGATGATGATG", 0, 12, False, 1),
]


def test_find_number_of_repetitions_known_answers():
    for unit, test, leading, want in REPETITION_KATS:
        assert oracle.find_repetitions(unit, test, leading) == want, (unit, test, leading)
        assert pyoracle.find_number_of_repetitions(unit, test, leading) == want, (unit, test, leading)
    for unit, uo, ul, test, to, tl, leading, want in REPETITION_KATS_FULL:
        assert oracle.find_repetitions(unit, test, leading, uo, ul, to, tl) == want, (unit, test, leading)


def test_pcr_error_model_table():
    # getErrorModelAdjustedQual (PairHMMLikelihoodCalculationEngine.java:356-358): max(10, round(40 - exp(n/(f*pi)) + 1))
    c = oracle.pcr_cache(3.0)
    assert list(c[:5]) == [40, 40, 40, 40, 39] and c[20] == 33     # 40-e^0+1, ..., 40-e^(20/9.42)+1 = 32.65
    assert list(oracle.pcr_cache(1.0)[[0, 5, 10, 12, 20]]) == [40, 36, 17, 10, 10]   # HOSTILE bottoms out at MIN_ADJUSTED_QSCORE
    assert all(np.diff(oracle.pcr_cache(2.0).astype(int)) <= 0)


def _low_complexity(rng, n):
    parts = []
    while sum(map(len, parts)) < n:
        unit = bytes(rng.choice(list(b"ACGT"), int(rng.integers(1, 10))).astype(np.uint8))
        parts.append(unit * int(rng.integers(1, 9)))
    return b"".join(parts)[:n]


def test_tandem_repeat_scan_two_restatements_and_pure_repeats():
    # homopolymers: every offset sees the whole run (capped at MAX_REPEAT_LENGTH = 20); no repeat at all: 1
    for n in (2, 5, 19, 20, 21, 60):
        assert [oracle.tandem_repeat_length(b"A" * n, o) for o in range(n)] == [min(n, 20)] * n
    assert [oracle.tandem_repeat_length(b"ACGT", o) for o in range(4)] == [1, 1, 1, 1]
    # the C oracle and the string-based Python restatement agree everywhere, pure repeats of the reference's test
    # (PairHMMLikelihoodCalculationEngineUnitTest.java:101-111) included
    rng = np.random.default_rng(5)
    seqs = [(u * k) for u in (b"A", b"AC", b"ACG", b"ACGT") for k in (1, 2, 3, 5, 10, 15)]
    seqs += [_low_complexity(rng, int(rng.integers(1, 120))) for _ in range(60)]
    for s in seqs:
        for o in range(len(s)):
            assert oracle.tandem_repeat_length(s, o) == pyoracle.find_tandem_repeat_units(s, o)[1], (s, o)


def test_apply_pcr_error_model_like_the_reference_test():
    # PairHMMLikelihoodCalculationEngineUnitTest.java:113-136: flat Q40 in, every base but the last gets the table value
    # of its repeat length
    cache = oracle.pcr_cache(3.0)
    for unit in (b"A", b"AC", b"ACG", b"ACGT"):
        for k in (1, 2, 3, 5, 10, 15):
            s = unit * k
            n = len(s)
            q40 = np.full(n, 40, np.uint8)
            _, iq, dq = oracle.modify_reads(s, q40, q40, q40, np.array([0, n]), [60], 3.0, bq_threshold=0, disable_cap_to_mapq=True)
            want = [cache[pyoracle.find_tandem_repeat_units(s, i - 1)[1]] for i in range(1, n)] + [40]
            assert list(iq) == want and list(dq) == want


def test_cap_minimum_read_qualities():
    # capMinimumReadQualities (PairHMMLikelihoodCalculationEngine.java:300-313): cap by MAPQ, base q < threshold -> 6,
    # ins/del < 6 -> 6; Java bytes are signed, so 200 counts as negative
    s = b"ACGTACGT"
    q = np.array([40, 17, 18, 5, 30, 200, 0, 41], np.uint8)
    i = np.array([45, 5, 6, 0, 200, 45, 45, 45], np.uint8)
    d = np.array([3, 45, 45, 45, 45, 45, 45, 130], np.uint8)
    bq, iq, dq = oracle.modify_reads(s, q, i, d, np.array([0, 8]), [25], 0.0, 18, False)
    assert list(bq) == [25, 6, 18, 6, 25, 25, 6, 25]
    assert list(iq) == [45, 6, 6, 6, 6, 45, 45, 45]
    assert list(dq) == [6, 45, 45, 45, 45, 45, 45, 6]
    bq, _, _ = oracle.modify_reads(s, q, i, d, np.array([0, 8]), [25], 0.0, 18, True)
    assert list(bq) == [40, 6, 18, 6, 30, 6, 6, 41]


def test_normalize_caps_worst_likelihood():
    # AlleleLikelihoodsUnitTest.java:357-387: normalizeLikelihoods(-0.001, true) == max(lk, best - 0.001)
    rng = np.random.default_rng(9)
    for nr, nh in ((1, 1), (7, 2), (40, 5), (3, 16)):
        lk = -rng.random((nr, nh)) * 30
        got = oracle.normalize(lk.ravel(), nr, nh, ref_hap=0, max_diff_cap=-0.001, symmetric=True).reshape(nh, nr)
        want = np.maximum(lk, lk.max(axis=1, keepdims=True) - 0.001) if nh > 1 else lk
        assert np.array_equal(got, want.T)
        # not symmetric: the reference haplotype does not compete for "best" (searchBestAllele, :505-533)
        got = oracle.normalize(lk.ravel(), nr, nh, ref_hap=0, max_diff_cap=-4.5, symmetric=False).reshape(nh, nr)
        if nh > 1:
            best_alt = lk[:, 1:].max(axis=1, keepdims=True)
            assert np.array_equal(got, np.maximum(lk, best_alt - 4.5).T)
        # -inf: untouched
        assert np.array_equal(oracle.normalize(lk.ravel(), nr, nh, 0, -np.inf, True).reshape(nh, nr), lk.T)


def test_filter_poorly_modeled_thresholds():
    # log10MinTrueLikelihood (ReadLikelihoodCalculationEngine.java:95-113): min(2, ceil(0.02 R)) * -4
    q = np.full(300, 30, np.uint8)
    assert oracle.min_true_likelihood(q[:10]) == -4.0 and oracle.min_true_likelihood(q[:50]) == -4.0
    assert oracle.min_true_likelihood(q[:51]) == -8.0 and oracle.min_true_likelihood(q[:300]) == -8.0
    assert oracle.min_true_likelihood(q[:0]) == 0.0
    # dynamic (:66-81, 118-151): min(static uncapped, -0.1 * (sum mean + k sqrt(sum var))); Q30 row of the table
    dyn = lambda k: -0.1 * (150 * 0.039111985 + k * np.sqrt(150 * 1.207526336))
    assert oracle.min_true_likelihood(q[:150], dynamic=True, dynamic_scale=1.0) == -12.0      # static uncapped ceil(3) * -4 wins
    assert dyn(10.0) < -12.0
    assert abs(oracle.min_true_likelihood(q[:150], dynamic=True, dynamic_scale=10.0) - dyn(10.0)) < 1e-12
    assert abs(oracle.min_true_likelihood(q[:20], dynamic=True, dynamic_scale=1.0) - min(-4.0, -0.1 * (20 * 0.039111985 + np.sqrt(20 * 1.207526336)))) < 1e-12
    # AlleleLikelihoodsUnitTest.java:192-217: reads whose best allele is below the threshold are dropped, others stay
    nr, nh = 11, 3
    lk = np.full((nh, nr), -1.0)
    lk[:, ::2] = -200.0
    ro = np.arange(nr + 1) * 20
    keep = oracle.filter_poorly_modeled(lk.ravel(), nr, nh, np.full(nr * 20, 30, np.uint8), ro)
    assert list(keep) == [0, 1] * 5 + [0]


# ---- device ---------------------------------------------------------------------------------------------------------
def _raw_batch(seed, n_units=5, with_bi_bd=False):
    """a batch as modifyReadQualities receives it: raw base quals, flat Q45 (or BI/BD-like) indel quals, low-complexity reads"""
    rng = np.random.default_rng(seed)
    regions, mapq = [], []
    for _ in range(n_units):
        H = int(rng.integers(80, 300))
        hap0 = np.frombuffer(_low_complexity(rng, H), dtype=np.uint8).copy()
        haps = [hap0.tobytes()]
        for _ in range(int(rng.integers(0, 6))):
            h = hap0.copy()
            p = int(rng.integers(0, H))
            h[p] = ord("A") if h[p] != ord("A") else ord("C")
            haps.append(h[: int(rng.integers(max(1, H - 20), H + 1))].tobytes())
        reads = []
        for _ in range(int(rng.integers(1, 40))):
            R = int(rng.integers(1, min(H, 260) + 1))
            off = int(rng.integers(0, H - R + 1))
            rd = hap0[off:off + R].copy()
            err = rng.random(R) < 0.02
            rd[err] = ord("T")
            q = np.clip(rng.normal(28, 10, R), 2, 41).astype(np.uint8)
            if rng.random() < 0.1:
                q[:] = rng.integers(0, 256, R).astype(np.uint8)
                q[q == 255] = 254
            if with_bi_bd:
                iq = rng.integers(0, 60, R).astype(np.uint8)
                dq = rng.integers(0, 60, R).astype(np.uint8)
            else:
                iq = np.full(R, 45, np.uint8)
                dq = iq.copy()
            reads.append((rd, q, iq, dq, np.full(R, 10, np.uint8)))
            mapq.append(int(rng.choice([0, 3, 20, 29, 60, 255])))
        regions.append((reads, haps))
    return Batch.from_units(regions), np.array(mapq, np.uint8)


def _oracle_regions(b, mapq, ref_hap, lk_source=None, **kw):
    """the whole chain on the CPU oracle; lk_source: use these read-major likelihoods instead of the oracle's PairHMM"""
    rate = kw.get("pcr_rate_factor", 3.0)
    q, i, d = oracle.modify_reads(b.read_bases, b.base_q, b.ins_q, b.del_q, b.read_off, mapq, rate,
                                  kw.get("base_quality_score_threshold", 18), kw.get("disable_cap_to_mapq", False))
    mod = Batch(b.read_bases, q, i, d, b.gcp, b.read_off, b.hap_bases, b.hap_off, b.units)
    raw = oracle_batch(mod) if lk_source is None else lk_source
    out = np.full(b.n_out, np.nan)
    keep = np.ones(len(b.read_off) - 1, np.uint8)
    for k, u in enumerate(b.units):
        r0, r1, h0, h1, o = (int(u[x]) for x in ("read_begin", "read_end", "hap_begin", "hap_end", "out_off"))
        nr, nh = r1 - r0, h1 - h0
        if nr == 0 or nh == 0:
            continue
        norm = oracle.normalize(raw[o:o + nr * nh], nr, nh, -1 if ref_hap is None else int(ref_hap[k]),
                                kw.get("log10_global_read_mismapping_rate", -4.5), kw.get("symmetric", False))
        out[o:o + nr * nh] = norm
        if kw.get("filter_poorly", True):
            keep[r0:r1] = oracle.filter_poorly_modeled(norm, nr, nh, q, b.read_off[r0:r1 + 1], kw.get("expected_error_rate_per_base", 0.02),
                                                       kw.get("dynamic_disqualification", False), kw.get("read_disqualification_scale", 1.0))
    return {"lk": out, "keep": keep, "base_q": q, "ins_q": i, "del_q": d, "raw": raw}


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [
    dict(),
    dict(pcr_rate_factor=1.0, symmetric=True, base_quality_score_threshold=10),
    dict(pcr_rate_factor=0.0, disable_cap_to_mapq=True, log10_global_read_mismapping_rate=-0.001),
    dict(dynamic_disqualification=True, read_disqualification_scale=1.5, expected_error_rate_per_base=0.04),
    dict(filter_poorly=False, log10_global_read_mismapping_rate=-np.inf),
])
def test_region_steps_on_device(variant):
    with GpuPhmm() as hmm:
        for seed, bibd in ((1, False), (2, True), (3, False)):
            b, mapq = _raw_batch(seed, n_units=6, with_bi_bd=bibd)
            ref_hap = np.array([(-1 if k % 3 == 2 else k % 2 * 0) for k in range(len(b.units))], np.int32)
            got = hmm.compute_regions(b, mapq, ref_hap, **variant)
            want = _oracle_regions(b, mapq, ref_hap, **variant)
            # integer steps: bit-exact
            for name in ("base_q", "ins_q", "del_q"):
                assert np.array_equal(got[name], want[name]), name
            # the kernel ran on the modified qualities: 1e-4 of the double-precision chain
            fin = np.isfinite(want["lk"])
            assert np.array_equal(np.isfinite(got["lk"]), fin)
            assert np.abs(got["lk"][fin] - want["lk"][fin]).max() <= 1e-4
            # post steps applied by the oracle to the DEVICE's raw likelihoods: bit-exact matrix and keep flags
            # (raw = the un-normalised read-major likelihoods of the same call, what getLogLikelihoodArray() returns)
            mod = Batch(b.read_bases, got["base_q"], got["ins_q"], got["del_q"], b.gcp, b.read_off, b.hap_bases, b.hap_off, b.units)
            rawfin = np.isfinite(got["raw"])
            assert np.abs(got["raw"][rawfin] - hmm.compute(mod)[rawfin]).max() < 1e-5   # same kernels up to the quality-class route
            exact = _oracle_regions(b, mapq, ref_hap, lk_source=got["raw"], **variant)
            assert np.array_equal(got["lk"][fin], exact["lk"][fin])
            assert np.array_equal(got["keep"], exact["keep"])
            # and the flags agree with the all-double chain except for reads within 1e-4 of their threshold
            flips = np.nonzero(got["keep"] != want["keep"])[0]
            assert len(flips) <= 0.01 * len(mapq) + 1


@pytest.mark.gpu
def test_region_steps_realistic_batch_uses_fast_kernels():
    # config-2-like batch with MAPQ 60 and the default CONSERVATIVE model: ins == del after the model, so the
    # symmetric-quality kernel runs; result equals the plain path on pre-modified qualities, transposed and capped
    b = synth.config2(40)
    mapq = np.full(len(b.read_off) - 1, 60, np.uint8)
    with GpuPhmm() as hmm:
        got = hmm.compute_regions(b, mapq, None)
        q, i, d = oracle.modify_reads(b.read_bases, b.base_q, b.ins_q, b.del_q, b.read_off, mapq)
        assert np.array_equal(got["ins_q"], i) and np.array_equal(got["del_q"], d) and np.array_equal(got["base_q"], q)
        assert (i == d).all() and i.max() == 45 and i.min() < 40
        mod = Batch(b.read_bases, q, i, d, b.gcp, b.read_off, b.hap_bases, b.hap_off, b.units)
        assert np.abs(got["raw"] - hmm.compute(mod)).max() < 1e-5
        want = _oracle_regions(b, mapq, None, lk_source=got["raw"])
        assert np.array_equal(got["lk"], want["lk"]) and np.array_equal(got["keep"], want["keep"])
        assert 0 < (got["keep"] == 0).sum() < 0.2 * len(mapq)


@pytest.mark.gpu
def test_region_steps_argument_errors():
    b, mapq = _raw_batch(4, n_units=2)
    with GpuPhmm() as hmm:
        with pytest.raises(native.GpuPhmmError) as e:
            hmm.compute_regions(b, mapq, None, log10_global_read_mismapping_rate=0.5)
        assert e.value.code == native.ERR_INVALID_ARG
        with pytest.raises(native.GpuPhmmError) as e:
            hmm.compute_regions(b, mapq, np.array([99, 0], np.int32))
        assert e.value.code == native.ERR_INVALID_ARG
        # and the handle still works
        assert np.isfinite(hmm.compute_regions(b, mapq, None)["lk"]).all()


@pytest.mark.gpu
def test_region_steps_through_the_plugin_mirror():
    # the call PairHMMLikelihoodCalculationEngine.computeReadLikelihoods would make (java/patches/...regionSteps.patch)
    from gatk_b200.pairhmm import Implementation, LikelihoodMatrix, PairHMMNativeArguments, Read, RegionSteps
    rng = np.random.default_rng(12)
    hap = np.frombuffer(_low_complexity(rng, 200), dtype=np.uint8)
    alt = hap.copy()
    alt[100] = ord("A") if alt[100] != ord("A") else ord("C")
    haps = [hap.tobytes(), alt.tobytes()]
    reads = []
    for k in range(30):
        off = int(rng.integers(0, 100))
        src = alt if k % 2 else hap
        bases = src[off:off + 100].copy()
        if k % 10 == 9:
            bases = np.frombuffer(bytes(rng.choice(list(b"ACGT"), 100).astype(np.uint8)), dtype=np.uint8)   # garbage read: must be dropped
        reads.append(Read(bases.tobytes(), np.full(100, 30, np.uint8), mapping_quality=60 if k % 7 else 10))
    hmm = Implementation.CUDA_LOGLESS_CACHING.makeNewHMM(PairHMMNativeArguments())
    try:
        hmm.initialize(haps)
        matrix = LikelihoodMatrix(list(reversed(haps)), len(reads))     # the matrix may order the alleles differently
        drop = hmm.computeRegionLikelihoods(matrix, reads, 10, RegionSteps(), referenceHaplotype=haps[0])
    finally:
        hmm.close()
    b = Batch.single_unit([(r.bases, r.base_quals, np.full(100, 45, np.uint8), np.full(100, 45, np.uint8), np.full(100, 10, np.uint8)) for r in reads], haps)
    want = _oracle_regions(b, np.array([r.mapping_quality for r in reads], np.uint8), np.array([0], np.int32))
    assert np.abs(matrix.values[::-1].ravel() - want["lk"]).max() <= 1e-4
    # garbage reads and the MAPQ-10 reads (every base quality floored to 6: 0.75^100 < 1e-8) are dropped
    assert drop == [int(k) for k in np.nonzero(want["keep"] == 0)[0]] == [0, 7, 9, 14, 19, 21, 28, 29]
    assert all(np.array_equal(r.hmm_base_qualities, want["base_q"][k * 100:(k + 1) * 100]) for k, r in enumerate(reads))
    assert (reads[0].hmm_base_qualities == 6).all() and (reads[1].hmm_base_qualities == 30).all()   # MAPQ 10 < threshold 18 -> 6


@pytest.mark.gpu
def test_region_steps_async_queue_merges_equal_requests():
    # gphmm_submit_regions: requests with equal parameters are merged into one GPU batch, others run on their own,
    # plain submits can be interleaved, every ticket gets exactly the synchronous result
    with GpuPhmm() as hmm:
        jobs = []
        for k in range(10):
            b, mapq = _raw_batch(100 + k, n_units=2)
            ref = np.zeros(len(b.units), np.int32)
            params = dict(pcr_rate_factor=1.0) if k in (4, 5) else {}
            jobs.append((b, mapq, ref, params))
        tickets = []
        for k, (b, mapq, ref, params) in enumerate(jobs):
            if k == 7:
                tickets.append(("plain", hmm.submit(b)))
            else:
                tickets.append(("regions", hmm.submit_regions(b, mapq, ref, **params)))
        for (kind, t), (b, mapq, ref, params) in zip(tickets, jobs):
            got = hmm.wait(t)
            if kind == "plain":
                assert np.array_equal(got, hmm.compute(b))
                continue
            want = hmm.compute_regions(b, mapq, ref, **params)
            for name in ("base_q", "ins_q", "del_q"):
                assert np.array_equal(got[name], want[name]), name
            assert (got["keep"] != want["keep"]).sum() <= 1   # a read within float noise of its threshold may flip
            assert np.abs(got["raw"] - want["raw"]).max() < 1e-5
            # merged batches may split haplotype groups differently (bigger chunk): same numbers to float rounding
            assert np.abs(got["lk"] - want["lk"]).max() < 1e-5


@pytest.mark.gpu
def test_region_steps_with_forced_double_precision():
    # --native-pair-hmm-use-double-precision: same steps around the fp64 kernels
    b, mapq = _raw_batch(21, n_units=4)
    ref = np.zeros(len(b.units), np.int32)
    with GpuPhmm() as f32, GpuPhmm(force_fp64=True) as f64:
        a, c = f32.compute_regions(b, mapq, ref), f64.compute_regions(b, mapq, ref)
        assert f64.stats()["rescued_pairs"] == f64.stats()["pairs"] > 0
    for name in ("base_q", "ins_q", "del_q"):
        assert np.array_equal(a[name], c[name])
    want = _oracle_regions(b, mapq, ref)
    assert np.abs(c["lk"] - want["lk"]).max() < 1e-9 and np.array_equal(c["keep"], want["keep"])
    assert np.abs(a["lk"] - c["lk"]).max() < 1e-4


@pytest.mark.gpu
def test_region_pipeline_keeps_order_and_results():
    # K regions in flight, finished strictly in order (HC/HaplotypeCaller.java:283-285), same numbers as one call per region
    from gatk_b200.pipeline import RegionPipeline
    regions = list(range(23))
    data = {k: _raw_batch(300 + k, n_units=1) for k in regions}
    fed, finished = [], []

    def feed(k):
        fed.append(k)
        b, mapq = data[k]
        return b, mapq, np.zeros(1, np.int32)

    def drain(k, res):
        finished.append((k, len(fed)))
        return k, res

    with GpuPhmm() as hmm:
        out = list(RegionPipeline(hmm, lookahead=5).run(regions, feed, drain))
        assert [k for k, _ in out] == regions
        # region k was finished only after region k+4 had been fed (look-ahead), except at the end of the stream
        assert all(n_fed >= min(k + 5, len(regions)) for k, n_fed in finished)
        for k, res in out:
            b, mapq = data[k]
            want = hmm.compute_regions(b, mapq, np.zeros(1, np.int32))
            assert np.array_equal(res["keep"], want["keep"]) and np.array_equal(res["base_q"], want["base_q"])
            assert np.abs(res["lk"] - want["lk"]).max() < 1e-5
        with pytest.raises(ValueError):
            RegionPipeline(hmm, lookahead=0)


@pytest.mark.gpu
def test_region_steps_edge_shapes():
    # units without reads / without haplotypes, a zero-length read, 1-base reads, a single haplotype (no normalisation:
    # AlleleLikelihoods.java:427-431), an empty batch
    e = np.zeros(0, np.uint8)
    q1 = lambda n, v: np.full(n, v, np.uint8)
    reads = [(b"A", q1(1, 30), q1(1, 45), q1(1, 45), q1(1, 10)),
             (b"ACGTACGTACGTACGTACGTACGTACGTACGTACGT", q1(36, 25), q1(36, 40), q1(36, 40), q1(36, 10)),
             (b"", e, e, e, e),
             (b"TTTTTTTTTTTTTTTT", q1(16, 12), q1(16, 45), q1(16, 45), q1(16, 10))]
    haps = [b"A", b"ACGT", b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT"]
    one = Batch.single_unit(reads, haps)
    units = np.array([(0, 0, 0, 3, 0), (0, 4, 0, 0, 0), (0, 4, 0, 3, 0), (1, 4, 2, 3, 12), (0, 2, 0, 1, 15)], dtype=native.UNIT_DTYPE)
    b = Batch(one.read_bases, one.base_q, one.ins_q, one.del_q, one.gcp, one.read_off, one.hap_bases, one.hap_off, units)
    mapq = np.array([60, 20, 60, 60], np.uint8)
    ref = np.array([0, -1, 1, 0, 0], np.int32)
    with GpuPhmm() as hmm:
        got = hmm.compute_regions(b, mapq, ref, pcr_rate_factor=2.0)
        q, i, d = oracle.modify_reads(b.read_bases, b.base_q, b.ins_q, b.del_q, b.read_off, mapq, 2.0)
        assert np.array_equal(got["base_q"], q) and np.array_equal(got["ins_q"], i) and np.array_equal(got["del_q"], d)
        for k, u in enumerate(units):
            r0, r1, h0, h1, o = (int(u[x]) for x in ("read_begin", "read_end", "hap_begin", "hap_end", "out_off"))
            nr, nh = r1 - r0, h1 - h0
            if nr == 0 or nh == 0:
                continue
            raw = got["raw"][o:o + nr * nh]
            norm = oracle.normalize(raw, nr, nh, int(ref[k]), -4.5, False)
            assert np.array_equal(got["lk"][o:o + nr * nh], norm, equal_nan=True), k
        # flags of the last unit that owns each read (units must not share reads in production; here 4 writes reads 0-1, 3 writes 1-3)
        last = {0: 4, 1: 4, 2: 3, 3: 3}
        for r, k in last.items():
            u = units[k]
            r0, r1, h0, h1, o = (int(u[x]) for x in ("read_begin", "read_end", "hap_begin", "hap_end", "out_off"))
            nr, nh = r1 - r0, h1 - h0
            norm = oracle.normalize(got["raw"][o:o + nr * nh], nr, nh, int(ref[k]), -4.5, False)
            keep = oracle.filter_poorly_modeled(norm, nr, nh, q, b.read_off[r0:r1 + 1])
            assert got["keep"][r] == keep[r - r0], (r, k)
        empty = Batch(e, e, e, e, e, [0], e, [0], np.zeros(0, dtype=native.UNIT_DTYPE))
        res = hmm.compute_regions(empty, e, None)
        assert res["lk"].size == 0 and res["keep"].size == 0
