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
    """mode 0: no flags; 1: SNP sites; 2: SNPs + well-formed deletions; 3: arbitrary flag bytes; 4: sparse (a handful of
    undetermined sites per haplotype, like the partially determined haplotypes of one assembly region); 5: sparse with
    several deletions of 1-40 columns, some of them adjacent, nested or cut off by the end of the haplotype; 6: sparse SNP
    sites with a deletion every ~12 columns"""
    pd = np.zeros(H, np.uint8)
    if mode == 6:   # a handful of SNP sites, a well-formed deletion of 1-6 columns every ~12 columns (touching ones included)
        for _ in range(int(rng.integers(1, 5))):
            pd[int(rng.integers(0, H))] |= SNP | int(rng.choice([A, C, G, T]))
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
    if mode == 5:
        for _ in range(int(rng.integers(0, 4))):
            pd[int(rng.integers(0, H))] |= SNP | int(rng.choice([A, C, G, T]))
        j = int(rng.integers(0, 30))
        while j < H:
            n = int(rng.choice([1, 1, 2, 5, 12, 40]))
            pd[j] |= DEL_START
            if j + n - 1 < H or rng.random() < 0.5:
                pd[min(H - 1, j + n - 1)] |= DEL_END
            j += n + int(rng.choice([0, 1, 2, 3, 30, 60, 60, 120, 120, 200]))
        return pd
    if mode == 4:
        for _ in range(int(rng.integers(1, 5))):
            pd[int(rng.integers(0, H))] |= SNP | int(rng.choice([A, C, G, T]))
        if H > 12 and rng.random() < 0.5:
            j = int(rng.integers(0, H - 8))
            pd[j] |= DEL_START
            pd[j + int(rng.integers(0, 8))] |= DEL_END
        return pd
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


# ---- the SIMPLE window form of the fast kernels, modelled on the CPU ------------------------------------------------------
def simple_events(pd):
    """host_pdhmm.inl pd_simple_events: the (a, b) column pairs (1-based) of a haplotype whose deletions are well formed
    (they may touch: a flag met in the AFTER_DEL state starts the next event on that very column), or None when a deletion is
    still open or just closed at the last column (LoglessPDPairHMM.java:59 carries that state into the next row)"""
    H, state, a, ev = len(pd), "N", 0, []
    for j in range(1, H + 1):
        ds, de = bool(pd[j - 1] & DEL_START), bool(pd[j - 1] & DEL_END)
        if state in ("N", "A"):
            if ds or de:
                a = j
                if de:
                    ev.append((a, j))
                    state = "A"
                else:
                    state = "I"
            else:
                state = "N"
        elif de:
            ev.append((a, j))
            state = "A"
    return ev if state == "N" else None


def simple_model(hap, pd, read, base_q, ins_q, del_q, gcp):
    """What phmm_pd_fast_kernel<K, SIMPLE> does, column by column in double: the plain LoglessPairHMM update on every column,
    plus, per deletion event (a, b), in this order within a column: before column b + 1 take max(saved, value) of column b IN
    PLACE; on column a SAVE the state (= column a - 1, merged if a is the column after the previous event); after the update
    of column b recompute the insertions from max(saved, value) of the row above.  The likelihood sum adds the unmerged M + I."""
    H, R = len(hap), len(read)
    ev = simple_events(pd)
    assert ev is not None
    first, last, after = {a for a, _ in ev}, {b for _, b in ev}, {b + 1 for _, b in ev}
    eps = lambda q: 10.0 ** (q / -10.0)
    bit = {ord("A"): A, ord("C"): C, ord("G"): G, ord("T"): T}
    M, I, D = np.zeros(R + 1), np.zeros(R + 1), np.zeros(R + 1)
    D[0] = 2.0 ** 1020 / H
    tMM = np.array([0.0] + [max(0.0, 1.0 - eps(ins_q[i]) - eps(del_q[i])) for i in range(R)])
    tMI = np.array([0.0] + [eps(q) for q in ins_q])
    tMD = np.array([0.0] + [eps(q) for q in del_q])
    tII = np.array([0.0] + [eps(q) for q in gcp])
    tIM = 1.0 - tII
    saved, total = None, 0.0
    for j in range(1, H + 1):
        if j in after:
            M, I, D = np.maximum(M, saved[0]), np.maximum(I, saved[1]), np.maximum(D, saved[2])
        if j in first:
            saved = (M.copy(), I.copy(), D.copy())
            saved[2][0] = 0.0  # row 0 has no branch (Java's branch matrices are 0.0 there)
        y, f = hap[j - 1], pd[j - 1]
        prior = np.zeros(R + 1)
        for i in range(1, R + 1):
            x, e = read[i - 1], eps(base_q[i - 1])
            match = x == y or x == ord("N") or y == ord("N") or ((f & SNP) and (f & bit[x]))
            prior[i] = 1.0 - e if match else e / 3.0
        Mn, In, Dn = np.zeros(R + 1), np.zeros(R + 1), np.zeros(R + 1)
        Dn[0] = 2.0 ** 1020 / H
        Mn[1:] = prior[1:] * (M[:-1] * tMM[1:] + I[:-1] * tIM[1:] + D[:-1] * tIM[1:])
        Dn[1:] = M[1:] * tMD[1:] + D[1:] * tII[1:]
        for i in range(1, R + 1):
            um, ui = Mn[i - 1], In[i - 1]
            if j in last:
                um, ui = max(um, saved[0][i - 1]), max(ui, saved[1][i - 1])
            In[i] = um * tMI[i] + ui * tII[i]
        M, I, D = Mn, In, Dn
        total += M[R] + I[R]
    return (np.log10(total) if total > 0 else -np.inf) - np.log10(2.0 ** 1020)


def test_simple_window_form_equals_the_reference_state_machine():
    # the reduction the SIMPLE kernels rest on, checked against the oracle without a GPU
    rng = np.random.default_rng(21)
    n = 0
    assert simple_events(np.array([0, DEL_START, 0, DEL_END, 0, 0], np.uint8)) == [(2, 4)]
    assert simple_events(np.array([0, DEL_END, 0, 0, DEL_START | DEL_END, 0], np.uint8)) == [(2, 2), (5, 5)]
    assert simple_events(np.array([0, DEL_END, 0, DEL_START | DEL_END, 0], np.uint8)) == [(2, 2), (4, 4)]
    assert simple_events(np.array([0, DEL_START, DEL_END, DEL_END, 0], np.uint8)) == [(2, 3), (4, 4)]    # a flag met in AFTER_DEL
    assert simple_events(np.array([0, DEL_END, DEL_START, 0, DEL_END, 0], np.uint8)) == [(2, 2), (3, 5)]
    assert simple_events(np.array([0, 0, DEL_START, 0], np.uint8)) is None                      # still open at the end
    assert simple_events(np.array([0, 0, DEL_START, DEL_END], np.uint8)) is None                # AFTER_DEL carried into the next row
    while n < 150:
        hap, read, q = random_pair(rng, 90, 60)
        pd = random_pd(rng, len(hap), int(rng.choice([2, 3, 4, 5])))   # dense, arbitrary, sparse, touching events
        if simple_events(pd) is None or not simple_events(pd):
            continue
        read[read == ord("N")] = ord("A")
        want = oracle.pd_logless(hap, pd, read, *q)
        got = simple_model(hap, pd, read, *q)
        assert abs(got - want) < 1e-9, (n, got, want)
        n += 1


# ---- device ---------------------------------------------------------------------------------------------------------
def _pd_batch(seed, n_units, max_reads, max_haps, read_len, hap_len, modes=(0, 1, 2, 3)):
    from gatk_b200.native import Batch
    rng = np.random.default_rng(seed)
    regions, pds = [], []
    for _ in range(n_units):
        H = int(rng.integers(hap_len[0], hap_len[1] + 1))
        hap0 = L[rng.integers(0, 4, H)]
        haps = []
        for _ in range(int(rng.integers(1, max_haps + 1))):
            h = hap0[: int(rng.integers(max(1, H - 30), H + 1))].copy()
            haps.append(h)
            pds.append(random_pd(rng, len(h), int(rng.choice(modes))))
        reads = []
        for _ in range(int(rng.integers(1, max_reads + 1))):
            R = int(rng.integers(read_len[0], read_len[1] + 1))
            rd = L[rng.integers(0, 4, R)]
            if rng.random() < 0.8:
                n = min(R, H)
                off = int(rng.integers(0, H - n + 1))
                rd[:n] = hap0[off:off + n]
                err = rng.random(R) < 0.03
                rd[err] = L[rng.integers(0, 4, int(err.sum()))]
            if rng.random() < 0.1:
                rd[int(rng.integers(0, R))] = ord("N")
            reads.append((rd, rng.integers(6, 41, R).astype(np.uint8), rng.integers(10, 60, R).astype(np.uint8),
                          rng.integers(10, 60, R).astype(np.uint8), rng.integers(5, 30, R).astype(np.uint8)))
        regions.append((reads, [h.tobytes() for h in haps]))
    return Batch.from_units(regions), np.concatenate(pds)


def _pd_oracle(b, pd):
    out = np.full(b.n_out, np.nan)
    for u in b.units:
        r0, r1, h0, h1, o = (int(u[x]) for x in ("read_begin", "read_end", "hap_begin", "hap_end", "out_off"))
        for r in range(r0, r1):
            s = slice(int(b.read_off[r]), int(b.read_off[r + 1]))
            for k in range(h0, h1):
                t = slice(int(b.hap_off[k]), int(b.hap_off[k + 1]))
                out[o + (r - r0) * (h1 - h0) + (k - h0)] = oracle.pd_logless(b.hap_bases[t], pd[t], b.read_bases[s], b.base_q[s], b.ins_q[s],
                                                                         b.del_q[s], b.gcp[s])
    return out


def _close(got, want, tol):
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin)
    assert np.array_equal(got[~fin], want[~fin])
    assert np.abs(got[fin] - want[fin]).max() <= tol


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [
    dict(n_units=6, max_reads=12, max_haps=6, read_len=(1, 60), hap_len=(1, 80)),        # K = 2 rows per lane
    dict(n_units=4, max_reads=10, max_haps=5, read_len=(64, 127), hap_len=(100, 200)),   # K = 4
    dict(n_units=3, max_reads=8, max_haps=5, read_len=(128, 256), hap_len=(200, 400)),   # K = 8, one strip
    dict(n_units=2, max_reads=5, max_haps=3, read_len=(257, 700), hap_len=(300, 500)),   # K = 8, several strips
])
def test_pdhmm_matches_oracle(shape):
    from gatk_b200.native import GpuPhmm
    with GpuPhmm() as hmm, GpuPhmm(force_fp64=True) as hmm64:
        for seed in range(3):
            b, pd = _pd_batch(seed, **shape)
            want = _pd_oracle(b, pd)
            _close(hmm.pd_compute(b, pd), want, 1e-4)
            _close(hmm64.pd_compute(b, pd), want, 1e-9)
            assert hmm64.stats()["rescued_pairs"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("read_len", [(1, 94), (95, 158), (159, 254)])  # 3 / 5 / 8 rows per lane
def test_pdhmm_sparse_flags_fast_kernels(read_len):
    # sparsely flagged haplotypes run the fast kernels: plain steps between the deletion events, the SIMPLE window form for
    # well-formed separate deletions (save at the first column, insertion merge at the last, in-place merge after it), the
    # first form for everything else (events next to each other, a deletion open at the end of the haplotype)
    from gatk_b200.native import GpuPhmm
    with GpuPhmm() as hmm:
        for seed in range(3):
            b, pd = _pd_batch(100 + seed, 4, 10, 6, read_len, (150, 420), modes=(4, 5, 5, 6))
            want = _pd_oracle(b, pd)
            _close(hmm.pd_compute(b, pd), want, 1e-4)
        # reads that end inside / right after a deletion window and reads shorter than one lane's rows
        rng = np.random.default_rng(9)
        hap = L[rng.integers(0, 4, 300)]
        pd = np.zeros(300, np.uint8)
        pd[0] |= DEL_START | DEL_END
        pd[40], pd[47] = DEL_START, DEL_END
        pd[100] |= DEL_END                  # a DEL_END without a start: a one-column event
        pd[200], pd[297] = DEL_START, DEL_END
        reads = []
        for R in (1, 2, 7, 8, 9, 39, 40, 41, 64, 94, 95, 150, 158, 159, 200, 247, 248, 249, 254):
            off = int(rng.integers(0, 300 - R + 1))
            reads.append((hap[off:off + R].copy(), rng.integers(6, 41, R).astype(np.uint8), rng.integers(10, 60, R).astype(np.uint8),
                          rng.integers(10, 60, R).astype(np.uint8), rng.integers(5, 30, R).astype(np.uint8)))
        from gatk_b200.native import Batch
        b = Batch.single_unit(reads, [hap.tobytes(), hap[:299].tobytes()])
        pd2 = np.concatenate([pd, pd[:299]])
        _close(hmm.pd_compute(b, pd2), _pd_oracle(b, pd2), 1e-4)


@pytest.mark.gpu
def test_pdhmm_special_cases():
    from gatk_b200 import native
    from gatk_b200.native import Batch, GpuPhmm
    with GpuPhmm() as hmm:
        # no flags at all: the plain PairHMM
        b, pd = _pd_batch(11, 4, 10, 6, (20, 250), (150, 300), modes=(0,))
        assert np.abs(hmm.pd_compute(b, pd) - hmm.compute(b)).max() < 1e-5
        # a flag on the LAST haplotype base carries the state into the next row (LoglessPDPairHMM.java:59)
        rng = np.random.default_rng(5)
        for last in (DEL_START, DEL_END, DEL_START | DEL_END):
            hap = L[rng.integers(0, 4, 50)]
            pd = random_pd(rng, 50, 2)
            pd[-1] |= last
            reads = [(hap[5:45].copy(), np.full(40, 30, np.uint8), np.full(40, 45, np.uint8), np.full(40, 45, np.uint8), np.full(40, 10, np.uint8))]
            b = Batch.single_unit(reads, [hap.tobytes()])
            _close(hmm.pd_compute(b, pd), _pd_oracle(b, pd), 1e-4)
        # hopeless pairs go through the fp64 redo, still Java's numbers
        hap = np.full(300, ord("A"), np.uint8)
        pd = np.zeros(300, np.uint8)
        pd[100], pd[120] = DEL_START, DEL_END
        reads = [(np.full(R, ord("C"), np.uint8), np.full(R, 40, np.uint8), np.full(R, 60, np.uint8), np.full(R, 60, np.uint8), np.full(R, 40, np.uint8))
                 for R in (10, 40, 100, 150)]
        b = Batch.single_unit(reads, [hap.tobytes()])
        hmm.reset_stats()
        got, want = hmm.pd_compute(b, pd), _pd_oracle(b, pd)
        assert hmm.stats()["rescued_pairs"] >= 2 and want.min() < -300
        _close(got, want, 1e-4)
        assert np.abs(got[want < -100] - want[want < -100]).max() < 1e-9
        # a read base that is not ACGT on a SNP column: the reference throws (LoglessPDPairHMM.java:202)
        hap = np.frombuffer(b"ACGTACGTAC", dtype=np.uint8)
        pd = np.zeros(10, np.uint8)
        pd[3] = SNP | A
        bad = Batch.single_unit([(np.frombuffer(b"ACGRACGT", dtype=np.uint8), np.full(8, 30, np.uint8), np.full(8, 45, np.uint8),
                                  np.full(8, 45, np.uint8), np.full(8, 10, np.uint8))], [hap.tobytes()])
        with pytest.raises(native.GpuPhmmError) as e:
            hmm.pd_compute(bad, pd)
        assert e.value.code == native.ERR_INVALID_ARG
        with pytest.raises(ValueError):
            hmm.pd_compute(bad, pd[:5])


@pytest.mark.gpu
def test_pdhmm_plugin_mirror():
    # the call PDPairHMMLikelihoodCalculationEngine makes through the plugin (VectorLoglessPairPDHMM.java:71-147)
    from gatk_b200.pairhmm import (CudaLoglessPairPDHMM, LikelihoodMatrix, PartiallyDeterminedHaplotype, Read,
                                   StandardPairHMMInputScoreImputator)
    rng = np.random.default_rng(8)
    hap = L[rng.integers(0, 4, 120)]
    alleles = [PartiallyDeterminedHaplotype(hap.tobytes(), random_pd(rng, 120, 4)), PartiallyDeterminedHaplotype(hap[:100].tobytes(), random_pd(rng, 100, 2))]
    reads = [Read(hap[o:o + 60].tobytes(), rng.integers(10, 41, 60).astype(np.uint8)) for o in (0, 17, 40, 60)]
    hmm = CudaLoglessPairPDHMM()
    try:
        hmm.computeLog10Likelihoods(LikelihoodMatrix([], 0), [], StandardPairHMMInputScoreImputator(10), alleles)
        assert hmm.getLogLikelihoodArray() is None
        matrix = LikelihoodMatrix([a.bases for a in alleles], len(reads))
        hmm.computeLog10Likelihoods(matrix, reads, StandardPairHMMInputScoreImputator(10), alleles)
    finally:
        hmm.close()
    la = hmm.getLogLikelihoodArray()
    assert la.shape == (8,)
    for r, read in enumerate(reads):
        for a, allele in enumerate(alleles):
            want = oracle.pd_logless(allele.bases, allele.alternate_bases, read.bases, read.base_quals, np.full(60, 45, np.uint8),
                                     np.full(60, 45, np.uint8), np.full(60, 10, np.uint8))
            assert abs(la[r * 2 + a] - want) <= 1e-4 and matrix.values[a, r] == la[r * 2 + a]
