"""Differential fuzzing of the CUDA path against the double-precision oracle: random ragged batches with every
quality regime (flat, symmetric, per-base wild), N / exotic bytes, tiny and long reads, forced small chunks, prefix
sharing on/off, forced fp64, and the fused region steps.  Run on a GPU box: python tools/fuzz.py [seconds] [seed0]"""
import sys
import time

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np

from gatk_b200 import synth
from gatk_b200.native import Batch, GpuPhmm
from oracle import oracle
from phmm_testutil import oracle_batch
from test_pdhmm import _pd_oracle, random_pd
from test_smith_waterman import _random_pairs

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
TOL = 1e-4


def requalify(b, rng, mode):
    n = len(b.read_bases)
    ins, dele, gcp = b.ins_q.copy(), b.del_q.copy(), b.gcp.copy()
    if mode == "sym":          # ins == del per base, flat gcp per read
        ins = np.where(rng.random(n) < 0.3, rng.integers(1, 71, n), 40).astype(np.uint8)
        dele = ins.copy()
        for r in range(len(b.read_off) - 1):
            gcp[b.read_off[r]:b.read_off[r + 1]] = int(rng.choice([10, 10, 12, 3, 30]))
    elif mode == "flatmix":    # several flat triples
        for r in range(len(b.read_off) - 1):
            t = [(45, 45, 10), (40, 42, 10), (30, 30, 8), (45, 40, 20), (35, 45, 12), (20, 25, 6)][int(rng.integers(0, 6))]
            s = slice(b.read_off[r], b.read_off[r + 1])
            ins[s], dele[s], gcp[s] = t
    elif mode == "dragstr":    # per-base gap-open and gap-continuation qualities (general kernel; half-warp form: MODE_GEN)
        r = rng.random(n)
        ins = np.where(r < 0.3, rng.integers(12, 41, n), 40).astype(np.uint8)
        dele = ins.copy() if rng.random() < 0.5 else np.minimum(ins + rng.integers(0, 8, n), 60).astype(np.uint8)
        gcp = np.where(r < 0.3, rng.integers(4, 14, n), 10).astype(np.uint8)
    elif mode == "extreme":    # the whole legal range
        ins = rng.integers(0, 128, n).astype(np.uint8)
        dele = rng.integers(0, 128, n).astype(np.uint8)
        gcp = rng.integers(0, 128, n).astype(np.uint8)
    return Batch(b.read_bases, b.base_q, ins, dele, gcp, b.read_off, b.hap_bases, b.hap_off, b.units)


def check(got, want, what):
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin), what
    if fin.any():
        err = np.abs(got[fin] - want[fin]).max()
        assert err <= TOL, "%s: max err %.3g" % (what, err)
        return err
    return 0.0


t_end = time.time() + budget
n_batches = n_pairs = n_pd = n_sw = n_paired = 0
worst = 0.0
handles = {
    "default": GpuPhmm(),
    "small-chunks": GpuPhmm(chunk_cells=3_000_000, host_threads=3),
    "no-sharing": GpuPhmm(no_prefix_sharing=True),
    "fp64": GpuPhmm(force_fp64=True),
}
seed = seed0
try:
    while time.time() < t_end:
        rng = np.random.default_rng(seed)
        shape = int(rng.integers(0, 4))
        if shape == 0:
            b = synth.random_batch(seed, n_units=int(rng.integers(1, 8)), max_reads=20, max_haps=9, wild_quals=bool(rng.integers(0, 2)))
        elif shape == 1:
            b = synth.random_batch(seed, n_units=2, max_reads=6, max_haps=4, read_len=(200, 700), hap_len=(300, 900))
        elif shape == 2:
            b = synth.random_batch(seed, n_units=int(rng.integers(1, 30)), max_reads=4, max_haps=3, read_len=(1, 40), hap_len=(1, 60))
        else:
            lo = int(rng.choice([50, 64, 100]))
            # (a narrow length range gives reads of equal length: quarter-warp tasks, four reads per warp)
            b = synth.random_batch(seed, n_units=3, max_reads=40, max_haps=16, read_len=(lo, lo + 6) if rng.random() < 0.4 else (lo, 254), hap_len=(250, 500))
        mode = ["keep", "sym", "flatmix", "extreme", "dragstr"][int(rng.integers(0, 5))]
        b = requalify(b, rng, mode)
        if rng.random() < 0.2:   # exotic haplotype bytes
            hb = b.hap_bases.copy()
            hb[rng.random(len(hb)) < 0.01] = int(rng.choice([ord("R"), ord("n"), ord("a"), ord("*")]))
            b = Batch(b.read_bases, b.base_q, b.ins_q, b.del_q, b.gcp, b.read_off, hb, b.hap_off, b.units)
        want = oracle_batch(b)
        first = None
        for name, h in handles.items():
            got = h.compute(b)
            worst = max(worst, check(got, want, "seed %d %s %s" % (seed, mode, name)))
            first = got if first is None else first
        if shape == 3 and b.n_reads > 0:
            # half-warp kernels: the same units replicated until one chunk holds enough reads for two reads per warp
            # (prepared path = one chunk); every replica must carry the bits of the one-read-per-warp result
            copies = -(-5000 // b.n_reads)
            u = np.tile(b.units, copies)
            u["out_off"] = np.repeat(np.arange(copies) * b.n_out, len(b.units)) + np.tile(b.units["out_off"], copies)
            big = Batch(b.read_bases, b.base_q, b.ins_q, b.del_q, b.gcp, b.read_off, b.hap_bases, b.hap_off, u)
            hd = handles["default"]
            p = hd.prepare(big)
            out = np.full(big.n_out, np.nan)
            hd.run_prepared(p, out)
            hd.release_prepared(p)
            rep = out.reshape(copies, b.n_out)
            same = np.array_equal(rep, np.broadcast_to(first, rep.shape), equal_nan=True)
            if not same:
                # wild qualities may overflow fp32 in one layout's neighbourhood only (prefix-sharing state): those pairs
                # come back from the fp64 redo; anything else must be bit-identical
                diff = rep != np.broadcast_to(first, rep.shape)
                # general reads -- "dragstr", "extreme", and the classes of "flatmix" beyond the four flat classes a chunk keeps --
                # run different kernels in the two layouts: the same recurrence, the sum taken differently
                assert mode != "keep" and np.abs(rep - first)[diff].max() < 1e-5, "seed %d %s half-warp layout differs" % (seed, mode)
            check(out[:b.n_out], want, "seed %d %s half-warp" % (seed, mode))
            n_paired += 1
        if mode != "extreme":
            # region steps: integer parts bit-exact, matrix/flags bit-exact given the device likelihoods
            n_reads = len(b.read_off) - 1
            mapq = rng.choice([0, 10, 25, 60, 255], n_reads).astype(np.uint8)
            ref = rng.integers(-1, 1, len(b.units)).astype(np.int32)
            kw = dict(pcr_rate_factor=float(rng.choice([0.0, 1.0, 2.0, 3.0])), symmetric=bool(rng.integers(0, 2)),
                      dynamic_disqualification=bool(rng.integers(0, 2)))
            got = handles["default"].compute_regions(b, mapq, ref, **kw)
            q, i, d = oracle.modify_reads(b.read_bases, b.base_q, b.ins_q, b.del_q, b.read_off, mapq, kw["pcr_rate_factor"])
            assert np.array_equal(got["base_q"], q) and np.array_equal(got["ins_q"], i) and np.array_equal(got["del_q"], d), "seed %d modify" % seed
            mod = Batch(b.read_bases, q, i, d, b.gcp, b.read_off, b.hap_bases, b.hap_off, b.units)
            raw = got["raw"]    # the un-normalised likelihoods of the same call
            worst = max(worst, check(raw, oracle_batch(mod), "seed %d regions raw" % seed))
            for k, u in enumerate(b.units):
                r0, r1, h0, h1, o = (int(u[x]) for x in ("read_begin", "read_end", "hap_begin", "hap_end", "out_off"))
                nr, nh = r1 - r0, h1 - h0
                if nr == 0 or nh == 0:
                    continue
                norm = oracle.normalize(raw[o:o + nr * nh], nr, nh, int(ref[k]), -4.5, kw["symmetric"])
                keep = oracle.filter_poorly_modeled(norm, nr, nh, q, b.read_off[r0:r1 + 1], 0.02, kw["dynamic_disqualification"], 1.0)
                assert np.array_equal(norm, got["lk"][o:o + nr * nh]), "seed %d normalize unit %d" % (seed, k)
                assert np.array_equal(keep, got["keep"][r0:r1]), "seed %d keep unit %d" % (seed, k)
        if shape in (0, 2) and mode != "extreme" and b.n_out <= 400:
            # PD-HMM against its oracle (per-pair Python loop: small batches only); exotic haplotype bytes are fine, a
            # non-ACGT read base on a SNP column is the reference's exception -> strip the N's from the reads first
            rb = b.read_bases.copy()
            rb[rb == ord("N")] = ord("A")
            bp = Batch(rb, b.base_q, b.ins_q, b.del_q, b.gcp, b.read_off, b.hap_bases, b.hap_off, b.units)
            pd = np.concatenate([random_pd(rng, int(b.hap_off[k + 1] - b.hap_off[k]), int(rng.integers(0, 7))) for k in range(len(b.hap_off) - 1)])
            want_pd = _pd_oracle(bp, pd)
            for name in ("default", "fp64"):
                worst = max(worst, check(handles[name].pd_compute(bp, pd), want_pd, "seed %d pd %s" % (seed, name)))
            n_pd += 1
        if seed % 4 == 0:
            # Smith-Waterman: offsets and CIGARs identical with the oracle
            lens = [((1, 30), (1, 30)), ((50, 400), (10, 300)), ((200, 700), (150, 600))][int(rng.integers(0, 3))]
            refs, alts = _random_pairs(seed, int(rng.integers(1, 40)), *lens)
            params = [(3, -1, -4, -3), (25, -50, -110, -6), (200, -150, -260, -11), (10, -15, -30, -5),
                      (int(rng.integers(1, 50)), -int(rng.integers(1, 60)), -int(rng.integers(1, 120)), -int(rng.integers(1, 20)))][int(rng.integers(0, 5))]
            strategy = int(rng.integers(0, 4))
            got_sw = handles["default"].sw_align(refs, alts, params, strategy, cigar_capacity=1400)
            want_sw = [oracle.sw_align(r, a, params, strategy) for r, a in zip(refs, alts)]
            assert got_sw == want_sw, "seed %d smith-waterman %s strategy %d" % (seed, params, strategy)
            n_sw += len(refs)
        n_batches += 1
        n_pairs += b.n_out
        seed += 1
finally:
    for h in handles.values():
        h.close()
print("fuzz: %d batches (seeds %d..%d), %d pairs x 4 configurations + region steps, %d batches replicated into the half-warp layout (bit-identical), %d PD-HMM batches, %d Smith-Waterman alignments (bit-exact), worst |err| %.3g (bar %g): OK" % (
    n_batches, seed0, seed - 1, n_pairs, n_paired, n_pd, n_sw, worst, TOL))
