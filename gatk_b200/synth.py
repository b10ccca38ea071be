"""Synthetic read/haplotype batches of the shapes BASELINE.json names.

Pure input generation (numpy); no likelihood arithmetic.  Seeds follow the reference's
GATK_RANDOM_SEED = 47382911 (src/main/java/org/broadinstitute/hellbender/utils/Utils.java:52).
"""
import numpy as np

from .native import UNIT_DTYPE, Batch

SEED = 47382911
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _base_quals(rng, shape):
    """clipped N(32, 6) in [6, 41], then the HaplotypeCaller floor: < 18 -> 6
    (PairHMMLikelihoodCalculationEngine.java:307-316, --base-quality-score-threshold 18)."""
    q = np.clip(np.rint(rng.normal(32.0, 6.0, shape)), 6, 41).astype(np.uint8)
    q[q < 18] = 6
    return q


def _mutate_hap(rng, hap, n_events):
    """1..n SNPs / short indels applied to a haplotype (uint8 codes 0..3)."""
    h = hap
    for _ in range(n_events):
        # (short haplotypes only occur in the ragged test batches; longer ones keep their historical random stream)
        pos = int(rng.integers(5, len(h) - 5)) if len(h) > 10 else int(rng.integers(0, max(1, len(h) - 1)))
        kind = int(rng.integers(0, 3))
        if kind == 0:
            h = h.copy()
            h[pos] = (h[pos] + rng.integers(1, 4)) % 4
        elif kind == 1:
            h = np.concatenate([h[:pos], rng.integers(0, 4, int(rng.integers(1, 6)), dtype=np.uint8), h[pos:]])
        else:
            h = np.concatenate([h[:pos], h[pos + int(rng.integers(1, 6)):]])
            if len(h) == 0:
                h = hap[:1].copy()
    return h


def _region(rng, n_reads, read_lens, n_haps, hap_len, indel_q=45, gcp_q=10):
    """One active region: haplotypes derived from a random hap0, reads sampled from the haplotypes with
    per-base substitution errors at the rate their base quality states."""
    hap0 = rng.integers(0, 4, hap_len, dtype=np.uint8)
    haps = [hap0] + [_mutate_hap(rng, hap0, int(rng.integers(1, 4))) for _ in range(n_haps - 1)]
    src = rng.integers(0, n_haps, n_reads)
    total = int(np.sum(read_lens))
    bases = np.empty(total, dtype=np.uint8)
    pos = 0
    for r in range(n_reads):
        h = haps[src[r]]
        R = int(read_lens[r])
        if len(h) >= R:
            off = int(rng.integers(0, len(h) - R + 1))
            bases[pos:pos + R] = h[off:off + R]
        else:  # read longer than the haplotype (allowed: PairHMMUnitTest.java:24)
            bases[pos:pos + len(h)] = h
            bases[pos + len(h):pos + R] = rng.integers(0, 4, R - len(h), dtype=np.uint8)
        pos += R
    quals = _base_quals(rng, total)
    err = rng.random(total) < np.power(10.0, quals.astype(np.float64) / -10.0)
    bases = np.where(err, (bases + rng.integers(1, 4, total, dtype=np.uint8)) % 4, bases).astype(np.uint8)
    return haps, _ACGT[bases], quals, np.full(total, indel_q, np.uint8), np.full(total, indel_q, np.uint8), np.full(total, gcp_q, np.uint8)


def _assemble(regions, pinned=False):
    """regions: list of (haps, bases, q, iq, dq, gq, read_lens) -> Batch with one unit per region."""
    rb, bq, iq, dq, gq, hb = [], [], [], [], [], []
    read_lens, hap_lens, units = [], [], []
    r0 = h0 = out = 0
    for haps, bases, q, i, d, g, rl in regions:
        rb.append(bases); bq.append(q); iq.append(i); dq.append(d); gq.append(g)
        read_lens.append(np.asarray(rl, dtype=np.int64))
        for h in haps:
            hb.append(_ACGT[h] if h.dtype == np.uint8 and h.max(initial=0) < 4 else h)
            hap_lens.append(len(h))
        units.append((r0, r0 + len(rl), h0, h0 + len(haps), out))
        out += len(rl) * len(haps)
        r0 += len(rl)
        h0 += len(haps)
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.uint8)
    read_off = np.concatenate([[0], np.cumsum(cat(read_lens) if read_lens else np.zeros(0, np.int64))]).astype(np.int64)
    hap_off = np.concatenate([[0], np.cumsum(np.asarray(hap_lens, dtype=np.int64))]).astype(np.int64)
    u = np.array(units, dtype=UNIT_DTYPE) if units else np.zeros(0, dtype=UNIT_DTYPE)
    return Batch(cat(rb), cat(bq), cat(iq), cat(dq), cat(gq), read_off, cat(hb), hap_off, u, pinned=pinned)


def config1(seed=SEED, pinned=False):
    """BASELINE.json configs[0]: 1 active region, 128 reads (150 bp) x 8 haplotypes (200-300 bp)."""
    rng = np.random.default_rng(seed)
    read_lens = np.full(128, 150, dtype=np.int64)
    haps, b, q, i, d, g = _region(rng, 128, read_lens, 8, int(rng.integers(200, 301)))
    return _assemble([(haps, b, q, i, d, g, read_lens)], pinned=pinned)


def config1_many(n_regions, seed=SEED, pinned=False):
    """n_regions regions of the configs[0] shape (128 reads x 150 bp x 8 haplotypes of 200-300 bp), region k seeded seed + k."""
    regions = []
    for k in range(n_regions):
        rng = np.random.default_rng(seed + k)
        read_lens = np.full(128, 150, dtype=np.int64)
        haps, b, q, i, d, g = _region(rng, 128, read_lens, 8, int(rng.integers(200, 301)))
        regions.append((haps, b, q, i, d, g, read_lens))
    return _assemble(regions, pinned=pinned)


def config2(n_regions=10000, seed=SEED, pinned=False, first_region=0):
    """BASELINE.json configs[1]: 30x WGS-like batch: n_regions active regions, reads/region ~ Poisson(60),
    250 bp reads (10 % clipped to U[100,250]), 4-16 haplotypes of 300-500 bp."""
    regions = []
    for k in range(first_region, first_region + n_regions):
        rng = np.random.default_rng(seed + k)
        n_reads = max(1, int(rng.poisson(60)))
        read_lens = np.where(rng.random(n_reads) < 0.1, rng.integers(100, 251, n_reads), 250).astype(np.int64)
        n_haps = int(rng.integers(4, 17))
        haps, b, q, i, d, g = _region(rng, n_reads, read_lens, n_haps, int(rng.integers(300, 501)))
        regions.append((haps, b, q, i, d, g, read_lens))
    return _assemble(regions, pinned=pinned)


def config4(n_regions=24, reads_per_region=4000, n_haps=32, seed=SEED, pinned=False):
    """BASELINE.json configs[3] at the PairHMM boundary: Mutect2-like 500x depth -- thousands of 150 bp reads
    (10 % clipped to U[100,150]) per region against many haplotypes of 300-500 bp (heavy batching per region)."""
    regions = []
    for k in range(n_regions):
        rng = np.random.default_rng(seed + 104729 * k)
        read_lens = np.where(rng.random(reads_per_region) < 0.1, rng.integers(100, 151, reads_per_region), 150).astype(np.int64)
        haps, b, q, i, d, g = _region(rng, reads_per_region, read_lens, n_haps, int(rng.integers(300, 501)))
        regions.append((haps, b, q, i, d, g, read_lens))
    return _assemble(regions, pinned=pinned)


def config5(hap_len=1000, n_regions=64, reads_per_region=64, n_haps=8, bad_fraction=0.1, seed=SEED, pinned=False):
    """BASELINE.json configs[4]: indel-heavy long-haplotype sweep.  250 bp reads against haplotypes of
    `hap_len`; a `bad_fraction` of the reads carries 3-10 indels of 1-20 bp plus 5 % mismatches at Q40 so
    that their likelihood falls under the fp32 range and the fp64 rescue pass has to redo them."""
    regions = []
    for k in range(n_regions):
        rng = np.random.default_rng(seed + 7919 * k + hap_len)
        read_lens = np.full(reads_per_region, 250, dtype=np.int64)
        haps, b, q, i, d, g = _region(rng, reads_per_region, read_lens, n_haps, hap_len)
        b = b.copy(); q = q.copy()
        for r in np.nonzero(rng.random(reads_per_region) < bad_fraction)[0]:
            seg = b[r * 250:(r + 1) * 250].copy()
            for _ in range(int(rng.integers(3, 11))):
                pos, ln = int(rng.integers(1, 220)), int(rng.integers(1, 21))
                if rng.random() < 0.5:
                    seg = np.concatenate([seg[:pos], _ACGT[rng.integers(0, 4, ln)], seg[pos:]])[:250]
                else:
                    seg = np.concatenate([seg[:pos], seg[pos + ln:], _ACGT[rng.integers(0, 4, ln)]])[:250]
            mm = rng.random(250) < 0.05
            seg = np.where(mm, _ACGT[rng.integers(0, 4, 250)], seg).astype(np.uint8)
            b[r * 250:(r + 1) * 250] = seg
            q[r * 250:(r + 1) * 250] = 40
        regions.append((haps, b, q, i, d, g, read_lens))
    return _assemble(regions, pinned=pinned)


def random_batch(seed, n_units=3, max_reads=12, max_haps=5, read_len=(1, 300), hap_len=(1, 400), wild_quals=False,
                 n_frac=0.02, pinned=False):
    """Small ragged batches for parity tests: arbitrary lengths, optional per-base random indel/gcp quals, some N's."""
    rng = np.random.default_rng(seed)
    regions = []
    for _ in range(n_units):
        n_reads = int(rng.integers(1, max_reads + 1))
        n_haps = int(rng.integers(1, max_haps + 1))
        H = int(rng.integers(hap_len[0], hap_len[1] + 1))
        read_lens = rng.integers(read_len[0], read_len[1] + 1, n_reads).astype(np.int64)
        hap0 = rng.integers(0, 4, H, dtype=np.uint8)
        haps = [hap0] + [(_mutate_hap(rng, hap0, int(rng.integers(1, 4))) if H > 12 else rng.integers(0, 4, int(rng.integers(1, H + 3)), dtype=np.uint8))
                         for _ in range(n_haps - 1)]
        total = int(read_lens.sum())
        bases = np.empty(total, dtype=np.uint8)
        pos = 0
        for r in range(n_reads):
            h = haps[int(rng.integers(0, n_haps))]
            R = int(read_lens[r])
            if len(h) >= R:
                off = int(rng.integers(0, len(h) - R + 1))
                bases[pos:pos + R] = h[off:off + R]
            else:
                bases[pos:pos + len(h)] = h
                bases[pos + len(h):pos + R] = rng.integers(0, 4, R - len(h), dtype=np.uint8)
            pos += R
        if wild_quals:
            q = rng.integers(2, 61, total).astype(np.uint8)
            iq = rng.integers(15, 61, total).astype(np.uint8)
            dq = rng.integers(15, 61, total).astype(np.uint8)
            gq = rng.integers(3, 41, total).astype(np.uint8)
        else:
            q = _base_quals(rng, total)
            iq = np.full(total, 45, np.uint8); dq = np.full(total, 45, np.uint8); gq = np.full(total, 10, np.uint8)
        err = rng.random(total) < np.power(10.0, q.astype(np.float64) / -10.0)
        bases = np.where(err, (bases + rng.integers(1, 4, total, dtype=np.uint8)) % 4, bases).astype(np.uint8)
        letters = _ACGT[bases].copy()
        letters[rng.random(total) < n_frac] = ord("N")
        hap_letters = []
        for h in haps:
            hl = _ACGT[h].copy()
            hl[rng.random(len(hl)) < n_frac] = ord("N")
            hap_letters.append(hl)
        regions.append((hap_letters, letters, q, iq, dq, gq, read_lens))
    return _assemble(regions, pinned=pinned)
