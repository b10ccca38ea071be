"""Smith-Waterman throughput (gphmm_sw_align) next to the CPU oracle.  Run on a GPU box: python tools/sw_bench.py [pairs]"""
import sys
import time

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np

from gatk_b200.native import GpuPhmm
from oracle import oracle as O
from test_smith_waterman import ALIGNMENT_TO_BEST_HAPLOTYPE, NEW_SW_PARAMETERS, _mutated

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rng = np.random.default_rng(7)


def make(n, ref_len, alt_len, exact_frac):
    refs, alts = [], []
    for _ in range(n):
        nr = int(rng.integers(ref_len[0], ref_len[1] + 1))
        ref = bytes(rng.choice(list(b"ACGT"), nr).astype(np.uint8))
        na = min(nr, int(rng.integers(alt_len[0], alt_len[1] + 1)))
        a = int(rng.integers(0, nr - na + 1))
        alt = ref[a:a + na] if rng.random() < exact_frac else _mutated(rng, ref[a:a + na], int(rng.integers(1, 5)))
        refs.append(ref)
        alts.append(alt)
    return refs, alts


with GpuPhmm() as h:
    for label, params, strategy, shape in (
            ("reads to their best haplotype (250 bp vs 300-500 bp, SOFTCLIP, every read differs from the haplotype)", ALIGNMENT_TO_BEST_HAPLOTYPE, O.SW_SOFTCLIP, ((300, 500), (100, 250), 0.0)),
            ("same, 70 % of the reads match exactly (substring shortcut)", ALIGNMENT_TO_BEST_HAPLOTYPE, O.SW_SOFTCLIP, ((300, 500), (100, 250), 0.7)),
            ("haplotypes to the reference (300-600 bp both, SOFTCLIP)", NEW_SW_PARAMETERS, O.SW_SOFTCLIP, ((300, 600), (300, 600), 0.0))):
        refs, alts = make(n_pairs, *shape)
        cells = sum(len(r) * len(a) for r, a in zip(refs, alts))
        got = h.sw_align(refs, alts, params, strategy, cigar_capacity=128)
        h.reset_stats()
        t = time.perf_counter()
        got = h.sw_align(refs, alts, params, strategy, cigar_capacity=128)
        dt = time.perf_counter() - t
        s = h.stats()
        m = min(300, n_pairs)
        t = time.perf_counter()
        want = [O.sw_align(r, a, params, strategy) for r, a in zip(refs[:m], alts[:m])]
        t_cpu = (time.perf_counter() - t) / m
        assert got[:m] == want
        print("%s\n  %d pairs, %.3g cells: device %.2f ms = %.2f M alignments/s (%.0f GCUPS); wall incl. Python packing %.0f ms; "
              "CPU oracle 1 thread %.0f us per alignment = %.4f M/s; first %d results identical"
              % (label, n_pairs, cells, s["device_ms"], n_pairs / s["device_ms"] / 1e3, cells / s["device_ms"] / 1e6, dt * 1e3, t_cpu * 1e6, 1e-6 / t_cpu, m))
