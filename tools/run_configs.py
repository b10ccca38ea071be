#!/usr/bin/env python
"""Runs the BASELINE.json configs that need only the PairHMM boundary (1, 2, 5 and a Mutect2-like deep-coverage batch)
through the C ABI on one GPU and prints a markdown table (GCUPS resident / end-to-end, rescued pairs, max error against
the CPU oracle on a sample).  usage: python tools/run_configs.py [--quick]"""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from gatk_b200 import synth
from gatk_b200.native import GpuPhmm, Batch
from phmm_testutil import oracle_batch

quick = "--quick" in sys.argv


def deep_coverage(n_regions=24, reads=4000, n_haps=32, seed=synth.SEED):
    """configs[3] at the PairHMM boundary (synth.config4)"""
    return synth.config4(n_regions, reads, n_haps, seed=seed, pinned=True)


def measure(name, batch, hmm, sample_units=4, steps=3):
    out = np.zeros(batch.n_out)
    p = hmm.prepare(batch)
    for _ in range(2):
        hmm.run_prepared(p, out)
    hmm.reset_stats()
    for _ in range(steps):
        hmm.run_prepared(p, out)
    st = hmm.stats()
    hmm.release_prepared(p)
    resident = st["cells"] / st["device_ms"] / 1e6
    for _ in range(2):
        hmm.compute(batch, out)
    t = time.perf_counter()
    for _ in range(steps):
        hmm.compute(batch, out)
    e2e = batch.cells() * steps / (time.perf_counter() - t) / 1e9
    sub = Batch(batch.read_bases, batch.base_q, batch.ins_q, batch.del_q, batch.gcp, batch.read_off, batch.hap_bases, batch.hap_off, batch.units[:sample_units])
    want = oracle_batch(sub)
    got = np.full_like(want, np.nan)
    for u in sub.units:
        n = int((u["read_end"] - u["read_begin"]) * (u["hap_end"] - u["hap_begin"]))
        got[int(u["out_off"]):int(u["out_off"]) + n] = out[int(u["out_off"]):int(u["out_off"]) + n]
    fin = np.isfinite(want)
    err = float(np.abs(got[fin] - want[fin]).max())
    print("| %s | %d | %.3g | %.0f | %.0f | %.2g | %d (%.1f %%) | %.1f %% |" % (
        name, batch.pairs(), batch.cells(), resident, e2e, err, st["rescued_pairs"] // steps,
        100.0 * st["rescued_pairs"] / max(1, st["pairs"]), 100.0 * st["skipped_cells"] / max(1, st["cells"])), flush=True)


print("| config | pairs | cells | GCUPS resident | GCUPS end-to-end | max err vs oracle (sample) | pairs redone in fp64 | cells skipped (prefix sharing) |")
print("|---|---|---|---|---|---|---|---|")
with GpuPhmm() as h:
    measure("1: 1 region, 128 x 150 bp x 8 hap", synth.config1(pinned=True), h, sample_units=1, steps=20)
    measure("2: %d regions, 250 bp x 4-16 hap" % (1000 if quick else 10000), synth.config2(1000 if quick else 10000, pinned=True), h)
    measure("4 (boundary only): 24 regions x 4000 reads x 32 hap", deep_coverage(6 if quick else 24), h, sample_units=1)
    for H in (250, 500, 750, 1000):
        for bad in (0.0, 0.1, 0.5):
            measure("5: H=%d, %.0f %% indel-heavy reads" % (H, 100 * bad), synth.config5(hap_len=H, n_regions=64 if quick else 256, reads_per_region=64, n_haps=8, bad_fraction=bad, pinned=True), h, sample_units=2)
with GpuPhmm(force_fp64=True) as h:
    measure("2 (forced fp64): 300 regions", synth.config2(300, pinned=True), h)
