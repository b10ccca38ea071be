"""How far do fp32 PairHMM likelihoods move the step after them: diploid genotype likelihoods (a proxy for VCF identity).

BASELINE.json asks for VCF calls identical with the reference's; that needs a JVM and is not run.
This tool measures the nearest thing that can be measured here: for every region of a configs[1]-like batch the
read x haplotype matrix is computed twice -- by the double-precision oracle (the Java LoglessPairHMM restated) and by the
fp32 path (`--backend gpu`: libgpuphmm on cuda:0; `--backend model`: tests/model/kernel_model.c, the CPU model of the
kernels' arithmetic, for the build box) -- and both go through the reference's next steps:

  normalizeLikelihoods + filterPoorlyModeledEvidence   (oracle/region_steps_oracle.c, pinned to the reference's tests)
  diploid genotype likelihoods over the haplotypes      GenotypeLikelihoodCalculator.computeLog10GenotypeLikelihoods
      hom  a/a: sum_r L[a][r]                            (J/tools/walkers/genotyper/GenotypeLikelihoodCalculator.java:86)
      het  a/b: sum_r approxLog10Sum(L[a][r], L[b][r]) - n*log10(2)                                       (:89-98)
      with MathUtils.approximateLog10SumLog10 and its 1e-4-step Jacobian table (J/utils/MathUtils.java:406-423,467-480)
  PL = round(-10 * (GL - max GL))                        (htsjdk GenotypeLikelihoods.GLsToPLs)

(Haplotypes stand in for alleles: GATK marginalises haplotypes to the alleles of each site first, which takes a max over
rows and cannot increase a difference.)  Reported: regions whose best genotype differs, PL entries that differ, the
largest PL difference, reads whose keep/drop decision differs.  Test infrastructure -- it runs the oracle.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from gatk_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402
from phmm_testutil import oracle_batch  # noqa: E402

_STEP = 1e-4
_TABLE = np.log10(1.0 + np.power(10.0, -np.arange(int(8.0 / _STEP) + 1) * _STEP))


def approx_log10_sum(a, b):
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    with np.errstate(invalid="ignore"):
        diff = np.where(np.isneginf(lo), np.inf, hi - lo)      # a == -inf returns b (:471-472)
    idx = np.floor(np.minimum(diff, 8.0) / _STEP + 0.5).astype(np.int64)   # MathUtils.fastRound for x >= 0
    return hi + np.where(diff < 8.0, _TABLE[idx], 0.0)


def diploid_pls(lk_allele_major, n_reads, n_haps, keep):
    L = lk_allele_major.reshape(n_haps, n_reads)[:, keep.astype(bool)]
    n = L.shape[1]
    gls = []
    for b in range(n_haps):          # canonical order: (0,0) (0,1) (1,1) (0,2) ...
        for a in range(b + 1):
            gls.append(L[a].sum() if a == b else approx_log10_sum(L[a], L[b]).sum() - n * np.log10(2.0))
    gls = np.array(gls)
    return np.rint(-10.0 * (gls - gls.max())).astype(np.int64), gls


def model_backend():
    import test_kernel_model as tkm
    lib = tkm._build(False)
    eps = np.array([oracle.qual_to_error_prob(q) for q in range(256)], dtype=np.float64)
    m2m = np.zeros(((tkm.MAX_Q + 1) * (tkm.MAX_Q + 2)) // 2, dtype=np.float64)
    for mx in range(tkm.MAX_Q + 1):
        for mn in range(mx + 1):
            m2m[((mx * (mx + 1)) >> 1) + mn] = oracle.match_to_match_prob(mn, mx)

    def compute(b):
        out = np.zeros(b.n_out)
        for u in b.units:
            r0, r1, h0, h1, o = (int(u[k]) for k in ("read_begin", "read_end", "hap_begin", "hap_end", "out_off"))
            nh = h1 - h0
            for r in range(r0, r1):
                s, e = int(b.read_off[r]), int(b.read_off[r + 1])
                for h in range(h0, h1):
                    rec = dict(read=bytes(b.read_bases[s:e]), hap=bytes(b.hap_bases[int(b.hap_off[h]):int(b.hap_off[h + 1])]),
                               base_q=b.base_q[s:e], ins_q=b.ins_q[s:e], del_q=b.del_q[s:e], gcp=b.gcp[s:e])
                    out[o + (r - r0) * nh + (h - h0)] = tkm._model(lib, False, (eps, m2m), rec)
        return out
    return compute


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--regions", type=int, default=40)
    ap.add_argument("--backend", choices=("gpu", "model"), default="gpu")
    args = ap.parse_args()
    b = synth.config2(args.regions)
    want = oracle_batch(b)
    if args.backend == "gpu":
        from gatk_b200.native import GpuPhmm
        with GpuPhmm() as hmm:
            got = hmm.compute(b)
    else:
        got = model_backend()(b)
    print("%d regions, %d pairs, backend %s: max |log10 lk - oracle| = %.3g" % (args.regions, len(want), args.backend, np.abs(got - want).max()))
    n_best_diff = n_pl = n_pl_diff = max_pl_diff = n_keep_diff = n_reads_total = 0
    for u in b.units:
        r0, r1, h0, h1, o = (int(u[k]) for k in ("read_begin", "read_end", "hap_begin", "hap_end", "out_off"))
        nr, nh = r1 - r0, h1 - h0
        s, e = int(b.read_off[r0]), int(b.read_off[r1])
        read_off = b.read_off[r0:r1 + 1] - s
        pls = []
        for lk in (want, got):
            norm = oracle.normalize(lk[o:o + nr * nh], nr, nh, 0, -4.5, False)
            keep = oracle.filter_poorly_modeled(norm, nr, nh, b.base_q[s:e], read_off)
            pls.append((diploid_pls(norm, nr, nh, keep)[0], keep))
        (pl_a, keep_a), (pl_b, keep_b) = pls
        n_reads_total += nr
        n_keep_diff += int((keep_a != keep_b).sum())
        n_best_diff += int(np.argmin(pl_a) != np.argmin(pl_b))
        n_pl += len(pl_a)
        d = np.abs(pl_a - pl_b)
        n_pl_diff += int((d > 0).sum())
        max_pl_diff = max(max_pl_diff, int(d.max()))
    print("best diploid genotype differs in %d of %d regions; %d of %d PL entries differ (largest difference %d); "
          "keep/drop differs for %d of %d reads" % (n_best_diff, len(b.units), n_pl_diff, n_pl, max_pl_diff, n_keep_diff, n_reads_total))
    return 0 if n_best_diff == 0 and n_keep_diff == 0 and max_pl_diff <= 1 else 1


if __name__ == "__main__":
    sys.exit(main())
