"""PD-HMM throughput (gphmm_pd_compute) on a configs[1]-shaped batch whose haplotypes carry SNP / deletion flags, next
to the CPU oracle (one thread, C restatement of LoglessPDPairHMM).  Run on a GPU box: python tools/pdhmm_bench.py [regions]"""
import sys
import time

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np

from gatk_b200 import synth
from gatk_b200.native import GpuPhmm
from oracle import oracle
from test_pdhmm import random_pd

n_regions = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
# second argument 150: the configs[0] shape (128 reads x 150 bases x 8 haplotypes of 200-300 bases) instead of configs[1]
b = synth.config1_many(n_regions, pinned=True) if len(sys.argv) > 2 and sys.argv[2] == "150" else synth.config2(n_regions, pinned=True)
rng = np.random.default_rng(1)
results = []
with GpuPhmm() as h:
    for label, mode in (("sparse flags (1-4 SNP sites and at most one deletion per haplotype)", 4), ("the same without the deletions (1-4 SNP sites)", -4),
                        ("1-4 SNP sites, a deletion every ~12 columns (every step a window step of the SIMPLE form)", 6),
                        ("dense flags (12 % SNP columns, a deletion every ~12 columns)", 2)):
        pd = np.concatenate([random_pd(rng, int(b.hap_off[k + 1] - b.hap_off[k]), abs(mode)) for k in range(len(b.hap_off) - 1)])
        if mode < 0:
            pd &= 0xf9  # clear DEL_START / DEL_END
        out = h.pd_compute(b, pd)
        best = 1e9
        for _ in range(3):
            h.reset_stats()
            t = time.perf_counter()
            h.pd_compute(b, pd, out)
            best = min(best, time.perf_counter() - t)
            s = h.stats()
        results.append((label, pd, out.copy(), best, s))
    h.compute(b)
    h.reset_stats()
    t = time.perf_counter()
    h.compute(b)
    t_plain = time.perf_counter() - t
label, pd, out, best, s = results[0]
cells = s["cells"]
# CPU oracle on a sample of pairs
n_done, t0, c_cells, worst = 0, time.perf_counter(), 0, 0.0
for u in b.units[:3]:
    r0, r1, h0, h1, o = (int(u[x]) for x in ("read_begin", "read_end", "hap_begin", "hap_end", "out_off"))
    for r in range(r0, min(r1, r0 + 20)):
        sl = slice(int(b.read_off[r]), int(b.read_off[r + 1]))
        for k in range(h0, h1):
            tl = slice(int(b.hap_off[k]), int(b.hap_off[k + 1]))
            v = oracle.pd_logless(b.hap_bases[tl], pd[tl], b.read_bases[sl], b.base_q[sl], b.ins_q[sl], b.del_q[sl], b.gcp[sl])
            worst = max(worst, abs(v - out[o + (r - r0) * (h1 - h0) + (k - h0)]))
            c_cells += (sl.stop - sl.start) * (tl.stop - tl.start)
            n_done += 1
t_cpu = time.perf_counter() - t0
print("regions %d pairs %d cells %.3g" % (n_regions, s["pairs"], cells))
for label, pd_, _, best_, s_ in results:
    print("gphmm_pd_compute, %s: %.1f ms wall = %.0f GCUPS e2e; device %.1f ms = %.0f GCUPS; %d pairs redone in fp64"
          % (label, best_ * 1e3, cells / best_ / 1e9, s_["device_ms"], cells / s_["device_ms"] / 1e6, s_["rescued_pairs"]))
print("gphmm_compute (plain PairHMM, same reads and haplotypes): %.1f ms wall = %.0f GCUPS" % (t_plain * 1e3, cells / t_plain / 1e9))
print("CPU oracle, 1 thread, %d pairs: %.2f GCUPS; max |GPU - oracle| on them %.2g" % (n_done, c_cells / t_cpu / 1e9, worst))
