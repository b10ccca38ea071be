"""Cost of the region steps (modifyReadQualities + normalizeLikelihoods + filterPoorlyModeledEvidence) on the device
next to the same steps on the CPU oracle.  Run on a GPU box: python tools/region_steps_bench.py [regions]"""
import sys
import time

sys.path.insert(0, ".")
import numpy as np

from gatk_b200 import synth
from gatk_b200.native import Batch, GpuPhmm
from oracle import oracle

n_regions = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
b = synth.config2(n_regions)
b = Batch(b.read_bases, b.base_q, b.ins_q, b.del_q, b.gcp, b.read_off, b.hap_bases, b.hap_off, b.units, pinned=True)
n_reads = len(b.read_off) - 1
mapq = np.full(n_reads, 60, np.uint8)

# CPU: the oracle's C restatement, one thread, on a sample of the reads
sample = min(n_reads, 20000)
t0 = time.perf_counter()
q, i, d = oracle.modify_reads(b.read_bases[: b.read_off[sample]], b.base_q, b.ins_q, b.del_q, b.read_off[: sample + 1], mapq)
t_cpu_mod = (time.perf_counter() - t0) / sample
q, i, d = oracle.modify_reads(b.read_bases, b.base_q, b.ins_q, b.del_q, b.read_off, mapq)
mod = Batch(b.read_bases, q, i, d, b.gcp, b.read_off, b.hap_bases, b.hap_off, b.units, pinned=True)

with GpuPhmm() as h:
    def timed(fn, reps=3):
        fn()
        best = 1e9
        for _ in range(reps):
            h.reset_stats()
            t = time.perf_counter()
            r = fn()
            best = min(best, time.perf_counter() - t)
        return best, h.stats(), r

    t_plain, s_plain, raw = timed(lambda: h.compute(mod))
    t_reg, s_reg, res = timed(lambda: h.compute_regions(b, mapq, None, want_quals=False, want_raw=False))
    t_regq, s_regq, _ = timed(lambda: h.compute_regions(b, mapq, None, want_quals=True, want_raw=True))

# CPU post steps on the device's raw likelihoods (bit-exact check + timing)
t0 = time.perf_counter()
ok = True
for u in b.units:
    r0, r1, h0, h1, o = (int(u[x]) for x in ("read_begin", "read_end", "hap_begin", "hap_end", "out_off"))
    nr, nh = r1 - r0, h1 - h0
    if nr == 0 or nh == 0:
        continue
    norm = oracle.normalize(raw[o:o + nr * nh], nr, nh, -1, -4.5, False)
    keep = oracle.filter_poorly_modeled(norm, nr, nh, q, b.read_off[r0:r1 + 1])
    ok = ok and np.array_equal(norm, res["lk"][o:o + nr * nh]) and np.array_equal(keep, res["keep"][r0:r1])
t_cpu_post = time.perf_counter() - t0

cells = s_plain["cells"]
print("regions %d reads %d pairs %d cells %.3g" % (n_regions, n_reads, s_plain["pairs"], cells))
print("gphmm_compute (qualities modified beforehand)      : %.1f ms  e2e %.0f GCUPS  launches %d" % (t_plain * 1e3, cells / t_plain / 1e9, s_plain["kernel_launches"]))
print("gphmm_compute_regions (pre + post steps on device) : %.1f ms  e2e %.0f GCUPS  launches %d" % (t_reg * 1e3, cells / t_reg / 1e9, s_reg["kernel_launches"]))
print("  ... also returning the three modified quality arrays and the raw likelihoods: %.1f ms" % (t_regq * 1e3))
print("CPU oracle (1 thread): modifyReadQualities %.2f us/read = %.1f ms for this batch; normalize+filter (python loop over units) %.1f ms"
      % (t_cpu_mod * 1e6, t_cpu_mod * n_reads * 1e3, t_cpu_post * 1e3))
print("device result == oracle post steps on the device's likelihoods: %s; reads dropped %d of %d" % (ok, int((res["keep"] == 0).sum()), n_reads))
