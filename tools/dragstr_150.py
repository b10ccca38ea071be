"""One 150-base batch with DRAGstr-like per-base qualities through the prepared path (the half-warp general kernel):
the ncu target of profiles/r02_gen_k10_ncu_full.txt.  Run on a GPU box: python tools/dragstr_150.py"""
import sys
sys.path.insert(0, ".")
import numpy as np
import bench
from gatk_b200 import synth
from gatk_b200.native import GpuPhmm

b = bench._dragstr_like(synth.config1_many(1500, pinned=True), np)
out = np.zeros(b.n_out)
with GpuPhmm() as h:
    p = h.prepare(b)
    for _ in range(2):
        h.run_prepared(p, out)
    h.reset_stats()
    h.run_prepared(p, out)
    s = h.stats()
    h.release_prepared(p)
print("150-base reads, DRAGstr-like qualities: %.0f GCUPS" % (s["cells"] / s["device_ms"] / 1e6))
