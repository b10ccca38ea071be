"""BASELINE.json configs[4] (indel-heavy long-haplotype sweep, exercises the fp64 rescue) sharded over the GPUs of one box by
ONE handle (the host work queue of SURVEY 8e; no collective): end-to-end GCUPS through gphmm_compute with pinned host
arrays for 1, 2, 4, 8 devices, results compared bit for bit with the 1-device run.
Run on a multi-GPU box: python tools/config5_multi_gpu.py [regions]"""
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch

from gatk_b200 import synth
from gatk_b200.native import GpuPhmm

n_regions = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
counts = [k for k in (1, 2, 4, 8) if k <= torch.cuda.device_count()]
print("config 5 sharded by one handle over n devices: %d regions x 64 reads (250 bp) x 8 haplotypes; end-to-end GCUPS (best of 3)" % n_regions)
print("| haplotype length | indel-heavy reads | pairs redone in fp64 | " + " | ".join("%d GPU" % k for k in counts) + " | identical results |")
print("|---|---|---|" + "---|" * (len(counts) + 1))
for H, bad in ((500, 0.1), (1000, 0.0), (1000, 0.1), (1000, 0.5)):
    if True:
        b = synth.config5(hap_len=H, n_regions=n_regions, reads_per_region=64, n_haps=8, bad_fraction=bad, pinned=True)
        cells = b.cells()
        ref, rates, same, redo = None, [], True, 0
        for n in counts:
            out = np.zeros(b.n_out)
            with GpuPhmm(devices=list(range(n))) as h:
                h.compute(b, out)
                best = 1e9
                for _ in range(3):
                    h.reset_stats()
                    t = time.perf_counter()
                    h.compute(b, out)
                    best = min(best, time.perf_counter() - t)
                redo = h.stats()["rescued_pairs"]
            rates.append(cells / best / 1e9)
            if ref is None:
                ref = out.copy()
            else:
                same = same and bool(np.array_equal(ref, out))
        print("| %d | %.0f %% | %d (%.1f %%) | %s | %s |" % (H, 100 * bad, redo, 100.0 * redo / b.pairs(), " | ".join("%.0f" % r for r in rates), "yes" if same else "NO"), flush=True)
