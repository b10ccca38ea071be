"""One handle over several GPUs (the in-process mode of the JNI drop-in, GATK_CUDA_PAIRHMM_DEVICES=0,1,...): end-to-end
GCUPS of gphmm_compute against the number of devices.  Run on a multi-GPU box: python tools/multi_device_probe.py [regions]"""
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch

from gatk_b200 import synth
from gatk_b200.native import GpuPhmm

n_regions = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
b = synth.config2(n_regions, pinned=True)
out = np.zeros(b.n_out)
base = None
for n in [k for k in (1, 2, 4, 8) if k <= torch.cuda.device_count()]:
    with GpuPhmm(devices=list(range(n))) as h:
        h.compute(b, out)
        best = 1e9
        for _ in range(3):
            h.reset_stats()
            t = time.perf_counter()
            h.compute(b, out)
            best = min(best, time.perf_counter() - t)
        cells = h.stats()["cells"]
    base = base or cells / best
    print("%d device(s), one handle: %.1f ms = %.0f GCUPS end to end (%.2f x)" % (n, best * 1e3, cells / best / 1e9, cells / best / base))
