import sys,time; sys.path.insert(0,"."); sys.path.insert(0,"tests")
import numpy as np
from gatk_b200 import synth
from gatk_b200.native import GpuPhmm
b = synth.config2(2000, pinned=True)
out=np.zeros(b.n_out)
for nps in (False, True):
    with GpuPhmm(no_prefix_sharing=nps) as h:
        p=h.prepare(b)
        h.run_prepared(p,out); h.run_prepared(p,out); h.reset_stats()
        h.run_prepared(p,out)
        s=h.stats()
        print("no_sharing" if nps else "sharing", "device ms %.2f  fp32 ms %.2f"%(s["device_ms"], s["fp32_kernel_ms"]), "skipped frac %.3f"%(s["skipped_cells"]/s["cells"]), "GCUPS %.0f"%(s["cells"]/s["device_ms"]/1e6))
        h.release_prepared(p)
