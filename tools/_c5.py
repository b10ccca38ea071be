import sys,time; sys.path.insert(0,"."); sys.path.insert(0,"tests")
import numpy as np
from gatk_b200 import synth
from gatk_b200.native import GpuPhmm, Batch
from phmm_testutil import oracle_batch
with GpuPhmm() as h:
    for H in (500,1000):
        for bad in (0.0,0.1,0.5):
            b=synth.config5(hap_len=H, n_regions=256, reads_per_region=64, n_haps=8, bad_fraction=bad, pinned=True)
            out=np.zeros(b.n_out); p=h.prepare(b); h.run_prepared(p,out); h.run_prepared(p,out); h.reset_stats(); h.run_prepared(p,out); s=h.stats(); h.release_prepared(p)
            sub=Batch(b.read_bases,b.base_q,b.ins_q,b.del_q,b.gcp,b.read_off,b.hap_bases,b.hap_off,b.units[:2])
            want=oracle_batch(sub); n=len(want)
            print("H=%d bad=%.1f GCUPS %.0f rescued %d max err %.2g (low-lk err %.2g)"%(H,bad,s["cells"]/s["device_ms"]/1e6,s["rescued_pairs"],np.abs(out[:n]-want).max(), np.abs(out[:n]-want)[want<-70].max() if (want<-70).any() else 0))
