"""GCUPS on batches whose gap-open qualities look like HaplotypeCaller's default --pcr-indel-model CONSERVATIVE output
(ins == del per base, mostly Q40 with lower values at tandem repeats, flat gcp) next to the flat-Q45 batches of
BASELINE.json: 250 bp (configs[1] shape) and 150 bp reads.  Run on a GPU box: python tools/realistic_quals.py"""
import sys,time; sys.path.insert(0,"."); sys.path.insert(0,"tests")
import numpy as np
from gatk_b200 import synth
from gatk_b200.native import GpuPhmm, Batch
def pcr_like(b, seed=3):
    rng=np.random.default_rng(seed); n=len(b.read_bases)
    q=np.full(n,40,np.uint8)
    r=rng.random(n); q[r<0.12]=39; q[r<0.04]=38; q[r<0.015]=rng.integers(25,38,int((r<0.015).sum())).astype(np.uint8)
    return Batch(b.read_bases,b.base_q,q,q.copy(),b.gcp,b.read_off,b.hap_bases,b.hap_off,b.units,pinned=True)
def dragstr_like(b, seed=4):
    # DragstrPairHMMInputScoreImputator.java:57-66: gop = ins = del per base (<= 40) and a per-base gcp, both from the STR context
    rng=np.random.default_rng(seed); n=len(b.read_bases)
    r=rng.random(n)
    gop=np.full(n,40,np.uint8); gop[r<0.25]=rng.integers(30,40,int((r<0.25).sum())).astype(np.uint8); gop[r<0.05]=rng.integers(15,30,int((r<0.05).sum())).astype(np.uint8)
    gcp=np.full(n,10,np.uint8); gcp[r<0.25]=rng.integers(6,12,int((r<0.25).sum())).astype(np.uint8)
    return Batch(b.read_bases,b.base_q,gop,gop.copy(),gcp,b.read_off,b.hap_bases,b.hap_off,b.units,pinned=True)
out=None
with GpuPhmm() as h:
    for name,b in (("250bp config2 x2000", synth.config2(2000)),):
        for label,bb in (("flat Q45", Batch(b.read_bases,b.base_q,b.ins_q,b.del_q,b.gcp,b.read_off,b.hap_bases,b.hap_off,b.units,pinned=True)), ("PCR-model-like", pcr_like(b)), ("DRAGstr-like (per-base gop and gcp: general kernel)", dragstr_like(b))):
            out=np.zeros(bb.n_out); p=h.prepare(bb); h.run_prepared(p,out); h.run_prepared(p,out); h.reset_stats(); h.run_prepared(p,out); s=h.stats(); h.release_prepared(p)
            print(name,label,"GCUPS %.0f"%(s["cells"]/s["device_ms"]/1e6))
    # 150 bp reads
    regions=[]
    for k in range(1500):
        rng=np.random.default_rng(1000+k); nr=max(1,int(rng.poisson(100))); rl=np.full(nr,150,np.int64)
        haps,bs,q,i,d,g=synth._region(rng,nr,rl,int(rng.integers(4,17)),int(rng.integers(250,401)))
        regions.append((haps,bs,q,i,d,g,rl))
    b=synth._assemble(regions)
    for label,bb in (("flat Q45", Batch(b.read_bases,b.base_q,b.ins_q,b.del_q,b.gcp,b.read_off,b.hap_bases,b.hap_off,b.units,pinned=True)), ("PCR-model-like", pcr_like(b)), ("DRAGstr-like (per-base gop and gcp: general kernel)", dragstr_like(b))):
        out=np.zeros(bb.n_out); p=h.prepare(bb); h.run_prepared(p,out); h.run_prepared(p,out); h.reset_stats(); h.run_prepared(p,out); s=h.stats(); h.release_prepared(p)
        print("150bp x1500 regions",label,"GCUPS %.0f"%(s["cells"]/s["device_ms"]/1e6))
