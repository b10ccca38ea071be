import sys; sys.path.insert(0,"."); sys.path.insert(0,"tests")
import numpy as np
from gatk_b200.native import GpuPhmm, Batch
from phmm_testutil import const_quals, oracle_batch
rng = np.random.default_rng(5)
hap = rng.integers(0, 4, 320, dtype=np.uint8)
letters = np.frombuffer(b"ACGT", dtype=np.uint8)
reads = []
for R in list(range(1, 301)):
    off = int(rng.integers(0, 320 - R + 1))
    rd = letters[hap[off:off + R]].copy()
    if R > 3:
        rd[R // 2] = ord("A") if rd[R // 2] != ord("A") else ord("C")
    reads.append((rd, const_quals(R, 30), const_quals(R, 45), const_quals(R, 45), const_quals(R, 10)))
haps=[letters[hap], letters[hap[:200]], letters[hap[7:]]]
b = Batch.single_unit(reads, haps)
print("sorted order:", sorted(range(3), key=lambda k: haps[k].tobytes()))
want=oracle_batch(b)
with GpuPhmm() as s, GpuPhmm(no_prefix_sharing=True) as p:
    a=s.compute(b); c=p.compute(b)
bad=np.nonzero(np.abs(a-want)>1e-4)[0]
print("bad (shared):", [(int(i)//3+1, int(i)%3, round(float(a[i]-want[i]),4)) for i in bad][:60])
bad2=np.nonzero(np.abs(c-want)>1e-4)[0]
print("bad (plain):", [(int(i)//3+1, int(i)%3, round(float(c[i]-want[i]),4)) for i in bad2][:60])
