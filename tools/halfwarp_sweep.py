"""Full-warp against half-warp forward kernels by read length (device-resident GCUPS, flat Q45 gap penalties).
Run on a GPU box: python tools/halfwarp_sweep.py  -- each configuration runs in a fresh process because the switch
(GPHMM_HALFWARP_MAX_READ) is read once per process."""
import os
import subprocess
import sys

CHILD = r'''
import sys
sys.path.insert(0, ".")
import numpy as np
from gatk_b200 import synth
from gatk_b200.native import GpuPhmm
L = int(sys.argv[1])
regions = []
for k in range(1200):
    rng = np.random.default_rng(5000 + k)
    nr = max(1, int(rng.poisson(80)))
    rl = np.full(nr, L, np.int64)
    haps, bs, q, i, d, g = synth._region(rng, nr, rl, int(rng.integers(4, 17)), int(rng.integers(L + 50, L + 250)))
    regions.append((haps, bs, q, i, d, g, rl))
b = synth._assemble(regions, pinned=True)
out = np.zeros(b.n_out)
with GpuPhmm() as h:
    p = h.prepare(b)
    for _ in range(2):
        h.run_prepared(p, out)
    h.reset_stats()
    for _ in range(3):
        h.run_prepared(p, out)
    s = h.stats()
    h.release_prepared(p)
print("%.0f" % (s["cells"] / s["device_ms"] / 1e6))
'''

if __name__ == "__main__":
    print("read length | full warp (GCUPS) | half warp (GCUPS) | rows per lane full / half")
    lengths = [int(x) for x in sys.argv[1:]] or [70, 76, 90, 100, 101, 125, 130, 150, 159, 175, 190, 207, 222, 235, 250]
    for L in lengths:
        res = []
        for mx in ("0", "254"):
            env = dict(os.environ, GPHMM_HALFWARP_MAX_READ=mx)
            r = subprocess.run([sys.executable, "-c", CHILD, str(L)], env=env, capture_output=True, text=True)
            res.append(r.stdout.strip() or ("ERR " + r.stderr[-200:]))
        k16 = next(k for k in (5, 6, 7, 8, 10, 12, 14, 16) if 16 * k >= L + 1)
        print("%d | %s | %s | %d / %d" % (L, res[0], res[1], (L + 1) // 32 + 1, k16), flush=True)
