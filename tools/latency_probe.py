"""Per-call latency of gphmm_compute for one small (region, sample) unit -- the regime a synchronous JNI caller sees."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from gatk_b200 import synth
from gatk_b200.native import GpuPhmm

b = synth.config1(pinned=True)
out = np.zeros(b.n_out)
with GpuPhmm() as h:
    for _ in range(20):
        h.compute(b, out)
    h.reset_stats()
    n = 200
    t = time.perf_counter()
    for _ in range(n):
        h.compute(b, out)
    dt = (time.perf_counter() - t) / n
    s = h.stats()
    print("config1 sync call: %.1f us/call, %.1f GCUPS, launches/call %.1f, device_ms/call %.3f, host_stage_ms/call %.3f" % (
        dt * 1e6, b.cells() / dt / 1e9, s["kernel_launches"] / n, s["device_ms"] / n, s["host_stage_ms"] / n))
    # async queue: 64 regions in flight
    batches = [synth.config1(seed=synth.SEED + k) for k in range(64)]
    outs = [np.zeros(x.n_out) for x in batches]
    for _ in range(2):  # warm-up incl. buffer growth for merged batches
        tk = [h.submit(x, o) for x, o in zip(batches, outs)]
        for t_ in tk:
            h.wait(t_)
    t = time.perf_counter()
    tickets = [h.submit(x, o) for x, o in zip(batches, outs)]
    for tk in tickets:
        h.wait(tk)
    dt = (time.perf_counter() - t) / len(batches)
    print("config1 async queue (64 in flight): %.1f us/region, %.1f GCUPS" % (dt * 1e6, batches[0].cells() / dt / 1e9))
