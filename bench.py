#!/usr/bin/env python
"""PairHMM GCUPS benchmark (BASELINE.json metric) -- one rank per GPU, weak scaling over region shards.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--regions R] [--impl ours|reference]

A step = one pass of the hot path over one synthetic batch of BASELINE.json configs[1]
(30x WGS-like: `--regions` active regions per GPU, Poisson(60) reads of 250 bp, 4-16 haplotypes of 300-500 bp).
  value    cells/s with the batch resident in HBM (gphmm_run_prepared: kernels + result download), CUDA-event
           time of the step on the launching stream, max over ranks
  e2e      same metric through gphmm_compute with pinned HOST buffers: H2D, kernels, D2H all inside the timer
  roofline FP32-pipe roofline of the fp32 forward kernels: 12 flop/cell (LoglessPairHMM.java:51-55) * cells /
           CUDA-event time of those kernels, against 2*SMs*128*SM-clock observed during the run
  cpu_baseline  the oracle (double-precision restatement of the Java LoglessPairHMM) on the host cores, bounded sample
`--impl reference` times that CPU restatement with all host threads instead (the reference itself is Java plus the
external GKL jar and cannot be built here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOAD = "configs[1]: synthetic 30x WGS-like batch, 250 bp reads x 4-16 haplotypes (300-500 bp), Poisson(60) reads/region"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        # under load = samples in the upper half of the observed power range
        if sm:
            import statistics
            thr = (max(pw) + min(pw)) / 2 if pw else 0
            loaded = [s for s, p in zip(sm, pw) if p >= thr] or sm
            return {"sm_mhz": statistics.median(loaded), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                    "power_w_max": max(pw) if pw else None}
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}


def cpu_baseline(batch, budget_s=15.0, threads=0):
    """Times the CPU oracle on a bounded prefix of the batch's units (about budget_s of CPU work)."""
    import numpy as np
    from oracle import oracle
    from gatk_b200.native import Batch
    from phmm_testutil import oracle_batch
    # every host core the process may use; NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1
    threads = threads or len(os.sched_getaffinity(0))
    # calibrate on a few units, then size the sample
    u = batch.units
    nr = u["read_end"] - u["read_begin"]
    rl = np.diff(batch.read_off); hl = np.diff(batch.hap_off)
    rc = np.concatenate([[0], np.cumsum(rl)]); hc = np.concatenate([[0], np.cumsum(hl)])
    cells_u = (rc[u["read_end"]] - rc[u["read_begin"]]) * (hc[u["hap_end"]] - hc[u["hap_begin"]])
    def run(n):
        sub = Batch(batch.read_bases, batch.base_q, batch.ins_q, batch.del_q, batch.gcp, batch.read_off, batch.hap_bases, batch.hap_off, u[:n])
        t0 = time.perf_counter()
        oracle_batch(sub, threads=threads)
        return time.perf_counter() - t0, int(cells_u[:n].sum())
    n0 = min(len(u), max(1, threads // 4))
    t, c = run(n0)
    rate = c / max(t, 1e-9)
    n = int(min(len(u), max(n0, np.searchsorted(np.cumsum(cells_u), rate * budget_s))))
    t, c = run(n)
    base = {"value": c / t / 1e9, "unit": "GCUPS", "cores": threads, "kind": "port",
            "sample": "first %d of %d regions of the same batch (%.3g cells, %.1f s), oracle/pairhmm_oracle.c fp64 scalar (restated Java LoglessPairHMM), OpenMP over reads" % (n, len(u), c, t)}
    # a second, faster CPU number from the same cores: fp32 SIMD with double redo, the shape of the GKL AVX path
    # (oracle/pairhmm_simd_baseline.c; reported next to the contract's cpu_baseline, never instead of it)
    try:
        un = np.stack([u[k] for k in ("read_begin", "read_end", "hap_begin", "hap_end", "out_off")], axis=1)
        m = len(u)
        t0 = time.perf_counter()
        _, resc = oracle.simd_batch(batch.read_bases, batch.base_q, batch.ins_q, batch.del_q, batch.gcp, batch.read_off,
                                    batch.hap_bases, batch.hap_off, un[:m], batch.n_out, threads=threads)
        ts = time.perf_counter() - t0
        base["simd_port"] = {"value": int(cells_u[:m].sum()) / ts / 1e9, "unit": "GCUPS", "cores": threads, "isa_bits": oracle.simd_isa(),
                             "sample": "%d regions (%.3g cells, %.1f s), oracle/pairhmm_simd_baseline.c: fp32, 16 reads per vector, fp64 redo of %d pairs" % (m, float(cells_u[:m].sum()), ts, resc)}
    except Exception as e:  # the baseline must never break the benchmark
        base["simd_port"] = {"error": str(e)}
    return base, t, c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--regions", type=int, default=int(os.environ.get("GPHMM_BENCH_REGIONS", "10000")))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-prefix-sharing", action="store_true", help="A/B switch: recompute every haplotype from column 1")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    from gatk_b200 import synth

    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import oracle
        batch = synth.config2(min(args.regions, 400))
        per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
        vals = []
        for s in range(args.warmup + args.steps):
            cb, t, c = cpu_baseline(batch, budget_s=per_step)
            if s >= args.warmup:
                vals.append((c, t, cb))
        cells = sum(v[0] for v in vals); secs = sum(v[1] for v in vals)
        v = cells / secs / 1e9
        cb = vals[-1][2]; cb["value"] = v
        print(json.dumps({"impl": "reference", "metric": "PairHMM GCUPS", "value": v, "unit": "GCUPS", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, len(vals)), "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": WORKLOAD, "note": "CPU restatement of the Java LoglessPairHMM (the reference needs a JVM + the GKL jar; neither exists in this image)"},
                          "cpu_baseline": cb, "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback exists)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from gatk_b200 import sharding
    from gatk_b200.native import GpuPhmm

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_gen = time.time()
    first_region, n_regions = sharding.region_slice(rank, world, args.regions)
    batch = synth.config2(n_regions, pinned=True, first_region=first_region)
    t_gen = time.time() - t_gen
    cells, pairs = batch.cells(), batch.pairs()
    out = np.full(batch.n_out, np.nan, dtype=np.float64)
    hmm = GpuPhmm(devices=[local_rank], no_prefix_sharing=args.no_prefix_sharing)
    prepared = hmm.prepare(batch)
    props = torch.cuda.get_device_properties(local_rank)

    # ---- value: inputs resident in HBM ----
    for _ in range(args.warmup):
        hmm.run_prepared(prepared, out)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    hmm.reset_stats()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hmm.run_prepared(prepared, out)
    barrier()
    wall = time.perf_counter() - t0
    st = hmm.stats()
    clocks = sampler.stop()
    dev_s = st["device_ms"] / 1e3
    f32_s = st["device_ms"] / 1e3  # whole device step (forward kernels are > 99 % of it); per-chunk kernel events overlap across streams
    assert np.all(np.isfinite(out)) and np.all(out <= 1e-9), "kernel output is not a valid log10 probability"

    # ---- e2e: host buffers in, host results out, through the public C-ABI call ----
    for _ in range(2):
        hmm.compute(batch, out)
    barrier()
    hmm.reset_stats()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hmm.compute(batch, out)
    barrier()
    e2e_wall = time.perf_counter() - t0
    st2 = hmm.stats()

    (dev_s, wall, e2e_wall, f32_s_max), (cells_all, pairs_all, rescued_all) = sharding.reduce_timing(
        [dev_s, wall, e2e_wall, f32_s], [float(cells), float(pairs), float(st["rescued_pairs"])], device="cuda")

    if rank == 0:
        steps = args.steps
        value = cells_all * steps / dev_s / 1e9
        e2e = cells_all * steps / e2e_wall / 1e9
        sm_mhz = clocks.get("sm_mhz") or float(props.clock_rate) / 1e3 if hasattr(props, "clock_rate") else clocks.get("sm_mhz")
        sm_for_peak = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
        peak_tflops = 2.0 * props.multi_processor_count * 128 * sm_for_peak * 1e6 / 1e12
        achieved_tflops = 12.0 * cells * steps / f32_s / 1e12  # rank 0's fp32 kernels
        res = {
            "metric": "PairHMM GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_s / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "regions_per_gpu": args.regions, "reads_per_gpu": batch.n_reads, "pairs": pairs_all,
                       "cells_per_step": cells_all, "input_bytes_per_gpu": batch.input_bytes(), "l2_policy": "inputs larger than L2 (%.0f MB/GPU streamed per step)" % (batch.input_bytes() / 1e6),
                       "fp64_rescued_pairs_per_step": rescued_all / steps,
                       "prefix_sharing": (not args.no_prefix_sharing), "cells_skipped_by_prefix_sharing_frac": st["skipped_cells"] / max(1, st["cells"]),
                       "cells_definition": "sum of R*H over pairs (LoglessPairHMM.java:47-49), no padding; prefix sharing skips executing some of them, results bit-identical",
                       "timing": "CUDA events on the library's launch stream, max over ranks; wall %.3f s" % wall},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "GCUPS", "h2d_bytes_per_step": st2["h2d_bytes"] / steps, "d2h_bytes_per_step": st2["d2h_bytes"] / steps,
                    "ms_per_step": 1e3 * e2e_wall / steps, "api": "gphmm_compute (C ABI) with pinned host arrays"},
            "gpu_launches": int(st["kernel_launches"]),
            "roofline": {"bound": "fp32", "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved_tflops / peak_tflops,
                         "traffic": 223994368, "kernel": "phmm_flat_f32_kernel<K> (+ phmm_fast_f32_kernel<K> for non-flat reads), all K buckets of the step",
                         "peak_source": "2 x %d SMs x 128 lanes x %.0f MHz (SM clock sampled during the timed region); MEASURED_PEAKS.json has no FP32 entry" % (props.multi_processor_count, sm_for_peak),
                         "flops_per_cell": 12, "kernel_ms_per_step": 1e3 * f32_s / steps,
                         "executed_frac": 12.0 * (cells - st["skipped_cells"] / steps) * steps / f32_s / 1e12 / peak_tflops,
                         "traffic_note": "bytes per launch of phmm_flat_f32_kernel<8> from one ncu --set full capture (22.5 MB DRAM read + 201.5 MB written -- write-back of the snapshot slabs -- in a 5.8 ms launch): the kernel is FP32-issue bound, HBM carries 39 GB/s; profiles/r01_flat_k8_final_ncu_full.txt",
                         "hbm_gbs_staging": (batch.input_bytes() + 8 * pairs) * steps / f32_s / 1e9},
        }
        if not args.no_cpu_baseline and world == 1:  # rank 0 at N=1 only
            small = synth.config2(min(args.regions, 400))
            res["cpu_baseline"], _, _ = cpu_baseline(small, budget_s=args.cpu_budget)
        print(json.dumps(res))
    hmm.release_prepared(prepared)
    hmm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
