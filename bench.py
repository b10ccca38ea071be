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
  e2e_from_pageable  the same call with ordinary (pageable) host arrays: the library copies them into pinned bounce buffers
           first -- the copy a JVM caller pays when the JNI shim packs heap arrays (gpuphmm_jni.cpp pack())
  read_150bp  the device-resident metric on 1 500 regions of the configs[0] shape (128 reads x 150 bp x 8 haplotypes),
           flat Q45 gap penalties, PCR-indel-model-like ones (ins == del per base) and DRAGstr-like ones (per-base gap-open and
           gap-continuation: the general kernel) -- the read length of configs[0], [2], [3]
  one_handle  (N > 1, rank 0) ONE gphmm handle over all N devices on the N x regions batch of all ranks: the in-process
           multi-GPU mode a JVM gets (host work queue, one stream set per GPU, no collective)
`--impl reference` times the CPU restatements with all host threads instead: value = the fp32 AVX-512 port with double redo
(the shape of GKL's AVX_LOGLESS_CACHING_OMP, which BASELINE.json's metric names), scalar_fp64_port = the restated Java
LoglessPairHMM.  The reference itself is Java plus the external GKL jar; no JDK exists in this image or on the GPU box
(profiles/r02_jdk_probe.txt), see DESIGN.md.
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


def _pcr_like(b, np, seed=3):
    """gap-open qualities as HaplotypeCaller's default --pcr-indel-model CONSERVATIVE leaves them: ins == del per base, mostly
    Q40, lower at (pretend) tandem repeats, flat gcp (PairHMMLikelihoodCalculationEngine.java:361-371)"""
    from gatk_b200.native import Batch
    rng = np.random.default_rng(seed)
    n = len(b.read_bases)
    q = np.full(n, 40, np.uint8)
    r = rng.random(n)
    q[r < 0.12] = 39
    q[r < 0.04] = 38
    q[r < 0.015] = rng.integers(25, 38, int((r < 0.015).sum())).astype(np.uint8)
    return Batch(b.read_bases, b.base_q, q, q.copy(), b.gcp, b.read_off, b.hap_bases, b.hap_off, b.units, pinned=True)


def _dragstr_like(b, np, seed=4):
    """per-base gap-open (ins == del) AND gap-continuation qualities from the STR context, as --dragstr-params-path leaves them
    (DragstrPairHMMInputScoreImputator.java:57-66): the general kernel's input"""
    from gatk_b200.native import Batch
    rng = np.random.default_rng(seed)
    n = len(b.read_bases)
    r = rng.random(n)
    gop = np.full(n, 40, np.uint8)
    gop[r < 0.25] = rng.integers(30, 40, int((r < 0.25).sum())).astype(np.uint8)
    gop[r < 0.05] = rng.integers(15, 30, int((r < 0.05).sum())).astype(np.uint8)
    gcp = np.full(n, 10, np.uint8)
    gcp[r < 0.25] = rng.integers(6, 12, int((r < 0.25).sum())).astype(np.uint8)
    return Batch(b.read_bases, b.base_q, gop, gop.copy(), gcp, b.read_off, b.hap_bases, b.hap_off, b.units, pinned=True)


def measure_read150(hmm, synth, np, n_regions=1500, steps=3):
    """device-resident GCUPS on n_regions regions of the configs[0] shape (128 reads x 150 bp x 8 haplotypes of 200-300 bp)"""
    flat = synth.config1_many(n_regions, pinned=True)
    res = {"workload": "configs[0] shape x %d regions: 128 reads x 150 bp x 8 haplotypes (200-300 bp); inputs resident in HBM, CUDA-event time" % n_regions,
           "cells_per_step": flat.cells(), "unit": "GCUPS"}
    for name, b in (("flat_q45", flat), ("pcr_model_quals", _pcr_like(flat, np)), ("dragstr_quals", _dragstr_like(flat, np))):
        out = np.zeros(b.n_out)
        p = hmm.prepare(b)
        for _ in range(2):
            hmm.run_prepared(p, out)
        hmm.reset_stats()
        for _ in range(steps):
            hmm.run_prepared(p, out)
        s = hmm.stats()
        hmm.release_prepared(p)
        res[name] = s["cells"] / s["device_ms"] / 1e6
    return res


def measure_one_handle(whole, outs, world, steps, own_e2e, np, torch):
    """ONE gphmm handle over devices 0..world-1 on the batch of all ranks (rank 0 only)."""
    from gatk_b200.native import GpuPhmm
    out = np.full(whole.n_out, np.nan, dtype=np.float64)
    cells = whole.cells()
    with GpuPhmm(devices=list(range(world))) as h:
        prepared = h.prepare(whole)
        for _ in range(2):
            h.run_prepared(prepared, out)
        h.reset_stats()
        for _ in range(steps):
            h.run_prepared(prepared, out)
        st = h.stats()
        h.release_prepared(prepared)
        value = cells * steps / (st["device_ms"] / 1e3) / 1e9
        for _ in range(2):
            h.compute(whole, out)
        for d in range(world):
            torch.cuda.synchronize(d)
        h.reset_stats()
        t0 = time.perf_counter()
        for _ in range(steps):
            h.compute(whole, out)
        wall = time.perf_counter() - t0
        st2 = h.stats()
    ref = np.concatenate(outs)
    same = bool(ref.shape == out.shape and np.array_equal(ref, out))
    e2e = cells * steps / wall / 1e9
    return {"value": value, "e2e": e2e, "unit": "GCUPS", "devices": world, "regions": int(len(whole.units)), "cells_per_step": cells,
            "ms_per_step_e2e": 1e3 * wall / steps, "efficiency": e2e / (world * own_e2e),
            "efficiency_definition": "one_handle.e2e / (N x rank 0's own single-GPU e2e in this run: %.0f GCUPS)" % own_e2e,
            "bit_identical_with_per_rank_results": same,
            "max_abs_diff_vs_per_rank": None if same or ref.shape != out.shape else float(np.nanmax(np.abs(ref - out))),
            "host_stage_ms_per_step": st2["host_stage_ms"] / steps, "h2d_bytes_per_step": st2["h2d_bytes"] / steps, "d2h_bytes_per_step": st2["d2h_bytes"] / steps,
            "api": "one gphmm_t over all devices (gphmm_config.devices), gphmm_compute with pinned host arrays; no collective on the data path"}


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
    ap.add_argument("--no-read150", action="store_true", help="skip the 150 bp side measurement")
    ap.add_argument("--no-one-handle", action="store_true", help="N > 1: skip the one-handle-over-all-devices measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    from gatk_b200 import synth

    if args.impl == "reference":
        if rank != 0:
            return
        import numpy as np
        from oracle import oracle
        threads = len(os.sched_getaffinity(0))
        # bounded sample of the same workload: the first regions of configs[1], sized so that one step of the SIMD port
        # takes a few seconds (the whole run must end within minutes)
        n_sample = min(args.regions, int(os.environ.get("GPHMM_REF_REGIONS", "1200")))
        batch = synth.config2(n_sample)
        cells = batch.cells()
        u = batch.units
        un = np.stack([u[k] for k in ("read_begin", "read_end", "hap_begin", "hap_end", "out_off")], axis=1)
        times, resc = [], 0
        for s_ in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            _, resc = oracle.simd_batch(batch.read_bases, batch.base_q, batch.ins_q, batch.del_q, batch.gcp, batch.read_off,
                                        batch.hap_bases, batch.hap_off, un, batch.n_out, threads=threads)
            if s_ >= args.warmup:
                times.append(time.perf_counter() - t0)
        secs = sum(times)
        v = cells * len(times) / secs / 1e9
        scalar, t_sc, c_sc = cpu_baseline(synth.config2(min(n_sample, 400)), budget_s=min(args.cpu_budget, 10.0), threads=threads)
        scalar.pop("simd_port", None)
        sample = "first %d of the %d regions of configs[1] per step (%.3g cells)" % (n_sample, args.regions, cells)
        cb = {"value": v, "unit": "GCUPS", "cores": threads, "kind": "port", "isa_bits": oracle.simd_isa(),
              "sample": sample + "; oracle/pairhmm_simd_baseline.c: fp32, 16 reads per vector, fp64 redo of %d pairs -- a port with the shape of GKL AVX_LOGLESS_CACHING_OMP, not GKL" % resc}
        print(json.dumps({"impl": "reference", "metric": "PairHMM GCUPS", "value": v, "unit": "GCUPS", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, len(times)), "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": WORKLOAD, "sample": sample,
                                     "note": "CPU ports on the box's host cores: the reference needs a JVM + the GKL jar and neither exists here (profiles/r02_jdk_probe.txt)"},
                          "cpu_baseline": cb, "scalar_fp64_port": scalar,
                          "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback exists)"
    torch.cuda.set_device(local_rank)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")  # host-side barriers: ranks that wait must not spin a kernel on their GPU
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
    e2e_wall_local = e2e_wall
    st2 = hmm.stats()

    # ---- e2e from pageable host arrays (the JNI pack copy: heap -> pinned, inside the library), rank 0 only ----
    pageable = None
    if rank == 0:
        from gatk_b200.native import Batch
        plain = Batch(*[np.array(getattr(batch, k)) for k in ("read_bases", "base_q", "ins_q", "del_q", "gcp", "read_off", "hap_bases", "hap_off", "units")])
        out_pg = np.full(batch.n_out, np.nan, dtype=np.float64)
        hmm.compute(plain, out_pg)
        t0 = time.perf_counter()
        for _ in range(max(1, min(args.steps, 3))):
            hmm.compute(plain, out_pg)
        torch.cuda.synchronize()
        pg_wall = (time.perf_counter() - t0) / max(1, min(args.steps, 3))
        pageable = {"value": cells / pg_wall / 1e9, "unit": "GCUPS", "ms_per_step": 1e3 * pg_wall, "bit_identical_with_pinned": bool(np.array_equal(out_pg, out)),
                    "api": "gphmm_compute (C ABI) with pageable host arrays: + one host copy of every input byte into pinned bounce buffers (what gpuphmm_jni.cpp pack() costs a JVM caller)"}
        del plain, out_pg

    # ---- 150 bp reads (configs[0] shape x 1 500 regions), rank 0 at N=1 only ----
    read150 = None
    if rank == 0 and world == 1 and not args.no_read150:
        read150 = measure_read150(hmm, synth, np)
    peak_measured = hmm.measure_fp32_peak(30.0) if rank == 0 else None

    (dev_s, wall, e2e_wall, f32_s_max), (cells_all, pairs_all, rescued_all) = sharding.reduce_timing(
        [dev_s, wall, e2e_wall, f32_s], [float(cells), float(pairs), float(st["rescued_pairs"])], device="cuda")

    # ---- N > 1: ONE handle over all devices on the shards of all ranks (rank 0 drives; the others wait on the CPU) ----
    one_handle = None
    if world > 1 and not args.no_one_handle:
        hmm.release_prepared(prepared)
        prepared = None
        own_e2e = cells * args.steps / e2e_wall_local / 1e9
        hmm.close()
        hmm = None
        torch.cuda.synchronize()
        res_all = sharding.collect_shards(batch, rank, world, tag="gphmm_bench", pinned=True, group=cpu_group, extra=out)
        if rank == 0:
            whole, outs = res_all
            one_handle = measure_one_handle(whole, outs, world, args.steps, own_e2e, np, torch)
            del whole, outs
        dist.barrier(group=cpu_group)

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
                         "traffic": 220070400, "kernel": "phmm_flat_f32_kernel<K> (+ phmm_fast_f32_kernel<K> for non-flat reads), all K buckets of the step",
                         "peak_source": "2 x %d SMs x 128 lanes x %.0f MHz (SM clock sampled during the timed region); MEASURED_PEAKS.json has no FP32 entry" % (props.multi_processor_count, sm_for_peak),
                         "flops_per_cell": 12, "kernel_ms_per_step": 1e3 * f32_s / steps,
                         "executed_frac": 12.0 * (cells - st["skipped_cells"] / steps) * steps / f32_s / 1e12 / peak_tflops,
                         "executed_frac_note": "frac counts every cell of the R x H matrices at the reference's 12 flop/cell (SURVEY 8d); executed_frac counts only the cells the kernels really sweep (prefix sharing skips the rest); a swept cell costs 5 FP instructions, so the FMA-pipe utilisation is executed_frac * 5/6",
                         "peak_measured": peak_measured, "frac_of_peak_measured": (achieved_tflops / peak_measured) if peak_measured else None,
                         "executed_frac_of_peak_measured": (12.0 * (cells - st["skipped_cells"] / steps) * steps / f32_s / 1e12 / peak_measured) if peak_measured else None,
                         "peak_measured_source": "gphmm_measure_fp32_peak: dense FFMA with constant-bank operands, 64 warps/SM, ~30 ms, same process and clocks",
                         "traffic_note": "bytes per launch of the dominant kernel phmm_flat_f32_kernel<16,0,16> (two 250-base reads per warp) from one ncu --set full capture: 20.8 MB DRAM read + 199.3 MB written in a 4.5 ms launch = 49 GB/s.  The reads are the inputs; the writes are write-back of the per-CTA snapshot slabs (scratch that is re-read from L2, 73 KB stored per task), not result traffic: the kernel is FP32-issue bound and HBM is 0.7 % busy; profiles/r02_flat16_k16_final_ncu_full.txt",
                         "hbm_gbs_staging": (batch.input_bytes() + 8 * pairs) * steps / f32_s / 1e9},
        }
        if pageable:
            res["e2e_from_pageable"] = pageable
        if read150:
            res["read_150bp"] = read150
        if one_handle:
            res["one_handle"] = one_handle
        if not args.no_cpu_baseline and world == 1:  # rank 0 at N=1 only
            small = synth.config2(min(args.regions, 400))
            res["cpu_baseline"], _, _ = cpu_baseline(small, budget_s=args.cpu_budget)
        print(json.dumps(res))
    if prepared is not None:
        hmm.release_prepared(prepared)
    if hmm is not None:
        hmm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
