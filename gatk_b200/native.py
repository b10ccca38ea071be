"""ctypes binding of libgpuphmm.so (include/gpuphmm.h).  No arithmetic happens in Python."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_u8p = ctypes.POINTER(ctypes.c_uint8)
_i64p = ctypes.POINTER(ctypes.c_int64)
_i32p = ctypes.POINTER(ctypes.c_int32)
_f64p = ctypes.POINTER(ctypes.c_double)

ERR_INVALID_ARG, ERR_NO_DEVICE, ERR_CUDA, ERR_BAD_QUAL, ERR_ALPHABET, ERR_NOMEM, ERR_BAD_TICKET, ERR_TOO_LARGE = range(-1, -9, -1)


class GpuPhmmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("gpuphmm error %d: %s" % (code, msg))
        self.code = code


class _Config(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_int32), ("n_devices", ctypes.c_int32), ("devices", _i32p),
                ("force_fp64", ctypes.c_int32), ("host_threads", ctypes.c_int32), ("tristate_off", ctypes.c_int32),
                ("no_prefix_sharing", ctypes.c_int32), ("chunk_cells", ctypes.c_int64), ("chunk_bytes", ctypes.c_int64)]


class _Unit(ctypes.Structure):
    _fields_ = [("read_begin", ctypes.c_int64), ("read_end", ctypes.c_int64), ("hap_begin", ctypes.c_int64),
                ("hap_end", ctypes.c_int64), ("out_off", ctypes.c_int64)]


class _Batch(ctypes.Structure):
    _fields_ = [("read_bases", ctypes.c_void_p), ("base_q", ctypes.c_void_p), ("ins_q", ctypes.c_void_p),
                ("del_q", ctypes.c_void_p), ("gcp", ctypes.c_void_p), ("read_off", ctypes.c_void_p),
                ("n_reads", ctypes.c_int64), ("hap_bases", ctypes.c_void_p), ("hap_off", ctypes.c_void_p),
                ("n_haps", ctypes.c_int64), ("units", ctypes.c_void_p), ("n_units", ctypes.c_int64)]


class Stats(ctypes.Structure):
    _fields_ = [("pairs", ctypes.c_int64), ("cells", ctypes.c_int64), ("rescued_pairs", ctypes.c_int64),
                ("skipped_cells", ctypes.c_int64), ("h2d_bytes", ctypes.c_int64), ("d2h_bytes", ctypes.c_int64),
                ("kernel_launches", ctypes.c_int64), ("fp32_kernel_ms", ctypes.c_double), ("fp64_kernel_ms", ctypes.c_double),
                ("device_ms", ctypes.c_double), ("host_stage_ms", ctypes.c_double), ("wall_ms", ctypes.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class _RegionSteps(ctypes.Structure):
    """gphmm_region_steps (include/gpuphmm.h)"""
    _fields_ = [("struct_size", ctypes.c_int32), ("flags", ctypes.c_int32), ("pcr_rate_factor", ctypes.c_double),
                ("base_quality_score_threshold", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("log10_global_read_mismapping_rate", ctypes.c_double), ("expected_error_rate_per_base", ctypes.c_double),
                ("read_disqualification_scale", ctypes.c_double), ("mapq", ctypes.c_void_p), ("ref_hap", ctypes.c_void_p),
                ("keep", ctypes.c_void_p), ("hmm_base_q", ctypes.c_void_p), ("hmm_ins_q", ctypes.c_void_p),
                ("hmm_del_q", ctypes.c_void_p), ("raw_lk", ctypes.c_void_p)]


class _SwParams(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_int32), ("match_value", ctypes.c_int32), ("mismatch_penalty", ctypes.c_int32),
                ("gap_open_penalty", ctypes.c_int32), ("gap_extend_penalty", ctypes.c_int32), ("overhang_strategy", ctypes.c_int32)]


class _SwBatch(ctypes.Structure):
    _fields_ = [("ref_bases", ctypes.c_void_p), ("ref_off", ctypes.c_void_p), ("alt_bases", ctypes.c_void_p),
                ("alt_off", ctypes.c_void_p), ("n_pairs", ctypes.c_int64)]


SW_SOFTCLIP, SW_INDEL, SW_LEADING_INDEL, SW_IGNORE = 0, 1, 2, 3
ERR_TOO_LARGE = -8

RS_DISABLE_CAP_TO_MAPQ, RS_SYMMETRIC_NORMALIZE, RS_FILTER_POORLY, RS_DYNAMIC_DISQ = 1, 2, 4, 8

UNIT_DTYPE = np.dtype([("read_begin", "<i8"), ("read_end", "<i8"), ("hap_begin", "<i8"), ("hap_end", "<i8"), ("out_off", "<i8")])

EXPORTS = ["gphmm_abi_version", "gphmm_device_count", "gphmm_strerror", "gphmm_create", "gphmm_destroy",
           "gphmm_last_error", "gphmm_compute", "gphmm_compute_regions", "gphmm_submit_regions", "gphmm_pd_compute", "gphmm_sw_align", "gphmm_submit", "gphmm_wait", "gphmm_prepare", "gphmm_run_prepared",
           "gphmm_release_prepared", "gphmm_get_stats", "gphmm_reset_stats", "gphmm_plan_stats", "gphmm_measure_fp32_peak", "gphmm_host_alloc", "gphmm_host_free"]


def lib_path():
    # GPHMM_LIB: another build of the same library (A/B measurements of two versions on one box)
    return os.environ.get("GPHMM_LIB") or os.path.join(_HERE, "lib", "libgpuphmm.so")


def load_library():
    """Loads the CUDA library.  Fails loudly when it has not been built -- there is no fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        # not built yet: compile it in-tree (nvcc, sm_100a).  There is still no fallback -- if this fails we raise.
        try:
            from . import build as _build
            _build.build()
        except Exception as e:
            raise GpuPhmmError(ERR_NO_DEVICE, "%s is missing and could not be built (%s): run `python -m gatk_b200.build`" % (path, e))
    L = ctypes.CDLL(path)
    L.gphmm_abi_version.restype = ctypes.c_int
    L.gphmm_device_count.restype = ctypes.c_int
    L.gphmm_strerror.restype = ctypes.c_char_p
    L.gphmm_strerror.argtypes = [ctypes.c_int]
    L.gphmm_create.restype = ctypes.c_int
    L.gphmm_create.argtypes = [ctypes.POINTER(_Config), ctypes.POINTER(ctypes.c_void_p)]
    L.gphmm_destroy.restype = None
    L.gphmm_destroy.argtypes = [ctypes.c_void_p]
    L.gphmm_last_error.restype = ctypes.c_char_p
    L.gphmm_last_error.argtypes = [ctypes.c_void_p]
    L.gphmm_compute.restype = ctypes.c_int
    L.gphmm_compute.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Batch), ctypes.c_void_p]
    L.gphmm_compute_regions.restype = ctypes.c_int
    L.gphmm_compute_regions.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Batch), ctypes.POINTER(_RegionSteps), ctypes.c_void_p]
    L.gphmm_pd_compute.restype = ctypes.c_int
    L.gphmm_pd_compute.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Batch), ctypes.c_void_p, ctypes.c_void_p]
    L.gphmm_sw_align.restype = ctypes.c_int
    L.gphmm_sw_align.argtypes = [ctypes.c_void_p, ctypes.POINTER(_SwBatch), ctypes.POINTER(_SwParams), ctypes.c_int32, ctypes.c_void_p,
                                 ctypes.c_void_p, ctypes.c_void_p]
    L.gphmm_submit_regions.restype = ctypes.c_int
    L.gphmm_submit_regions.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Batch), ctypes.POINTER(_RegionSteps), ctypes.c_void_p,
                                       ctypes.POINTER(ctypes.c_uint64)]
    L.gphmm_submit.restype = ctypes.c_int
    L.gphmm_submit.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Batch), ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
    L.gphmm_wait.restype = ctypes.c_int
    L.gphmm_wait.argtypes = [ctypes.c_void_p, ctypes.c_uint64]
    L.gphmm_prepare.restype = ctypes.c_int
    L.gphmm_prepare.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Batch), ctypes.POINTER(ctypes.c_void_p)]
    L.gphmm_run_prepared.restype = ctypes.c_int
    L.gphmm_run_prepared.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.gphmm_release_prepared.restype = None
    L.gphmm_release_prepared.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.gphmm_get_stats.restype = ctypes.c_int
    L.gphmm_get_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(Stats)]
    L.gphmm_reset_stats.restype = None
    L.gphmm_reset_stats.argtypes = [ctypes.c_void_p]
    L.gphmm_measure_fp32_peak.restype = ctypes.c_int
    L.gphmm_measure_fp32_peak.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.POINTER(ctypes.c_double)]
    L.gphmm_plan_stats.restype = ctypes.c_int
    L.gphmm_plan_stats.argtypes = [ctypes.POINTER(_Batch), ctypes.c_int, ctypes.POINTER(ctypes.c_int64)]
    L.gphmm_host_alloc.restype = ctypes.c_void_p
    L.gphmm_host_alloc.argtypes = [ctypes.c_size_t]
    L.gphmm_host_free.restype = None
    L.gphmm_host_free.argtypes = [ctypes.c_void_p]
    _LIB = L
    return L


class Batch:
    """Flat SoA batch (gphmm_batch).  Arrays are numpy; `pinned=True` places the big per-base arrays in
    CUDA pinned host memory obtained from the library so that staging is a straight DMA."""

    def __init__(self, read_bases, base_q, ins_q, del_q, gcp, read_off, hap_bases, hap_off, units, pinned=False):
        self._pinned_ptrs = []
        conv = (lambda a: self._pin(np.ascontiguousarray(a, dtype=np.uint8))) if pinned else (lambda a: np.ascontiguousarray(a, dtype=np.uint8))
        self.read_bases, self.base_q, self.ins_q, self.del_q, self.gcp = (conv(a) for a in (read_bases, base_q, ins_q, del_q, gcp))
        self.hap_bases = np.ascontiguousarray(hap_bases, dtype=np.uint8)
        self.read_off = np.ascontiguousarray(read_off, dtype=np.int64)
        self.hap_off = np.ascontiguousarray(hap_off, dtype=np.int64)
        u = np.asarray(units)
        if u.dtype != UNIT_DTYPE:
            uu = np.zeros(len(u), dtype=UNIT_DTYPE)
            u = np.asarray(u, dtype=np.int64).reshape(-1, 5)
            for k, name in enumerate(UNIT_DTYPE.names):
                uu[name] = u[:, k]
            u = uu
        self.units = np.ascontiguousarray(u)
        n = len(self.read_bases)
        for a in (self.base_q, self.ins_q, self.del_q, self.gcp):
            if len(a) != n:
                raise ValueError("per-base read arrays differ in length")  # PairHMM.java:286-292
        if len(self.read_off) < 1 or self.read_off[-1] != n:
            raise ValueError("read_off does not cover the read arrays")
        if len(self.hap_off) < 1 or self.hap_off[-1] != len(self.hap_bases):
            raise ValueError("hap_off does not cover hap_bases")

    def _pin(self, a):
        L = load_library()
        nbytes = max(a.nbytes, 1)
        ptr = L.gphmm_host_alloc(nbytes)
        if not ptr:
            raise GpuPhmmError(ERR_NOMEM, "pinned allocation failed")
        self._pinned_ptrs.append(ptr)
        buf = (ctypes.c_uint8 * nbytes).from_address(ptr)
        out = np.frombuffer(buf, dtype=np.uint8, count=a.size)
        out[:] = a
        return out

    def __del__(self):
        try:
            L = load_library()
            for p in self._pinned_ptrs:
                L.gphmm_host_free(p)
        except Exception:
            pass
        self._pinned_ptrs = []

    @property
    def n_reads(self):
        return len(self.read_off) - 1

    @property
    def n_haps(self):
        return len(self.hap_off) - 1

    @property
    def n_out(self):
        if len(self.units) == 0:
            return 0
        nr = self.units["read_end"] - self.units["read_begin"]
        nh = self.units["hap_end"] - self.units["hap_begin"]
        return int(np.max(self.units["out_off"] + nr * nh))

    def cells(self):
        rl = np.diff(self.read_off)
        hl = np.diff(self.hap_off)
        rc = np.concatenate([[0], np.cumsum(rl)])
        hc = np.concatenate([[0], np.cumsum(hl)])
        u = self.units
        return int(np.sum((rc[u["read_end"]] - rc[u["read_begin"]]) * (hc[u["hap_end"]] - hc[u["hap_begin"]])))

    def pairs(self):
        u = self.units
        return int(np.sum((u["read_end"] - u["read_begin"]) * (u["hap_end"] - u["hap_begin"])))

    def input_bytes(self):
        return 5 * int(self.read_off[-1]) + int(self.hap_off[-1])

    def c_struct(self):
        b = _Batch()
        b.read_bases, b.base_q, b.ins_q, b.del_q, b.gcp = (a.ctypes.data for a in (self.read_bases, self.base_q, self.ins_q, self.del_q, self.gcp))
        b.read_off = self.read_off.ctypes.data
        b.n_reads = self.n_reads
        b.hap_bases = self.hap_bases.ctypes.data
        b.hap_off = self.hap_off.ctypes.data
        b.n_haps = self.n_haps
        b.units = self.units.ctypes.data
        b.n_units = len(self.units)
        return b

    @staticmethod
    def concat(batches, pinned=False):
        """The batches back to back as one batch (units, reads, haplotypes and output slots renumbered): what one handle
        over several GPUs is given when every shard of a sharded job is collected in one process."""
        cat = lambda xs, dt: np.concatenate(xs) if len(xs) else np.zeros(0, dt)
        read_off, hap_off, units = [np.zeros(1, np.int64)], [np.zeros(1, np.int64)], []
        r0 = h0 = b0 = c0 = o0 = 0
        for b in batches:
            read_off.append(b.read_off[1:] + b0)
            hap_off.append(b.hap_off[1:] + c0)
            u = b.units.copy()
            u["read_begin"] += r0; u["read_end"] += r0; u["hap_begin"] += h0; u["hap_end"] += h0; u["out_off"] += o0
            units.append(u)
            r0 += b.n_reads; h0 += b.n_haps; b0 += int(b.read_off[-1]); c0 += int(b.hap_off[-1]); o0 += b.n_out
        cols = [cat([getattr(b, name) for b in batches], np.uint8) for name in ("read_bases", "base_q", "ins_q", "del_q", "gcp", "hap_bases")]
        return Batch(cols[0], cols[1], cols[2], cols[3], cols[4], cat(read_off, np.int64), cols[5], cat(hap_off, np.int64),
                     cat(units, UNIT_DTYPE), pinned=pinned)

    @staticmethod
    def single_unit(reads, haps):
        """reads: list of (bases, base_q, ins_q, del_q, gcp); haps: list of bytes."""
        as_u8 = lambda x: np.frombuffer(x, dtype=np.uint8) if isinstance(x, (bytes, bytearray)) else np.asarray(x, dtype=np.uint8)
        cols = [[as_u8(r[k]) for r in reads] for k in range(5)]
        cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.uint8)
        read_off = np.concatenate([[0], np.cumsum([len(x) for x in cols[0]])]).astype(np.int64)
        hs = [as_u8(h) for h in haps]
        hap_off = np.concatenate([[0], np.cumsum([len(h) for h in hs])]).astype(np.int64)
        units = np.array([(0, len(reads), 0, len(haps), 0)], dtype=UNIT_DTYPE)
        return Batch(cat(cols[0]), cat(cols[1]), cat(cols[2]), cat(cols[3]), cat(cols[4]), read_off, cat(hs), hap_off, units)

    @staticmethod
    def from_units(regions):
        """regions: list of (reads, haps) as in single_unit; one unit per entry, outputs back to back."""
        as_u8 = lambda x: np.frombuffer(x, dtype=np.uint8) if isinstance(x, (bytes, bytearray)) else np.asarray(x, dtype=np.uint8)
        cols = [[], [], [], [], []]
        hs, units = [], []
        n_reads = n_haps = n_out = 0
        for reads, haps in regions:
            for r in reads:
                for k in range(5):
                    cols[k].append(as_u8(r[k]))
            hs.extend(as_u8(h) for h in haps)
            units.append((n_reads, n_reads + len(reads), n_haps, n_haps + len(haps), n_out))
            n_reads += len(reads)
            n_haps += len(haps)
            n_out += len(reads) * len(haps)
        cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.uint8)
        read_off = np.concatenate([[0], np.cumsum([len(x) for x in cols[0]])]).astype(np.int64)
        hap_off = np.concatenate([[0], np.cumsum([len(h) for h in hs])]).astype(np.int64)
        return Batch(cat(cols[0]), cat(cols[1]), cat(cols[2]), cat(cols[3]), cat(cols[4]), read_off, cat(hs), hap_off,
                     np.array(units, dtype=UNIT_DTYPE))


def plan_stats(batch, prefix_sharing=True):
    """Host-only planning totals (gphmm_plan_stats); works without a GPU."""
    L = load_library()
    out = (ctypes.c_int64 * 10)()
    b = batch.c_struct()
    rc = L.gphmm_plan_stats(ctypes.byref(b), int(prefix_sharing), out)
    if rc != 0:
        raise GpuPhmmError(rc, L.gphmm_strerror(rc).decode())
    keys = ("units", "passes", "segments", "free_steps", "checked_steps", "snapshots", "skipped_columns", "total_columns", "chunks", "tasks")
    return dict(zip(keys, [int(x) for x in out]))


class GpuPhmm:
    """RAII wrapper of a gphmm_t handle."""

    def __init__(self, devices=None, force_fp64=False, host_threads=0, tristate_off=False, chunk_cells=0, chunk_bytes=0,
                 no_prefix_sharing=False):
        self._L = load_library()
        self._h = ctypes.c_void_p()
        cfg = _Config()
        cfg.struct_size = ctypes.sizeof(_Config)
        self._dev_arr = None
        if devices:
            self._dev_arr = (ctypes.c_int32 * len(devices))(*devices)
            cfg.n_devices = len(devices)
            cfg.devices = ctypes.cast(self._dev_arr, _i32p)
        cfg.force_fp64 = int(force_fp64)
        cfg.host_threads = int(host_threads)
        cfg.tristate_off = int(tristate_off)
        cfg.no_prefix_sharing = int(no_prefix_sharing)
        cfg.chunk_cells = int(chunk_cells)
        cfg.chunk_bytes = int(chunk_bytes)
        rc = self._L.gphmm_create(ctypes.byref(cfg), ctypes.byref(self._h))
        if rc != 0:
            raise GpuPhmmError(rc, self._L.gphmm_strerror(rc).decode())
        self._pending = {}

    def _check(self, rc):
        if rc != 0:
            raise GpuPhmmError(rc, self._L.gphmm_last_error(self._h).decode() or self._L.gphmm_strerror(rc).decode())

    def close(self):
        if self._h:
            self._L.gphmm_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def compute(self, batch, out=None):
        if out is None:
            out = np.full(batch.n_out, np.nan, dtype=np.float64)
        b = batch.c_struct()
        self._check(self._L.gphmm_compute(self._h, ctypes.byref(b), out.ctypes.data))
        return out

    @staticmethod
    def _region_steps(batch, mapq, ref_hap, pcr_rate_factor=3.0, base_quality_score_threshold=18, disable_cap_to_mapq=False,
                      log10_global_read_mismapping_rate=-4.5, symmetric=False, filter_poorly=True,
                      expected_error_rate_per_base=0.02, dynamic_disqualification=False, read_disqualification_scale=1.0,
                      want_quals=True, want_raw=True):
        """-> (gphmm_region_steps, result dict, arrays to keep alive)"""
        n_reads = len(batch.read_off) - 1
        mapq = np.ascontiguousarray(mapq, dtype=np.uint8)
        if len(mapq) != n_reads:
            raise ValueError("mapq needs one entry per read")
        out = np.full(batch.n_out, np.nan, dtype=np.float64)
        keep = np.full(max(n_reads, 1), 255, dtype=np.uint8)
        rs = _RegionSteps()
        rs.struct_size = ctypes.sizeof(_RegionSteps)
        rs.flags = ((RS_DISABLE_CAP_TO_MAPQ if disable_cap_to_mapq else 0) | (RS_SYMMETRIC_NORMALIZE if symmetric else 0)
                    | (RS_FILTER_POORLY if filter_poorly else 0) | (RS_DYNAMIC_DISQ if dynamic_disqualification else 0))
        rs.pcr_rate_factor = float(pcr_rate_factor)
        rs.base_quality_score_threshold = int(base_quality_score_threshold)
        rs.log10_global_read_mismapping_rate = float(log10_global_read_mismapping_rate)
        rs.expected_error_rate_per_base = float(expected_error_rate_per_base)
        rs.read_disqualification_scale = float(read_disqualification_scale)
        rs.mapq = mapq.ctypes.data if n_reads else None
        if ref_hap is not None:
            ref_hap = np.ascontiguousarray(ref_hap, dtype=np.int32)
            if len(ref_hap) != len(batch.units):
                raise ValueError("ref_hap needs one entry per unit")
            rs.ref_hap = ref_hap.ctypes.data
        rs.keep = keep.ctypes.data
        res = {"lk": out, "keep": keep[:n_reads]}
        if want_raw:
            res["raw"] = np.full(batch.n_out, np.nan, dtype=np.float64)
            rs.raw_lk = res["raw"].ctypes.data
        if want_quals:
            n = len(batch.read_bases)
            for name, field in (("base_q", "hmm_base_q"), ("ins_q", "hmm_ins_q"), ("del_q", "hmm_del_q")):
                a = np.zeros(max(n, 1), dtype=np.uint8)
                setattr(rs, field, a.ctypes.data)
                res[name] = a[:n]
        return rs, res, (mapq, ref_hap, keep)

    def compute_regions(self, batch, mapq, ref_hap=None, **params):
        """gphmm_compute_regions: modifyReadQualities -> PairHMM -> normalizeLikelihoods -> filterPoorlyModeledEvidence
        on the device.  Returns a dict: lk (flat, per unit allele-major [h*nReads + r]), keep (per read),
        base_q / ins_q / del_q (the qualities the kernel used, when want_quals), raw (un-normalised read-major
        likelihoods, when want_raw).  params: see _region_steps."""
        rs, res, _alive = self._region_steps(batch, mapq, ref_hap, **params)
        b = batch.c_struct()
        self._check(self._L.gphmm_compute_regions(self._h, ctypes.byref(b), ctypes.byref(rs), res["lk"].ctypes.data))
        return res

    def pd_compute(self, batch, hap_pd_bases, out=None):
        """gphmm_pd_compute: LoglessPDPairHMM; hap_pd_bases is parallel to batch.hap_bases"""
        pd = np.ascontiguousarray(hap_pd_bases, dtype=np.uint8)
        if len(pd) != len(batch.hap_bases):
            raise ValueError("hap_pd_bases needs one byte per haplotype base")
        if out is None:
            out = np.full(batch.n_out, np.nan, dtype=np.float64)
        b = batch.c_struct()
        self._check(self._L.gphmm_pd_compute(self._h, ctypes.byref(b), pd.ctypes.data if len(pd) else None, out.ctypes.data))
        return out

    def sw_align(self, refs, alts, params, strategy, cigar_capacity=64):
        """gphmm_sw_align: refs / alts are lists of bytes (pair k aligns alts[k] to refs[k]); params = (match, mismatch, gap open,
        gap extend).  Returns a list of (offset, CIGAR string)."""
        n = len(refs)
        assert len(alts) == n
        cat = lambda xs: np.frombuffer(b"".join(bytes(x) for x in xs), dtype=np.uint8) if n else np.zeros(0, np.uint8)
        rb, ab = cat(refs), cat(alts)
        ro = np.concatenate([[0], np.cumsum([len(x) for x in refs])]).astype(np.int64)
        ao = np.concatenate([[0], np.cumsum([len(x) for x in alts])]).astype(np.int64)
        b = _SwBatch(rb.ctypes.data if len(rb) else None, ro.ctypes.data, ab.ctypes.data if len(ab) else None, ao.ctypes.data, n)
        p = _SwParams(ctypes.sizeof(_SwParams), int(params[0]), int(params[1]), int(params[2]), int(params[3]), int(strategy))
        offsets = np.zeros(max(n, 1), np.int32)
        n_elems = np.zeros(max(n, 1), np.int32)
        elems = np.zeros(max(n, 1) * cigar_capacity, np.uint32)
        self._check(self._L.gphmm_sw_align(self._h, ctypes.byref(b), ctypes.byref(p), cigar_capacity, offsets.ctypes.data, n_elems.ctypes.data,
                                           elems.ctypes.data))
        ops = "MIDS"
        return [(int(offsets[k]), "".join("%d%s" % (int(e) >> 4, ops[int(e) & 15]) for e in elems[k * cigar_capacity:k * cigar_capacity + n_elems[k]]))
                for k in range(n)]

    def submit_regions(self, batch, mapq, ref_hap=None, **params):
        """gphmm_submit_regions; wait(ticket) returns the result dict of compute_regions"""
        rs, res, alive = self._region_steps(batch, mapq, ref_hap, **params)
        b = batch.c_struct()
        t = ctypes.c_uint64(0)
        self._check(self._L.gphmm_submit_regions(self._h, ctypes.byref(b), ctypes.byref(rs), res["lk"].ctypes.data, ctypes.byref(t)))
        res["_alive"] = alive
        self._pending[t.value] = res
        return t.value

    def submit(self, batch, out=None):
        if out is None:
            out = np.full(batch.n_out, np.nan, dtype=np.float64)
        b = batch.c_struct()
        t = ctypes.c_uint64(0)
        self._check(self._L.gphmm_submit(self._h, ctypes.byref(b), out.ctypes.data, ctypes.byref(t)))
        self._pending[t.value] = out
        return t.value

    def wait(self, ticket):
        rc = self._L.gphmm_wait(self._h, ctypes.c_uint64(ticket))
        out = self._pending.pop(ticket, None)
        self._check(rc)
        return out

    def prepare(self, batch):
        b = batch.c_struct()
        p = ctypes.c_void_p()
        self._check(self._L.gphmm_prepare(self._h, ctypes.byref(b), ctypes.byref(p)))
        return p

    def run_prepared(self, prepared, out=None):
        self._check(self._L.gphmm_run_prepared(self._h, prepared, out.ctypes.data if out is not None else None))
        return out

    def release_prepared(self, prepared):
        self._L.gphmm_release_prepared(self._h, prepared)

    def stats(self):
        s = Stats()
        self._check(self._L.gphmm_get_stats(self._h, ctypes.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        self._L.gphmm_reset_stats(self._h)

    def measure_fp32_peak(self, millis=20.0):
        """Sustained FP32 FFMA rate of the handle's first device in TFLOP/s (gphmm_measure_fp32_peak; measurement aid)."""
        v = ctypes.c_double(0.0)
        self._check(self._L.gphmm_measure_fp32_peak(self._h, ctypes.c_double(millis), ctypes.byref(v)))
        return v.value
