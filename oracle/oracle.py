"""ctypes front-end of the CPU test oracle (oracle/pairhmm_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (gatk_b200) never imports this module.

Parity status: PINNED against the reference's fixtures by tests/test_oracle_golden.py.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libphmm_oracle.so")
_lib = None

_u8p = ctypes.POINTER(ctypes.c_uint8)
_i32p = ctypes.POINTER(ctypes.c_int32)
_f64p = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc, -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in ("pairhmm_oracle.c", "pairhmm_simd_baseline.c", "region_steps_oracle.c", "sw_oracle.c", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(x) for x in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libphmm_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.phmm_oracle_init.restype = None
        L.phmm_oracle_qual_to_error_prob.restype = ctypes.c_double
        L.phmm_oracle_qual_to_error_prob.argtypes = [ctypes.c_int]
        L.phmm_oracle_match_to_match_prob.restype = ctypes.c_double
        L.phmm_oracle_match_to_match_prob.argtypes = [ctypes.c_int, ctypes.c_int]
        L.phmm_oracle_qual_to_trans_probs.restype = ctypes.c_int
        L.phmm_oracle_qual_to_trans_probs.argtypes = [_f64p, ctypes.c_uint8, ctypes.c_uint8, ctypes.c_uint8]
        pair_args = [_u8p, ctypes.c_int, _u8p, _u8p, _u8p, _u8p, _u8p, ctypes.c_int]
        L.phmm_oracle_logless.restype = ctypes.c_int
        L.phmm_oracle_logless.argtypes = pair_args + [ctypes.c_int, _f64p]
        L.phmm_oracle_log10.restype = ctypes.c_int
        L.phmm_oracle_log10.argtypes = pair_args + [ctypes.c_int, ctypes.c_int, _f64p]
        L.phmm_oracle_pd_logless.restype = ctypes.c_int
        L.phmm_oracle_pd_logless.argtypes = [_u8p, _u8p, ctypes.c_int, _u8p, _u8p, _u8p, _u8p, _u8p, ctypes.c_int, ctypes.c_int, _f64p]
        L.phmm_oracle_unit.restype = ctypes.c_int
        L.phmm_oracle_unit.argtypes = [_u8p, _u8p, _u8p, _u8p, _u8p, _i32p, ctypes.c_int,
                                       _u8p, _i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f64p]
        L.phmm_oracle_max_threads.restype = ctypes.c_int
        L.phmm_simd_isa.restype = ctypes.c_int
        L.phmm_simd_batch.restype = ctypes.c_long
        L.phmm_simd_batch.argtypes = [_u8p, _u8p, _u8p, _u8p, _u8p, ctypes.POINTER(ctypes.c_int64), _u8p, ctypes.POINTER(ctypes.c_int64),
                                      ctypes.POINTER(ctypes.c_int64), ctypes.c_long, ctypes.c_int, _f64p]
        L.region_oracle_find_repetitions.restype = ctypes.c_int
        L.region_oracle_find_repetitions.argtypes = [_u8p, ctypes.c_int, ctypes.c_int, _u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.region_oracle_tandem_repeat_length.restype = ctypes.c_int
        L.region_oracle_tandem_repeat_length.argtypes = [_u8p, ctypes.c_int, ctypes.c_int]
        L.region_oracle_pcr_cache.restype = None
        L.region_oracle_pcr_cache.argtypes = [ctypes.c_double, _u8p]
        L.region_oracle_modify_read.restype = None
        L.region_oracle_modify_read.argtypes = [_u8p, _u8p, _u8p, _u8p, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int]
        L.region_oracle_normalize.restype = None
        L.region_oracle_normalize.argtypes = [_f64p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, _f64p]
        L.region_oracle_min_true_likelihood.restype = ctypes.c_double
        L.region_oracle_min_true_likelihood.argtypes = [_u8p, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_double]
        L.region_oracle_filter.restype = None
        L.region_oracle_filter.argtypes = [_f64p, ctypes.c_int, ctypes.c_int, _u8p, ctypes.POINTER(ctypes.c_int64), ctypes.c_double,
                                           ctypes.c_int, ctypes.c_double, _u8p]
        L.sw_oracle_align.restype = ctypes.c_int
        L.sw_oracle_align.argtypes = [_u8p, ctypes.c_int, _u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int, ctypes.POINTER(ctypes.c_uint32), ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        L.phmm_oracle_init()
        _lib = L
    return _lib


def _u8(a):
    a = np.ascontiguousarray(np.frombuffer(a, dtype=np.uint8) if isinstance(a, (bytes, bytearray)) else a, dtype=np.uint8)
    return a, a.ctypes.data_as(_u8p)


def _pair(fn, hap, read, base_q, ins_q, del_q, gcp, *extra):
    hap_a, hap_p = _u8(hap)
    rd_a, rd_p = _u8(read)
    arrs = [_u8(x) for x in (base_q, ins_q, del_q, gcp)]
    for a, _ in arrs:
        if len(a) != len(rd_a):
            # PairHMM.java:286-292 (IllegalArgumentException)
            raise ValueError("read bases and quals aren't the same size")
    out = ctypes.c_double(float("nan"))
    rc = fn(hap_p, len(hap_a), rd_p, arrs[0][1], arrs[1][1], arrs[2][1], arrs[3][1], len(rd_a), *extra, ctypes.byref(out))
    if rc != 0:
        raise ValueError("oracle error %d" % rc)
    return out.value


def logless(hap, read, base_q, ins_q, del_q, gcp, tristate_off=False):
    """LoglessPairHMM (Java double) log10 likelihood of one (read, haplotype) pair."""
    return _pair(lib().phmm_oracle_logless, hap, read, base_q, ins_q, del_q, gcp, int(tristate_off))


def log10hmm(hap, read, base_q, ins_q, del_q, gcp, exact=True, tristate_off=False):
    """Log10PairHMM: EXACT (exact=True) or ORIGINAL (exact=False)."""
    return _pair(lib().phmm_oracle_log10, hap, read, base_q, ins_q, del_q, gcp, int(exact), int(tristate_off))


def pd_logless(hap, pd, read, base_q, ins_q, del_q, gcp, tristate_off=False):
    """LoglessPDPairHMM for one (read, partially determined haplotype) pair; pd = PartiallyDeterminedHaplotype.getAlternateBases()"""
    (h, hp), (f, fp), (r, rp), (q, qp), (i, ip), (d, dp), (g, gp) = (_u8(x) for x in (hap, pd, read, base_q, ins_q, del_q, gcp))
    assert len(f) == len(h) and len(q) == len(i) == len(d) == len(g) == len(r)
    out = ctypes.c_double(0.0)
    rc = lib().phmm_oracle_pd_logless(hp, fp, len(h), rp, qp, ip, dp, gp, len(r), 1 if tristate_off else 0, ctypes.byref(out))
    if rc != 0:
        raise ValueError("oracle error %d" % rc)
    return out.value


def qual_to_error_prob(q):
    return lib().phmm_oracle_qual_to_error_prob(int(q))


def match_to_match_prob(i, d):
    return lib().phmm_oracle_match_to_match_prob(int(i), int(d))


def qual_to_trans_probs(ins_q, del_q, gcp):
    dest = (ctypes.c_double * 6)()
    rc = lib().phmm_oracle_qual_to_trans_probs(dest, ins_q, del_q, gcp)
    if rc != 0:
        raise ValueError("oracle error %d" % rc)
    return list(dest)


def max_threads():
    return lib().phmm_oracle_max_threads()


def unit(read_bases, base_q, ins_q, del_q, gcp, read_off, hap_bases, hap_off, tristate_off=False, threads=1):
    """One (region, sample) unit, flat SoA in, read-major double[nReads*nHaps] out."""
    rb, rbp = _u8(read_bases)
    bq, bqp = _u8(base_q)
    iq, iqp = _u8(ins_q)
    dq, dqp = _u8(del_q)
    gc, gcp_p = _u8(gcp)
    hb, hbp = _u8(hap_bases)
    ro = np.ascontiguousarray(read_off, dtype=np.int32)
    ho = np.ascontiguousarray(hap_off, dtype=np.int32)
    n_reads, n_haps = len(ro) - 1, len(ho) - 1
    out = np.full(n_reads * n_haps, np.nan, dtype=np.float64)
    rc = lib().phmm_oracle_unit(rbp, bqp, iqp, dqp, gcp_p, ro.ctypes.data_as(_i32p), n_reads,
                                hbp, ho.ctypes.data_as(_i32p), n_haps, int(tristate_off), int(threads),
                                out.ctypes.data_as(_f64p))
    if rc != 0:
        raise ValueError("oracle error %d" % rc)
    return out


def simd_isa():
    """512 (AVX-512), 256 (AVX2) or 0 (generic): the vector ISA the SIMD baseline uses on this host."""
    return lib().phmm_simd_isa()


def simd_batch(read_bases, base_q, ins_q, del_q, gcp, read_off, hap_bases, hap_off, units, n_out, threads=1):
    """Vectorised fp32 CPU baseline (pairhmm_simd_baseline.c) over a whole batch; units = int64[n,5] rows of
    (read_begin, read_end, hap_begin, hap_end, out_off).  Returns (log10 likelihoods, pairs redone in double)."""
    arrs = [_u8(x) for x in (read_bases, base_q, ins_q, del_q, gcp)]
    hb, hbp = _u8(hap_bases)
    ro = np.ascontiguousarray(read_off, dtype=np.int64)
    ho = np.ascontiguousarray(hap_off, dtype=np.int64)
    un = np.ascontiguousarray(units, dtype=np.int64).reshape(-1, 5)
    out = np.full(n_out, np.nan, dtype=np.float64)
    i64p = ctypes.POINTER(ctypes.c_int64)
    rescued = lib().phmm_simd_batch(arrs[0][1], arrs[1][1], arrs[2][1], arrs[3][1], arrs[4][1], ro.ctypes.data_as(i64p), hbp,
                                    ho.ctypes.data_as(i64p), un.ctypes.data_as(i64p), len(un), int(threads), out.ctypes.data_as(_f64p))
    return out, int(rescued)


# ---- steps either side of the kernel (oracle/region_steps_oracle.c) -------------------------------------------------
def _bytes(x):
    return np.ascontiguousarray(np.frombuffer(x, dtype=np.uint8) if isinstance(x, (bytes, bytearray)) else x, dtype=np.uint8)


def find_repetitions(unit, test, leading, unit_off=0, unit_len=None, test_off=0, test_len=None):
    """GATKVariantContextUtils.findNumberOfRepetitions (both overloads)"""
    u, t = _bytes(unit), _bytes(test)
    if unit_len is None:
        unit_len = len(u)
    if test_len is None:
        test_len = len(t)
    if len(t) == 0:
        t = np.zeros(1, np.uint8)
    return lib().region_oracle_find_repetitions(u.ctypes.data_as(_u8p), unit_off, unit_len, t.ctypes.data_as(_u8p), test_off, test_len,
                                                1 if leading else 0)


def tandem_repeat_length(bases, offset):
    b = _bytes(bases)
    return lib().region_oracle_tandem_repeat_length(b.ctypes.data_as(_u8p), len(b), offset)


def pcr_cache(rate_factor):
    out = np.zeros(21, np.uint8)
    lib().region_oracle_pcr_cache(float(rate_factor), out.ctypes.data_as(_u8p))
    return out


def modify_reads(read_bases, base_q, ins_q, del_q, read_off, mapq, rate_factor=3.0, bq_threshold=18, disable_cap_to_mapq=False):
    """modifyReadQualities on every read of flat arrays; returns new (base_q, ins_q, del_q)"""
    L = lib()
    rb = _bytes(read_bases)
    q, i, d = (np.array(x, dtype=np.uint8, copy=True) for x in (base_q, ins_q, del_q))
    for r in range(len(read_off) - 1):
        o, n = int(read_off[r]), int(read_off[r + 1] - read_off[r])
        if n == 0:
            continue
        L.region_oracle_modify_read(rb[o:].ctypes.data_as(_u8p), q[o:].ctypes.data_as(_u8p), i[o:].ctypes.data_as(_u8p),
                                    d[o:].ctypes.data_as(_u8p), n, int(mapq[r]), float(rate_factor), int(bq_threshold),
                                    1 if disable_cap_to_mapq else 0)
    return q, i, d


def normalize(lk, n_reads, n_haps, ref_hap=-1, max_diff_cap=-4.5, symmetric=False):
    """read-major lk -> normalised allele-major matrix (flat)"""
    a = np.ascontiguousarray(lk, dtype=np.float64)
    out = np.zeros(n_reads * n_haps, np.float64)
    lib().region_oracle_normalize(a.ctypes.data_as(_f64p), n_reads, n_haps, ref_hap, float(max_diff_cap), 1 if symmetric else 0,
                                  out.ctypes.data_as(_f64p))
    return out


def min_true_likelihood(hmm_base_q, max_error_per_base=0.02, dynamic=False, dynamic_scale=1.0):
    q = _bytes(hmm_base_q)
    p = q if len(q) else np.zeros(1, np.uint8)
    return lib().region_oracle_min_true_likelihood(p.ctypes.data_as(_u8p), len(q), float(max_error_per_base), 1 if dynamic else 0,
                                                   float(dynamic_scale))


def filter_poorly_modeled(lk_allele_major, n_reads, n_haps, hmm_base_q, read_off, max_error_per_base=0.02, dynamic=False,
                          dynamic_scale=1.0):
    a = np.ascontiguousarray(lk_allele_major, dtype=np.float64)
    q = _bytes(hmm_base_q)
    ro = np.ascontiguousarray(read_off, dtype=np.int64)
    keep = np.zeros(max(n_reads, 1), np.uint8)
    lib().region_oracle_filter(a.ctypes.data_as(_f64p), n_reads, n_haps, q.ctypes.data_as(_u8p), ro.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                               float(max_error_per_base), 1 if dynamic else 0, float(dynamic_scale), keep.ctypes.data_as(_u8p))
    return keep[:n_reads]


# ---- Smith-Waterman (oracle/sw_oracle.c) ---------------------------------------------------------------------------------
SW_SOFTCLIP, SW_INDEL, SW_LEADING_INDEL, SW_IGNORE = 0, 1, 2, 3
SW_OPS = "MIDS"


def cigar_string(elems):
    return "".join("%d%s" % (int(e) >> 4, SW_OPS[int(e) & 15]) for e in elems)


def sw_align(ref, alt, params, strategy):
    """SmithWatermanJavaAligner.align: params = (match, mismatch, gap open, gap extend) -> (offset, CIGAR string)"""
    r, a = _bytes(ref), _bytes(alt)
    cap = len(r) + len(a) + 4
    elems = np.zeros(cap, np.uint32)
    off = ctypes.c_int(0)
    n = lib().sw_oracle_align(r.ctypes.data_as(_u8p), len(r), a.ctypes.data_as(_u8p), len(a), *[int(x) for x in params], int(strategy),
                              elems.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), cap, ctypes.byref(off))
    if n < 0:
        raise ValueError("sw oracle error %d" % n)
    return off.value, cigar_string(elems[:n])
