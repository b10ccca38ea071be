"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/gpuphmm.h declares, and refuses to run without a GPU (no fallback).  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

from gatk_b200 import build as gbuild
from gatk_b200 import native, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    gbuild.build()
    return native.load_library()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gpuphmm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gphmm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(lib):
    names = _declared_symbols()
    assert set(names) == set(native.EXPORTS)
    for n in names:
        assert getattr(lib, n) is not None


def test_abi_version_and_strerror(lib):
    assert lib.gphmm_abi_version() == 1
    assert lib.gphmm_strerror(0) == b"ok"
    for code in range(-8, 0):
        assert len(lib.gphmm_strerror(code)) > 0


def test_struct_layouts_match_header():
    assert ctypes.sizeof(native._Unit) == 40 == native.UNIT_DTYPE.itemsize
    assert ctypes.sizeof(native._Batch) == 96
    assert ctypes.sizeof(native._RegionSteps) == 104
    assert ctypes.sizeof(native._SwParams) == 24 and ctypes.sizeof(native._SwBatch) == 40
    assert ctypes.sizeof(native._Config) == 48
    assert ctypes.sizeof(native.Stats) == 12 * 8


def test_no_gpu_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.gphmm_device_count() == 0
    with pytest.raises(native.GpuPhmmError) as e:
        native.GpuPhmm()
    assert e.value.code == native.ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    # the product path must never route through the CPU oracle
    pkg = os.path.join(ROOT, "gatk_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".java")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "phmm_oracle" not in text, f


def test_synthetic_shapes():
    b = synth.config1()
    assert b.n_reads == 128 and b.n_haps == 8 and b.pairs() == 1024
    assert np.all(np.diff(b.read_off) == 150)
    assert np.all((np.diff(b.hap_off) >= 190) & (np.diff(b.hap_off) <= 320))
    b2 = synth.config2(20)
    assert len(b2.units) == 20
    nh = b2.units["hap_end"] - b2.units["hap_begin"]
    assert nh.min() >= 4 and nh.max() <= 16
    assert np.diff(b2.read_off).max() == 250 and np.diff(b2.read_off).min() >= 100
    assert set(np.unique(b2.read_bases)) <= set(b"ACGT")
    assert b2.n_out == b2.pairs()


def test_batch_validation():
    b = synth.config1()
    with pytest.raises(ValueError):
        native.Batch(b.read_bases, b.base_q[:-1], b.ins_q, b.del_q, b.gcp, b.read_off, b.hap_bases, b.hap_off, b.units)


def test_jni_shim_syntax_and_symbols():
    # no JDK in this image: compile the shim against the minimal stub header (syntax + types only)
    import subprocess
    src = os.path.join(ROOT, "gatk_b200", "csrc", "gpuphmm_jni.cpp")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "tests", "jni_stub"), src], check=True)
    # every native method of the Java binding has its JNI export, and vice versa
    java = open(os.path.join(ROOT, "java/org/broadinstitute/hellbender/utils/pairhmm/CudaPairHMMBinding.java")).read()
    natives = set(re.findall(r"private static native [\w\[\]]+ (\w+)\(", java))
    exported = set(re.findall(r"JNIFN\((\w+)\)\(", open(src).read()))
    assert natives == exported and len(natives) == 11


def test_java_plugin_sources_present():
    hmm = open(os.path.join(ROOT, "java/org/broadinstitute/hellbender/utils/pairhmm/CudaLoglessPairHMM.java")).read()
    assert "extends LoglessPairHMM" in hmm and "HardwareFeatureException" in hmm
    for method in ("initialize(final List<Haplotype>", "computeLog10Likelihoods(final LikelihoodMatrix<GATKRead, Haplotype>", "public void close()",
                   "writeToResultsFileIfApplicable(", "mLogLikelihoodArray = flat"):
        assert method in hmm
    patch = open(os.path.join(ROOT, "java/patches/PairHMM.Implementation.patch")).read()
    assert "CUDA_LOGLESS_CACHING(args ->" in patch
    added = [l for l in patch.splitlines() if l.startswith("+") and not l.startswith("+++")]
    assert not any("FASTEST_AVAILABLE" in l for l in added)  # the default chain is untouched
    # caller side of the cross-region queue: the plugin's async form and the patch that uses it agree on names
    assert "public PendingLikelihoods submitLog10Likelihoods(" in hmm and "public void complete()" in hmm
    pipelined = open(os.path.join(ROOT, "java/patches/HaplotypeCaller.pipelinedCallRegion.patch")).read()
    engine = open(os.path.join(ROOT, "java/patches/PairHMMLikelihoodCalculationEngine.regionSteps.patch")).read()
    for name in ("cuda.submitLog10Likelihoods(", "PendingLikelihoods::complete", "submitReadLikelihoods("):
        assert name in engine
    for name in (".submitReadLikelihoods(", "callRegionDeferred(", "onTraversalSuccess"):
        assert name in pipelined
    sw = open(os.path.join(ROOT, "java/org/broadinstitute/hellbender/utils/smithwaterman/CudaSmithWatermanAligner.java")).read()
    realign = open(os.path.join(ROOT, "java/patches/AssemblyBasedCallerUtils.batchedRealign.patch")).read()
    assert "public List<SmithWatermanAlignment> alignBatch(" in sw and "gpu.alignBatch(haplotypes, reads, parameters, SWOverhangStrategy.SOFTCLIP)" in realign


def test_plugin_mirror_without_gpu():
    import torch
    from gatk_b200 import pairhmm
    assert [e.name for e in pairhmm.Implementation] == ["EXACT", "ORIGINAL", "LOGLESS_CACHING", "AVX_LOGLESS_CACHING",
                                                         "AVX_LOGLESS_CACHING_OMP", "CUDA_LOGLESS_CACHING", "FASTEST_AVAILABLE"]
    if not torch.cuda.is_available():
        with pytest.raises(pairhmm.HardwareFeatureException):  # VectorLoglessPairHMM.java:63-66 convention: throw, no fallback
            pairhmm.Implementation.CUDA_LOGLESS_CACHING.makeNewHMM(pairhmm.PairHMMNativeArguments())
    imp = pairhmm.StandardPairHMMInputScoreImputator(10)
    ins, dele, gcp = imp.impute(pairhmm.Read(b"ACGT", [30, 30, 30, 30]))
    assert list(ins) == [45] * 4 and list(dele) == [45] * 4 and list(gcp) == [10] * 4


def test_planner_host_only():
    # gphmm_plan_stats runs the real chunk planner without a GPU: schedule totals are consistent
    b = synth.config2(40)
    plain = native.plan_stats(b, False)
    shared = native.plan_stats(b, True)
    nh = int(np.sum(b.units["hap_end"] - b.units["hap_begin"]))
    assert plain["units"] == shared["units"] == 40 and plain["total_columns"] == int(b.hap_off[-1])
    assert plain["skipped_columns"] == 0 and plain["snapshots"] == 0
    # a small chunk is split into haplotype groups: one pass and one schedule per haplotype, no sharing lost or gained
    assert plain["passes"] == nh
    big = synth.config2(400)
    sp, ss = native.plan_stats(big, False), native.plan_stats(big, True)
    # enough reads: one task per read, or per PAIR of reads where two reads of a unit share a warp (half-warp kernels:
    # reads of 128-254 bases whose last rows fall on the same register slot); 90 % of these reads are 250 bases long
    assert sp["tasks"] == ss["tasks"] and big.n_reads // 2 <= ss["tasks"] <= int(0.62 * big.n_reads)
    steps = lambda st: st["free_steps"] + st["checked_steps"]
    # every computed column is one step of lane 0; each unit's sweep drains 31 more steps
    assert steps(sp) == sp["total_columns"] + sp["passes"] + 31 * sp["units"]
    assert 0.15 < ss["skipped_columns"] / ss["total_columns"] < 0.6
    # the near-depth rule of the snapshot planner: close to what the trie of the sorted haplotypes allows
    trie = 0
    for u in big.units:
        haps = sorted(bytes(big.hap_bases[big.hap_off[h]:big.hap_off[h + 1]]) for h in range(int(u["hap_begin"]), int(u["hap_end"])))
        prev = b""
        for h in haps:
            m = 0
            while m < min(len(h), len(prev)) and h[m] == prev[m]:
                m += 1
            trie += m
            prev = h
    assert ss["skipped_columns"] <= trie and ss["skipped_columns"] >= 0.93 * trie
    assert steps(ss) < 0.9 * steps(sp)
    assert ss["checked_steps"] <= 32 * (ss["passes"] + ss["snapshots"]) + ss["units"]
    # equal-length reads of up to 159 bases: four per warp (quarter-warp tasks) once the chunk is large enough
    quad = synth.config1_many(48)   # 48 regions x 128 reads x 150 bases
    assert native.plan_stats(quad)["tasks"] == quad.n_reads // 4
    few = synth.config1_many(8)     # too few reads to fill the GPU that way: one read per warp, haplotypes split into groups
    assert native.plan_stats(few)["tasks"] >= few.n_reads
