"""Host-side mirror of the reference's PairHMM plugin surface for the CUDA path, used by tests and examples.

The real plugin is Java (java/org/broadinstitute/hellbender/utils/pairhmm/CudaLoglessPairHMM.java); this image
has no JVM, so this module restates the same surface -- names, argument meaning, error behaviour -- over the
same C ABI, without any arithmetic of its own:

  PairHMM.Implementation.makeNewHMM(args)                    PairHMM.java:40-109
  PairHMMNativeArguments{maxNumberOfThreads,useDoublePrecision}   PairHMMNativeArgumentCollection.java:18-23
  initialize(haplotypes, perSampleReadList, readMaxLength, haplotypeMaxLength)   VectorLoglessPairHMM.java:89-102
  computeLog10Likelihoods(logLikelihoods, processedReads, inputScoreImputator)   VectorLoglessPairHMM.java:108-159
  getLogLikelihoodArray() / close()                          PairHMM.java:357-359,391-402
  computeRegionLikelihoods(matrix, clippedReads, constantGCP, steps)   the fused region steps (CudaLoglessPairHMM.java;
      replaces PairHMMLikelihoodCalculationEngine.java:194-203,283-316 + AlleleLikelihoods.java:416-458,1351-1376)
"""
import enum
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from .native import Batch, GpuPhmm, GpuPhmmError, ERR_NO_DEVICE


class HardwareFeatureException(RuntimeError):
    """UserException.HardwareFeatureException (exceptions/UserException.java:435-445)."""


@dataclass
class PairHMMNativeArguments:
    maxNumberOfThreads: int = 4          # --native-pair-hmm-threads
    useDoublePrecision: bool = False     # --native-pair-hmm-use-double-precision


@dataclass
class Read:
    """The five per-base arrays the binding reads from a GATKRead + its imputed scores (VectorLoglessPairHMM.java:120-129)."""
    bases: bytes
    base_quals: Sequence[int]
    ins_quals: Optional[Sequence[int]] = None   # BI tag or the flat default
    del_quals: Optional[Sequence[int]] = None   # BD tag or the flat default
    mapping_quality: int = 60                   # GATKRead.getMappingQuality (region steps only)
    hmm_base_qualities: Optional[np.ndarray] = None   # HMM_BASE_QUALITIES_TAG, set by computeRegionLikelihoods


@dataclass
class RegionSteps:
    """CudaLoglessPairHMM.RegionSteps: the PairHMMLikelihoodCalculationEngine fields behind the fused steps."""
    pcrRateFactor: float = 3.0                  # PCRErrorModel CONSERVATIVE; 0 = NONE
    baseQualityScoreThreshold: int = 18
    disableCapReadQualitiesToMapQ: bool = False
    log10GlobalReadMismappingRate: float = -4.5
    symmetricallyNormalizeAllelesToReference: bool = False
    filterPoorly: bool = True
    expectedErrorRatePerBase: float = 0.02
    dynamicDisqualification: bool = False
    readDisqualificationScale: float = 1.0


class StandardPairHMMInputScoreImputator:
    """StandardPairHMMInputScoreImputator.java:27-47: BI/BD tags if present else flat Q45 (ReadUtils.java:44,838-862),
    flat gap-continuation penalty (--pair-hmm-gap-continuation-penalty, default 10)."""
    DEFAULT_INSERTION_DELETION_QUAL = 45

    def __init__(self, constant_gcp=10):
        self.constant_gcp = constant_gcp

    def impute(self, read: Read):
        n = len(read.bases)
        flat = np.full(n, self.DEFAULT_INSERTION_DELETION_QUAL, dtype=np.uint8)
        ins = flat if read.ins_quals is None else np.asarray(read.ins_quals, dtype=np.uint8)
        dele = flat if read.del_quals is None else np.asarray(read.del_quals, dtype=np.uint8)
        return ins, dele, np.full(n, self.constant_gcp, dtype=np.uint8)


class LikelihoodMatrix:
    """Minimal LikelihoodMatrix: values[allele][read] (AlleleLikelihoods.java:71-74,1418-1421 is allele-major)."""

    def __init__(self, alleles: List[bytes], n_reads: int):
        self._alleles = list(alleles)
        self.values = np.full((len(alleles), n_reads), np.nan)

    def alleles(self):
        return self._alleles

    def numberOfAlleles(self):
        return len(self._alleles)

    def set(self, allele_index, read_index, value):
        self.values[allele_index, read_index] = value


class CudaLoglessPairHMM:
    """Mirror of CudaLoglessPairHMM.java / VectorLoglessPairHMM.java over libgpuphmm."""

    def __init__(self, args: Optional[PairHMMNativeArguments] = None, devices=None):
        args = args or PairHMMNativeArguments()
        try:
            self._hmm = GpuPhmm(devices=devices, force_fp64=args.useDoublePrecision, host_threads=args.maxNumberOfThreads)
        except GpuPhmmError as e:
            if e.code == ERR_NO_DEVICE:
                raise HardwareFeatureException("Machine does not support the CUDA PairHMM.") from e
            raise
        self._haps: List[bytes] = []
        self._hap_index: Dict[bytes, int] = {}
        self.mLogLikelihoodArray = None
        self._initialized = False

    def initialize(self, haplotypes: List[bytes], perSampleReadList=None, readMaxLength=0, haplotypeMaxLength=0):
        self._haps = [bytes(h) for h in haplotypes]
        self._hap_index = {h: i for i, h in enumerate(self._haps)}
        self._initialized = True

    def computeLog10Likelihoods(self, logLikelihoods: LikelihoodMatrix, processedReads: List[Read], inputScoreImputator):
        if not processedReads:
            return  # VectorLoglessPairHMM.java:110-112
        if not self._initialized:
            raise RuntimeError("Must call initialize before calling computeLog10Likelihoods")  # PairHMM.java:284
        rows = []
        for r in processedReads:
            ins, dele, gcp = inputScoreImputator.impute(r)
            q = np.asarray(r.base_quals, dtype=np.uint8)
            if not (len(q) == len(ins) == len(dele) == len(gcp) == len(r.bases)):
                raise ValueError("Read bases and read quals aren't the same size")  # PairHMM.java:286-292
            rows.append((r.bases, q, ins, dele, gcp))
        batch = Batch.single_unit(rows, self._haps)
        self.mLogLikelihoodArray = self._hmm.compute(batch)
        n_haps = len(self._haps)
        for r in range(len(processedReads)):
            for hap_idx, hap in enumerate(logLikelihoods.alleles()):
                logLikelihoods.set(hap_idx, r, self.mLogLikelihoodArray[r * n_haps + self._hap_index[bytes(hap)]])

    def computeRegionLikelihoods(self, logLikelihoods: LikelihoodMatrix, clippedReads: List[Read], constantGCP: int,
                                 steps: RegionSteps, referenceHaplotype: Optional[bytes] = None):
        """modifyReadQualities -> PairHMM -> normalizeLikelihoods -> filterPoorlyModeledEvidence in one GPU call.
        Fills the matrix with the normalised likelihoods of all reads and returns the indexes of the reads to remove."""
        if not clippedReads:
            return []
        if not self._initialized:
            raise RuntimeError("Must call initialize before calling computeRegionLikelihoods")
        imputator = StandardPairHMMInputScoreImputator(constantGCP)
        rows = []
        for r in clippedReads:
            ins, dele, gcp = imputator.impute(r)
            rows.append((r.bases, np.asarray(r.base_quals, dtype=np.uint8), ins, dele, gcp))
        batch = Batch.single_unit(rows, self._haps)
        ref = -1 if referenceHaplotype is None else self._hap_index[bytes(referenceHaplotype)]
        res = self._hmm.compute_regions(
            batch, [min(255, r.mapping_quality) for r in clippedReads], [ref], pcr_rate_factor=steps.pcrRateFactor,
            base_quality_score_threshold=steps.baseQualityScoreThreshold, disable_cap_to_mapq=steps.disableCapReadQualitiesToMapQ,
            log10_global_read_mismapping_rate=steps.log10GlobalReadMismappingRate,
            symmetric=steps.symmetricallyNormalizeAllelesToReference, filter_poorly=steps.filterPoorly,
            expected_error_rate_per_base=steps.expectedErrorRatePerBase, dynamic_disqualification=steps.dynamicDisqualification,
            read_disqualification_scale=steps.readDisqualificationScale)
        n_reads = len(clippedReads)
        for hap_idx, hap in enumerate(logLikelihoods.alleles()):
            column = self._hap_index[bytes(hap)] * n_reads
            logLikelihoods.values[hap_idx, :] = res["lk"][column:column + n_reads]
        for k, r in enumerate(clippedReads):
            r.hmm_base_qualities = res["base_q"][batch.read_off[k]:batch.read_off[k + 1]].copy()
        return [int(k) for k in np.nonzero(res["keep"] == 0)[0]]

    def getLogLikelihoodArray(self):
        return self.mLogLikelihoodArray

    def close(self):
        self._hmm.close()


@dataclass
class PartiallyDeterminedHaplotype:
    """bases + PartiallyDeterminedHaplotype.getAlternateBases() (utils/haplotype/PartiallyDeterminedHaplotype.java:59-65:
    SNP=1 DEL_START=2 DEL_END=4 A=8 C=16 G=32 T=64, one flag byte per base)"""
    bases: bytes
    alternate_bases: Sequence[int]


class CudaLoglessPairPDHMM:
    """Mirror of CudaLoglessPairPDHMM.java / VectorLoglessPairPDHMM.java:71-147 over gphmm_pd_compute.  The read-span filter
    (rangeForReadOverlapToDeterminedBases) needs read and allele coordinates and stays with the caller, as in Java."""

    def __init__(self, devices=None):
        try:
            self._hmm = GpuPhmm(devices=devices)
        except GpuPhmmError as e:
            if e.code == ERR_NO_DEVICE:
                raise HardwareFeatureException("Machine does not support the CUDA PDHMM.") from e
            raise
        self.mLogLikelihoodArray = None

    def computeLog10Likelihoods(self, logLikelihoods: LikelihoodMatrix, processedReads: List[Read], inputScoreImputator,
                                alleles: List[PartiallyDeterminedHaplotype]):
        if not processedReads:
            return  # VectorLoglessPairPDHMM.java:77-79
        rows = []
        for r in processedReads:
            ins, dele, gcp = inputScoreImputator.impute(r)
            rows.append((r.bases, np.asarray(r.base_quals, dtype=np.uint8), ins, dele, gcp))
        batch = Batch.single_unit(rows, [a.bases for a in alleles])
        pd = np.concatenate([np.asarray(a.alternate_bases, dtype=np.uint8) for a in alleles])
        self.mLogLikelihoodArray = self._hmm.pd_compute(batch, pd)
        n = len(alleles)
        for r in range(len(processedReads)):
            for a in range(n):
                logLikelihoods.set(a, r, self.mLogLikelihoodArray[r * n + a])

    def getLogLikelihoodArray(self):
        return self.mLogLikelihoodArray

    def close(self):
        self._hmm.close()


class Implementation(enum.Enum):
    """PairHMM.Implementation (PairHMM.java:40-109).  Only the CUDA entry is constructible here; the Java/AVX entries
    belong to the reference and are listed so that the registry reads the same.  FASTEST_AVAILABLE does not include
    CUDA_LOGLESS_CACHING."""
    EXACT = "EXACT"
    ORIGINAL = "ORIGINAL"
    LOGLESS_CACHING = "LOGLESS_CACHING"
    AVX_LOGLESS_CACHING = "AVX_LOGLESS_CACHING"
    AVX_LOGLESS_CACHING_OMP = "AVX_LOGLESS_CACHING_OMP"
    CUDA_LOGLESS_CACHING = "CUDA_LOGLESS_CACHING"
    FASTEST_AVAILABLE = "FASTEST_AVAILABLE"

    def makeNewHMM(self, args: Optional[PairHMMNativeArguments] = None):
        if self is Implementation.CUDA_LOGLESS_CACHING:
            return CudaLoglessPairHMM(args)
        raise NotImplementedError("%s is implemented by the reference (Java / GKL), not by this package" % self.name)
