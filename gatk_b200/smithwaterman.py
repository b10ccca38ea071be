"""Host-side mirror of the reference's Smith-Waterman plugin surface for the CUDA aligner (the real class is Java:
java/org/broadinstitute/hellbender/utils/smithwaterman/CudaSmithWatermanAligner.java).  No arithmetic here.

  SmithWatermanAligner.align(ref, alt, SWParameters, SWOverhangStrategy)      SmithWatermanAligner.java:20-33
  SmithWatermanAligner.Implementation / getAligner(type)                        SmithWatermanAligner.java:45-87
  SWParameters presets                                                          SmithWatermanAlignmentConstants.java:30-68
"""
import enum
from dataclasses import dataclass
from typing import List, Sequence

from .native import GpuPhmm, GpuPhmmError, ERR_NO_DEVICE
from .pairhmm import HardwareFeatureException


@dataclass(frozen=True)
class SWParameters:
    matchValue: int
    mismatchPenalty: int
    gapOpenPenalty: int
    gapExtendPenalty: int

    def as_tuple(self):
        return (self.matchValue, self.mismatchPenalty, self.gapOpenPenalty, self.gapExtendPenalty)


ORIGINAL_DEFAULT = SWParameters(3, -1, -4, -3)
STANDARD_NGS = SWParameters(25, -50, -110, -6)
NEW_SW_PARAMETERS = SWParameters(200, -150, -260, -11)
ALIGNMENT_TO_BEST_HAPLOTYPE_SW_PARAMETERS = SWParameters(10, -15, -30, -5)


class SWOverhangStrategy(enum.IntEnum):
    SOFTCLIP = 0
    INDEL = 1
    LEADING_INDEL = 2
    IGNORE = 3


@dataclass(frozen=True)
class SmithWatermanAlignment:
    cigar: str
    alignmentOffset: int

    def getCigar(self):
        return self.cigar

    def getAlignmentOffset(self):
        return self.alignmentOffset


class CudaSmithWatermanAligner:
    def __init__(self, devices=None):
        try:
            self._hmm = GpuPhmm(devices=devices)
        except GpuPhmmError as e:
            if e.code == ERR_NO_DEVICE:
                raise HardwareFeatureException("Machine does not support the CUDA Smith-Waterman.") from e
            raise

    def align(self, reference: bytes, alternate: bytes, parameters: SWParameters, overhangStrategy: SWOverhangStrategy):
        return self.alignBatch([reference], [alternate], parameters, overhangStrategy)[0]

    def alignBatch(self, references: Sequence[bytes], alternates: Sequence[bytes], parameters: SWParameters,
                   overhangStrategy: SWOverhangStrategy) -> List[SmithWatermanAlignment]:
        if parameters is None or overhangStrategy is None:
            raise ValueError("Null object is not allowed here.")       # Utils.nonNull
        if any(r is None or a is None or len(r) == 0 or len(a) == 0 for r, a in zip(references, alternates)):
            raise ValueError("Non-null, non-empty sequences are required for the Smith-Waterman calculation")  # SmithWatermanJavaAligner.java:64-66
        capacity = 32
        while True:
            try:
                res = self._hmm.sw_align(references, alternates, parameters.as_tuple(), int(overhangStrategy), cigar_capacity=capacity)
                return [SmithWatermanAlignment(c, o) for o, c in res]
            except GpuPhmmError as e:
                if e.code != -8:
                    raise
                capacity *= 8   # a CIGAR with more elements than expected: redo with more room

    def close(self):
        self._hmm.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class Implementation(enum.Enum):
    """SmithWatermanAligner.Implementation; only CUDA is constructible here."""
    FASTEST_AVAILABLE = "FASTEST_AVAILABLE"
    AVX_ENABLED = "AVX_ENABLED"
    JAVA = "JAVA"
    CUDA = "CUDA"


def getAligner(type_: Implementation):
    if type_ is Implementation.CUDA:
        return CudaSmithWatermanAligner()
    raise NotImplementedError("%s belongs to the reference (Java / GKL); this package only provides CUDA" % type_.name)
