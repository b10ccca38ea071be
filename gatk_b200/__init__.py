"""gatk_b200 -- B200-native PairHMM forward path behind GATK's `-pairHMM CUDA_LOGLESS_CACHING`.

The product is gatk_b200/lib/libgpuphmm.so (CUDA, C ABI in include/gpuphmm.h) plus the Java plugin
sources under java/.  The Python modules here are thin bindings used by tests and bench.py; they
contain no likelihood arithmetic and no CPU fallback.
"""
from .native import GpuPhmm, GpuPhmmError, Batch, lib_path, load_library  # noqa: F401
