"""Builds gatk_b200/lib/libgpuphmm.so in-tree with nvcc for sm_100a (no GPU needed)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "lib", "libgpuphmm.so")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".inl", ".cpp", ".h"))]
    srcs.append(os.path.join(_HERE, "..", "include", "gpuphmm.h"))
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force=False, verbose=False):
    if force or _stale():
        cmd = ["make", "-C", CSRC] + (["-B"] if force else [])
        subprocess.run(cmd, check=True, stdout=None if verbose else subprocess.DEVNULL)
    return LIB


if __name__ == "__main__":
    print(build(verbose=True))
