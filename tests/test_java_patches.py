"""The Java half of the drop-in (java/patches/*.patch) is mechanically true: every patch applies to the reference tree with
`git apply` (no fuzz, which `patch(1)` would allow), the series applies together (each reference file is touched by exactly
one patch), and the patched files still have balanced brackets.  No JDK exists in this image or on the GPU box
(profiles/r02_jdk_probe.txt), so this is as far as the check can go without a compiler.  Skipped when /root/reference is
absent (the GPU box)."""
import glob
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PATCHES = sorted(glob.glob(os.path.join(ROOT, "java", "patches", "*.patch")))

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "main", "java")), reason="reference tree not present")


def _targets(patch):
    return re.findall(r"^\+\+\+ b/(\S+)", open(patch).read(), flags=re.M)


def test_there_are_eight_patches_and_no_file_is_patched_twice():
    assert len(PATCHES) == 8
    seen = {}
    for p in PATCHES:
        for t in _targets(p):
            assert t not in seen, "%s is touched by %s and %s" % (t, seen[t], p)
            seen[t] = p
            assert os.path.isfile(os.path.join(REF, t)), t


@pytest.mark.parametrize("patch", PATCHES, ids=[os.path.basename(p) for p in PATCHES])
def test_patch_applies_to_the_reference_without_fuzz(patch):
    r = subprocess.run(["git", "apply", "--check", "--verbose", patch], cwd=REF, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def _strip_java(src):
    """comments, string and char literals removed (enough for bracket counting)"""
    out, i, n = [], 0, len(src)
    while i < n:
        c = src[i]
        if src.startswith("//", i):
            i = src.find("\n", i)
            i = n if i < 0 else i
        elif src.startswith("/*", i):
            i = src.find("*/", i + 2)
            i = n if i < 0 else i + 2
        elif c in "\"'":
            j = i + 1
            while j < n and src[j] != c:
                j += 2 if src[j] == "\\" else 1
            i = j + 1
        else:
            out.append(c)
            i += 1
    return "".join(out)


def _balanced(text):
    stack = []
    pairs = {")": "(", "]": "[", "}": "{"}
    for ch in text:
        if ch in "([{":
            stack.append(ch)
        elif ch in pairs:
            if not stack or stack.pop() != pairs[ch]:
                return False
    return not stack


def test_series_applies_together_and_brackets_stay_balanced(tmp_path):
    for p in PATCHES:
        for t in _targets(p):
            dst = tmp_path / t
            dst.parent.mkdir(parents=True, exist_ok=True)
            shutil.copy(os.path.join(REF, t), dst)
            os.chmod(dst, 0o644)
    for p in PATCHES:
        r = subprocess.run(["git", "apply", p], cwd=tmp_path, capture_output=True, text=True)
        assert r.returncode == 0, (p, r.stderr)
    for p in PATCHES:
        for t in _targets(p):
            before = _strip_java(open(os.path.join(REF, t)).read())
            after = _strip_java(open(tmp_path / t).read())
            assert _balanced(before), "bracket counter is wrong about the pristine " + t
            assert _balanced(after), t
    # the classes the patches call exist in java/, with the members they use
    cuda = open(os.path.join(ROOT, "java/org/broadinstitute/hellbender/utils/pairhmm/CudaLoglessPairHMM.java")).read()
    for member in ("class PendingLikelihoods", "void complete()", "PendingLikelihoods submitLog10Likelihoods(", "class RegionSteps", "int[] computeRegionLikelihoods("):
        assert member in cuda, member
    sw = open(os.path.join(ROOT, "java/org/broadinstitute/hellbender/utils/smithwaterman/CudaSmithWatermanAligner.java")).read()
    assert "List<SmithWatermanAlignment> alignBatch(" in sw
    assert os.path.isfile(os.path.join(ROOT, "java/org/broadinstitute/hellbender/utils/pairhmm/CudaLoglessPairPDHMM.java"))
    args = open(os.path.join(ROOT, "java/org/broadinstitute/hellbender/utils/pairhmm/CudaPairHMMArguments.java")).read()
    assert "extends PairHMMNativeArguments" in args and "public int[] devices" in args and "static int[] parseDeviceList(" in args
    assert "args instanceof CudaPairHMMArguments" in cuda


def test_our_java_sources_have_balanced_brackets():
    for f in glob.glob(os.path.join(ROOT, "java", "org", "**", "*.java"), recursive=True):
        assert _balanced(_strip_java(open(f).read())), f
