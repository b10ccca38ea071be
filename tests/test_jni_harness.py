"""The JNI shim (gatk_b200/csrc/gpuphmm_jni.cpp) executed without a JVM: tests/jni_stub/jni_harness.cpp implements the JNIEnv
of the stub header over a toy object model and calls the shim's native* entry points as CudaPairHMMBinding.java does."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tests", "jni_stub")
BIN = os.path.join(STUB, "jni_harness")
LIBDIR = os.path.join(ROOT, "gatk_b200", "lib")


def _build():
    srcs = [os.path.join(STUB, "jni_harness.cpp"), os.path.join(ROOT, "gatk_b200", "csrc", "gpuphmm_jni.cpp")]
    deps = srcs + [os.path.join(STUB, "jni.h"), os.path.join(ROOT, "include", "gpuphmm.h"), os.path.join(LIBDIR, "libgpuphmm.so")]
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + STUB, "-o", BIN] + srcs +
                              ["-L" + LIBDIR, "-lgpuphmm", "-Wl,-rpath," + LIBDIR])
    return BIN


def test_shim_without_a_gpu_raises_hardware_feature_exception():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible: the gpu mode of the harness covers this box")
    out = subprocess.run([_build(), "cpu"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "jni_harness cpu: ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_shim_agrees_with_the_c_abi_on_the_gpu():
    out = subprocess.run([_build(), "gpu"], capture_output=True, text=True, timeout=90)
    assert out.returncode == 0 and "jni_harness gpu: ok" in out.stdout, out.stdout + out.stderr
