"""The step after the likelihood matrix: diploid genotype likelihoods / PLs from fp32 results equal those from the
double-precision oracle (tools/genotype_stability.py; the measurable proxy for the north star's VCF identity)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "tools", "genotype_stability.py")


def test_approx_log10_sum_matches_the_reference_known_answers():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import genotype_stability as gs
    # as MathUtilsUnitTest.testApproximateLogSumLog (MathUtilsUnitTest.java:160-168): within 1e-4 of the exact sum
    for a, b in ((0.0, 0.0), (-1.0, 0.0), (0.0, -1.0), (-2.2, -3.5), (-1.0, -7.1), (5.0, 6.2), (38.1, 16.2), (-38.1, 6.2), (-19.1, -37.1)):
        exact = np.log10(10.0 ** a + 10.0 ** b)
        assert abs(float(gs.approx_log10_sum(np.float64(a), np.float64(b))) - exact) < 1e-4
    assert float(gs.approx_log10_sum(np.float64(-np.inf), np.float64(-3.0))) == -3.0
    assert float(gs.approx_log10_sum(np.float64(0.0), np.float64(-9.0))) == 0.0   # beyond MAX_TOLERANCE the smaller term is dropped


def test_pls_from_the_model_of_the_kernel_arithmetic_equal_the_oracle_pls():
    out = subprocess.run([sys.executable, TOOL, "--backend", "model", "--regions", "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "best diploid genotype differs in 0 of 2 regions; 0 of" in out.stdout


@pytest.mark.gpu
def test_pls_from_the_gpu_equal_the_oracle_pls():
    out = subprocess.run([sys.executable, TOOL, "--backend", "gpu", "--regions", "100"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "best diploid genotype differs in 0 of 100 regions" in out.stdout and "keep/drop differs for 0 of" in out.stdout
