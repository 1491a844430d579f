"""GPU test of solveRobustPGO through the C++ drop-in (reference: tests/testPGO.cpp:193-271).
Kept in its own file, last in collection order."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_robust_pgo_rejects_the_wrong_loop_closure():
    exe = os.path.join(ROOT, "dpgo_b200", "host", "bin", "robust_pgo_test")
    if not os.path.exists(exe):
        pytest.skip("host binaries not built (run __graft_entry__.build())")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
