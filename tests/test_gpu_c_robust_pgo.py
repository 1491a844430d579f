"""GPU test of solveRobustPGO through the C++ drop-in (reference: tests/testPGO.cpp:193-271).
Kept in its own file, last in collection order."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_robust_pgo_rejects_the_wrong_loop_closure():
    from test_gpu_z_host import _need, _run
    out = _run([_need("robust_pgo_test")], 120)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
