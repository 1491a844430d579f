"""GPU test of the three-phase form of the two-level preconditioner (precon_mode 3, opt-in).  Its
host side (index plan, layouts, algebra) is covered on the CPU by test_three_phase_plan_cpu.py; its
CUDA side was written after this round's GPU budget was spent and has not run on a device yet, so
outside DPGO_B200_EXPERIMENTAL=1 a failure is reported as an expected failure (the variant is not
used by any default path).  Runs in its own process, last in collection order."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STRICT = os.environ.get("DPGO_B200_EXPERIMENTAL") == "1"


@pytest.mark.xfail(condition=not STRICT, reason="precon_mode 3: first device run pending", strict=False)
def test_three_phase_preconditioner_matches_the_oracle():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "gpu_three_phase_check.py")],
                         capture_output=True, text=True, timeout=300)
    print(out.stdout[-4000:])
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
