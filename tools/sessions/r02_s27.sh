#!/bin/bash
# Round-2 GPU session 27 (1 GPU): strip waves issued by the lanes of a warp (one stage per lane) and the interior wave of
# the next tCG iteration put in flight by the CTA's last warp at the start of the finish phase: parity, solve times.
O=gpurun_out/s27
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_a_parity.py tests/test_gpu_b_team.py -x -q -m gpu > $O/pytest.log 2>&1; tail -2 $O/pytest.log
timeout 200 python tools/dd_probe.py --barrier-ab > $O/solve_times.jsonl 2> $O/solve_times.err
python - <<'PY'
import json
for l in open("gpurun_out/s27/solve_times.jsonl"):
    d = json.loads(l); print(d["problem"], d["mode"], d["optimize_ms"], d["apply_us"], d["outer"], d["tcg"], d["barriers"], repr(d["two_f"]), [d["phase_ms"][i] for i in (8, 9, 10, 11, 12, 2, 5, 3, 4)])
PY
for name in torus3D city10000; do python - $name <<'PY'
import sys, json
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import dd_probe
z, d, n, T0 = dd_probe.fixture(sys.argv[1])
dd_probe.run(sys.argv[1], z, d, n, T0, 5 if d == 3 else 3, 2, reps=5)
PY
done
