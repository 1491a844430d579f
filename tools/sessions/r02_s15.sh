#!/bin/bash
# Round-2 GPU session 15 (1 GPU): device chordal initialization (first run), per-pose kernels staged through shared
# memory, fused solver without the folded phases (they did not pay: profiles/r02_fold_ab.jsonl).
O=gpurun_out/s15
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_a_parity.py tests/test_gpu_b_team.py -x -q -m gpu -s --durations=8 > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest_time.txt
tail -15 $O/pytest_gpu.log; grep -h "rotation_iterations" $O/pytest_gpu.log | head
timeout 300 python tools/pose_op_scale.py 64 100 > $O/pose_op_scale.jsonl 2> $O/pose_op_scale.err; cat $O/pose_op_scale.jsonl; tail -2 $O/pose_op_scale.err
timeout 200 python tools/dd_probe.py --barrier-ab > $O/solve_times.jsonl 2> $O/solve_times.err
python - <<'PY'
import json
for l in open("gpurun_out/s15/solve_times.jsonl"):
    d = json.loads(l); print(d["problem"], d["mode"], d["optimize_ms"], d["outer"], d["tcg"], d["barriers"], repr(d["two_f"]))
PY
