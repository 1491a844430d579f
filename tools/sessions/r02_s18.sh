#!/bin/bash
# Round-2 GPU session 18 (1 GPU): Q*X at scale with the grid capped at w resident waves (DPGO_QX_RESIDENT_WAVES) against
# one CTA per 64 poses; per-pose kernels after the cheaper Jacobi tests.
O=gpurun_out/s18
mkdir -p $O
for w in 0 1 2 4 0 1; do
  DPGO_QX_RESIDENT_WAVES=$w timeout 300 python tools/qx_scale.py 64 100 > $O/qx_w$w.tmp 2> $O/qx_w$w.err
  python - $w $O/qx_w$w.tmp >> $O/qx_waves.jsonl <<'PY'
import json, sys
for l in open(sys.argv[2]):
    d = json.loads(l); d["resident_waves"] = int(sys.argv[1]); print(json.dumps(d))
PY
  rm -f $O/qx_w$w.tmp
done
python - <<'PY'
import json
for l in open("gpurun_out/s18/qx_waves.jsonl"):
    d = json.loads(l)
    print(d["resident_waves"], d["n"], [(v["variant"], round(v["flushed_us"], 1), round(v["frac_of_measured_peak"], 3)) for v in d["variants"]])
PY
timeout 300 python tools/pose_op_scale.py 64 100 > $O/pose_op_scale.jsonl 2> $O/pose_op_scale.err
python - <<'PY'
import json
for l in open("gpurun_out/s18/pose_op_scale.jsonl"):
    d = json.loads(l); print(d["n"], [(o["op"], round(o["flushed_us"], 1), round(o["frac_of_measured_peak"], 3)) for o in d["ops"]])
PY
timeout 300 python -m pytest tests/test_gpu_a_parity.py -x -q -m gpu > $O/pytest_a.log 2>&1; tail -2 $O/pytest_a.log
