#!/bin/bash
# Round-2 GPU session 13 (1 GPU): diagonal-block kernel with the shortened latency chain; strip tuning sweep.
O=gpurun_out/s13
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_a_parity.py tests/test_gpu_b_team.py -x -q -m gpu > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest_time.txt
tail -2 $O/pytest_gpu.log
timeout 300 python tools/gnc_update_probe.py > $O/gnc_probe.jsonl 2> $O/gnc_probe.err; cat $O/gnc_probe.jsonl; tail -2 $O/gnc_probe.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_gnc.csv \
  python tools/gnc_update_probe.py > $O/gnc_probe_ncu.log 2>&1
python tools/ncu_digest.py launches $O/launches_gnc.csv $O/launches_gnc_summary.json | tail -6
timeout 300 python tools/setup_time.py > $O/setup_time.jsonl 2> $O/setup_time.err; cat $O/setup_time.jsonl
timeout 400 python tools/dd_probe.py --strip-tuning > $O/strip_tuning.jsonl 2> $O/strip_tuning.err
python - <<'PY'
import json
for l in open("gpurun_out/s13/strip_tuning.jsonl"):
    d = json.loads(l); print(d["problem"], d["tuning"], d["optimize_ms"], d["apply_us"], [d["phase_ms"][i] for i in (8, 10, 12, 2)])
PY
