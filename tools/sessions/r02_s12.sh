#!/bin/bash
# Round-2 GPU session 12 (1 GPU): faster diagonal-block kernel, symbolic part of the two-level set-up kept across
# weight-only rebuilds, stream-ordered scratch: suite, GNC update breakdown, set-up times, launch list.
O=gpurun_out/s12
mkdir -p $O
( time timeout 900 python -m pytest tests/ -x -q -m gpu --durations=5 > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest_time.txt
tail -3 $O/pytest_gpu.log
timeout 300 python tools/gnc_update_probe.py > $O/gnc_probe.jsonl 2> $O/gnc_probe.err; cat $O/gnc_probe.jsonl; tail -2 $O/gnc_probe.err
timeout 300 python tools/setup_time.py > $O/setup_time.jsonl 2> $O/setup_time.err; cat $O/setup_time.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_gnc.csv \
  python tools/gnc_update_probe.py > $O/gnc_probe_ncu.log 2>&1
python tools/ncu_digest.py launches $O/launches_gnc.csv $O/launches_gnc_summary.json | tail -12
timeout 300 python - > $O/anchors.jsonl 2> $O/anchors.err <<'PY'
import json, sys
sys.path.insert(0, '.')
from tools import bench_team
ds = dict(dataset='city10000', agents=4, r=3)
for gnc in (5, 0):
    t = bench_team.measure(20, 5, 0, 1, 0, schedule='all', mode='device', gnc=gnc, **ds)
    print(json.dumps(dict(ds, gnc=gnc, value=t['value'], ms_per_step=t['ms_per_step'], weight_updates=t['weight_updates'])), flush=True)
PY
cat $O/anchors.jsonl; tail -2 $O/anchors.err
