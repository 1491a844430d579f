#!/bin/bash
# Round-2 GPU session 11 (1 GPU): suite with the rounding kernel and the device rounding in the C++ drop-in;
# time breakdown of a GNC weight update; launch list of the preconditioner set-up of a city10000 agent.
O=gpurun_out/s11
mkdir -p $O
( time timeout 900 python -m pytest tests/ -x -q -m gpu --durations=6 > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest_time.txt
tail -3 $O/pytest_gpu.log
timeout 300 python tools/gnc_update_probe.py > $O/gnc_probe.jsonl 2> $O/gnc_probe.err; cat $O/gnc_probe.jsonl; tail -2 $O/gnc_probe.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_gnc.csv \
  python tools/gnc_update_probe.py > $O/gnc_probe_ncu.log 2>&1
python tools/ncu_digest.py launches $O/launches_gnc.csv $O/launches_gnc_summary.json | tail -25
