#!/bin/bash
# Round-2 GPU session 9 (1 GPU): permuted residual copy in the fused solver, team bench path incl. the asynchronous
# series on one device, ncu launch list and full capture of the fused kernel, Q*X automatic choice.
O=gpurun_out/s9
mkdir -p $O
( time timeout 900 python -m pytest tests/ -x -q -m gpu --durations=6 > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest_time.txt
tail -3 $O/pytest_gpu.log
( time timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err ) 2> $O/bench_n1_time.txt; echo "bench n1 rc=$?"; cat $O/bench_n1_time.txt; tail -c 400 $O/bench_n1.err
timeout 600 python bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_team_1gpu.json 2> $O/bench_team_1gpu.err; echo "team rc=$?"; tail -c 600 $O/bench_team_1gpu.err
timeout 300 python tools/qx_scale.py 64 100 > $O/qx_scale.jsonl 2> $O/qx_scale.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 3 --team-steps 0 --qx-scale "" --cpu-steps 1 --example 0 > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_rtr_fused -s 3 -c 1 -o $O/fused_full -f \
  python bench.py --steps 2 --warmup 3 --team-steps 0 --qx-scale "" --cpu-steps 1 --example 0 > $O/ncu_fused.log 2>&1
ncu -i $O/fused_full.ncu-rep --page raw --csv > $O/fused_full_raw.csv 2>/dev/null
rm -f $O/fused_full.ncu-rep
python - <<'PY'
import json
b = json.load(open("gpurun_out/s9/bench_n1.json"))
print("N1", b["value"], b["ms_per_step"], b["e2e"]["value"], [ (q["n"], round(q["frac"],3)) for q in b["roofline"]["qx_scale"]])
print(b["fused_phase_ms"])
t = json.load(open("gpurun_out/s9/bench_team_1gpu.json"))
print("team", t["value"], t["e2e"]["value"], t["speedup_vs_1gpu_same_workload"], json.dumps(t["asynchronous_peer_mailboxes"])[:400])
for l in open("gpurun_out/s9/qx_scale.jsonl"):
    d = json.loads(l); print(d["L"], [(v["variant"], round(v["flushed_us"],1), round(v["frac_of_measured_peak"],3)) for v in d["variants"]])
PY
