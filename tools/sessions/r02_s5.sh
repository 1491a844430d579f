#!/bin/bash
# Round-2 GPU session 5 (1 GPU): 256-bit loads in the lane-group SpMM (Q*X at scale, fused solver), the new bench
# line (qx_scale, anchors through dpgo_exchange), launch list + full ncu capture of the fused kernel.
O=gpurun_out/s5
mkdir -p $O
( time timeout 900 python -m pytest tests/ -x -q -m gpu --durations=8 > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest_time.txt
tail -3 $O/pytest_gpu.log
timeout 500 python tools/qx_scale.py 64 100 > $O/qx_scale.jsonl 2> $O/qx_scale.err
python -c "
import json
for l in open('$O/qx_scale.jsonl'):
    d=json.loads(l); print(d['L'], [(v['variant'], round(v['flushed_us'],1), round(v['frac_of_measured_peak'],3), v['rel_diff_vs_variant0']) for v in d['variants']])
"
( time timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err ) 2> $O/bench_time.txt; echo "bench rc=$?"; cat $O/bench_time.txt
tail -c 600 $O/bench_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_n1_reference.json 2> $O/bench_n1_reference.err
timeout 300 python tools/dd_probe.py --barrier-ab > $O/probe.jsonl 2> $O/probe.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 3 --team-steps 0 --qx-scale "" --cpu-steps 1 > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_rtr_fused -s 3 -c 1 -o $O/fused_full -f \
  python bench.py --steps 2 --warmup 3 --team-steps 0 --qx-scale "" --cpu-steps 1 > $O/ncu_fused.log 2>&1
ncu -i $O/fused_full.ncu-rep --page details --csv > $O/fused_full_details.csv 2>/dev/null
ncu -i $O/fused_full.ncu-rep --page raw --csv > $O/fused_full_raw.csv 2>/dev/null
ls -la $O
