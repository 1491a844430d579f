#!/bin/bash
# Round-2 GPU session 4: grid-barrier forms A/B (release/acquire counter, relaxed counter + fences, grid.sync with a
# CTA barrier in front), Q*X kernel variants at roofline scale with an ncu capture of the lane-group kernel, GPU
# suite incl. the native exchange on one device.
O=gpurun_out/s4
mkdir -p $O
( time timeout 900 python -m pytest tests/ -x -q -m gpu --durations=8 > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest_time.txt
tail -3 $O/pytest_gpu.log
for v in "" _bar2 _presync; do
  for rep in 1 2; do
    DPGO_B200_LIB=$PWD/dpgo_b200/libdpgo_b200$v.so timeout 200 python tools/dd_probe.py --barrier-ab \
      > $O/ab${v}_$rep.jsonl 2> $O/ab${v}_$rep.err
  done
done
for f in $O/ab*.jsonl; do echo $f; python -c "
import json,sys
for l in open('$f'):
    d=json.loads(l); print('  ',d['problem'],d['mode'],d['optimize_ms'])
"; done
timeout 500 python tools/qx_scale.py 64 100 > $O/qx_scale.jsonl 2> $O/qx_scale.err
python -c "
import json
for l in open('$O/qx_scale.jsonl'):
    d=json.loads(l); print(d['L'], [(v['variant'], round(v['flushed_us'],1), round(v['frac_of_measured_peak'],3), v['rel_diff_vs_variant0']) for v in d['variants']])
"
for v in 0 3; do
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_qx -s 3 -c 1 -o $O/qx_v$v -f \
  python tools/ncu_qx.py 64 $v 6 > $O/ncu_qx_v$v.log 2>&1
ncu -i $O/qx_v$v.ncu-rep --page details --csv > $O/qx_v${v}_details.csv 2>/dev/null
done
timeout 300 python bench.py --team-steps 0 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
ls $O
