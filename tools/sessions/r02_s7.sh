#!/bin/bash
# Round-2 GPU session 7 (1 GPU): GNC rounds through the native exchange, Q*X variants with the clean flush
# (incl. the shared-memory tile-staged form), 1-GPU anchors of torus3D / city10000, ncu of the Q*X forms.
O=gpurun_out/s7
mkdir -p $O
( time timeout 900 python -m pytest tests/ -x -q -m gpu --durations=8 > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest_time.txt
tail -3 $O/pytest_gpu.log
timeout 400 python tools/qx_scale.py 64 100 > $O/qx_scale.jsonl 2> $O/qx_scale.err
python -c "
import json
for l in open('$O/qx_scale.jsonl'):
    d=json.loads(l); print(d['L'], [(v['variant'], round(v['back_to_back_us'],1), round(v['flushed_us'],1), round(v['frac_of_measured_peak'],3), v['rel_diff_vs_variant0']) for v in d['variants']])
"
timeout 400 python - > $O/anchors.jsonl 2> $O/anchors.err <<'PY'
import json, sys
sys.path.insert(0, '.')
from tools import bench_team
for ds, gnc in ((dict(dataset='city10000', agents=4, r=3), 5), (dict(dataset='city10000', agents=4, r=3), 0),
                (dict(dataset='torus3D', agents=8, r=5), 0)):
    for sched in ('all', 'colored'):
        t = bench_team.measure(20, 5, 0, 1, 0, schedule=sched, mode='device', gnc=gnc, **ds)
        c2, gn = bench_team.central_eval(t['X'], ds['dataset'], ds['r'], 0)
        print(json.dumps(dict(ds, schedule=sched, gnc=gnc, value=t['value'], ms_per_step=t['ms_per_step'],
                              weight_updates=t['weight_updates'], cost2=c2, gradnorm=gn)), flush=True)
PY
cat $O/anchors.jsonl; tail -3 $O/anchors.err
for v in 0 2; do
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_qx -s 3 -c 1 -o $O/qx_v$v -f \
  python tools/ncu_qx.py 64 $v 6 > $O/ncu_qx_v$v.log 2>&1
ncu -i $O/qx_v$v.ncu-rep --page details --csv > $O/qx_v${v}_details.csv 2>/dev/null
done
ls $O
