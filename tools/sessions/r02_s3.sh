#!/bin/bash
# Round-2 GPU session 3: hand-rolled grid barrier vs cooperative-groups grid.sync (A/B builds on one box),
# staged Q*X kernel at roofline scale, GPU suite with the new tests.
O=gpurun_out/s3
mkdir -p $O
( time timeout 900 python -m pytest tests/ -x -q -m gpu --durations=8 > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest_time.txt
tail -3 $O/pytest_gpu.log
for v in "" _cgsync _presync; do
  for rep in 1 2; do
    DPGO_B200_LIB=$PWD/dpgo_b200/libdpgo_b200$v.so timeout 200 python tools/dd_probe.py --barrier-ab \
      > $O/ab${v}_$rep.jsonl 2> $O/ab${v}_$rep.err
  done
done
grep -h optimize_ms $O/ab*.jsonl | cut -c1-400
timeout 500 python tools/qx_scale.py 64 100 > $O/qx_scale.jsonl 2> $O/qx_scale.err
cat $O/qx_scale.jsonl | cut -c1-1500
timeout 300 python bench.py --team-steps 0 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
