#!/bin/bash
# Round-2 GPU session 14 (1 GPU): first device run of the folded phases of the fused two-level solver (tCG update
# folded into the first strip pass, direction update folded into the Hessian pass, interior wave prefetched across
# tCG iterations): parity suite, then A/B of the builds on one box.
O=gpurun_out/s14
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_a_parity.py tests/test_gpu_b_team.py -x -q -m gpu > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest_time.txt
tail -3 $O/pytest_gpu.log
for v in product nofold pf pffs product nofold; do
  lib=dpgo_b200/libdpgo_b200_$v.so
  [ $v = product ] && lib=dpgo_b200/libdpgo_b200.so
  DPGO_B200_LIB=$lib timeout 200 python tools/dd_probe.py --fold-ab > $O/fold_$v.tmp 2> $O/fold_$v.err
  python - $v $O/fold_$v.tmp >> $O/fold_ab.jsonl <<'PY'
import json, sys
for l in open(sys.argv[2]):
    d = json.loads(l); d["build"] = sys.argv[1]; print(json.dumps(d))
PY
  rm -f $O/fold_$v.tmp
done
python - <<'PY'
import json
for l in open("gpurun_out/s14/fold_ab.jsonl"):
    d = json.loads(l)
    print(d["build"], d["problem"], d["optimize_ms"], d["outer"], d["tcg"], d["barriers"], repr(d["two_f"]), [d["phase_ms"][i] for i in (8, 9, 10, 11, 12, 2, 5, 3, 4)])
PY
timeout 300 python tools/pose_op_scale.py 64 100 > $O/pose_op_scale.jsonl 2> $O/pose_op_scale.err; cat $O/pose_op_scale.jsonl; tail -2 $O/pose_op_scale.err
