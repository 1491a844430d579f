#!/bin/bash
# Round-2 GPU session 19 (1 GPU): first device run of the tensor-core Q*X (variant 4): parity, then at scale beside the
# other forms (uncapped grid and one resident wave).
O=gpurun_out/s19
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_a_parity.py -x -q -m gpu -k "qx" > $O/pytest_qx.log 2>&1; tail -12 $O/pytest_qx.log
for w in 0 1; do
  DPGO_QX_RESIDENT_WAVES=$w timeout 300 python tools/qx_scale.py 64 100 > $O/qx_w$w.tmp 2> $O/qx_w$w.err
  python - $w $O/qx_w$w.tmp >> $O/qx_variants.jsonl <<'PY'
import json, sys
for l in open(sys.argv[2]):
    d = json.loads(l); d["resident_waves"] = int(sys.argv[1]); print(json.dumps(d))
PY
  rm -f $O/qx_w$w.tmp
done
python - <<'PY'
import json
for l in open("gpurun_out/s19/qx_variants.jsonl"):
    d = json.loads(l)
    print(d["resident_waves"], d["n"], [(v["variant"], round(v["flushed_us"], 1), round(v["frac_of_measured_peak"], 3), v["rel_diff_vs_variant0"]) for v in d["variants"]])
PY
