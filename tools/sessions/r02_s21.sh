#!/bin/bash
# Round-2 GPU session 21 (1 GPU): ncu --set full of the two Jacobi per-pose kernels at 262 144 poses (what bounds them).
O=gpurun_out/s21
mkdir -p $O
for k in k_polar k_round; do
  timeout 300 ncu --set full --clock-control none -k "regex:$k" -s 5 -c 1 -o $O/${k}_full -f \
    python tools/pose_op_scale.py 64 > $O/ncu_$k.log 2>&1
  ncu -i $O/${k}_full.ncu-rep --page raw --csv > $O/${k}_full_raw.csv 2>/dev/null
  rm -f $O/${k}_full.ncu-rep
  python tools/ncu_digest.py full $O/${k}_full_raw.csv $O/${k}_full.json $k | tail -1
done
