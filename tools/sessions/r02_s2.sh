#!/bin/bash
# Round-2 GPU session 2: fresh-box driver-order GPU suite with the in-tree dense linear algebra (no cuSOLVER /
# cuBLAS), smoke wall time, preconditioner set-up times, per-CTA phase trace, ncu of the strip kernels.
O=gpurun_out/s2
mkdir -p $O
( time timeout 1300 python -m pytest tests/ -x -q -m gpu --durations=15 > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest_time.txt
tail -4 $O/pytest_gpu.log; cat $O/pytest_time.txt
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1 ) 2> $O/smoke_time.txt
tail -2 $O/smoke.log; cat $O/smoke_time.txt
timeout 300 python tools/setup_time.py > $O/setup_time.jsonl 2> $O/setup_time.err; cat $O/setup_time.jsonl
DPGO_B200_LIB=$PWD/dpgo_b200/libdpgo_b200_trace.so timeout 300 python tools/phase_trace.py sphere2500 5 2 4 > $O/phase_trace.jsonl 2> $O/phase_trace.err
timeout 300 python bench.py --team-steps 0 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --precon-mode 3 --team-steps 0 > $O/bench_n1_mode3.json 2> $O/bench_n1_mode3.err
# ncu: strip kernels of the three-phase apply (L2-resident), full sections on 6 launches after warm-up
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_strip_gemv3 -s 30 -c 6 -o $O/strip3 -f \
  python tools/ncu_precon_apply.py sphere2500 5 3 > $O/ncu_strip3.log 2>&1
ncu -i $O/strip3.ncu-rep --page details --csv > $O/strip3_details.csv 2>/dev/null
ls -la $O
