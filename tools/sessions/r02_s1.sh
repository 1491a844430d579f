#!/bin/bash
# Round-2 GPU session 1: cold-start diagnosis, driver-order GPU suite, first device run of the
# code written blind at the end of round 1 (modes 3/4, async solve, Q*X prefetch), bench lines.
mkdir -p gpurun_out/s1
O=gpurun_out/s1
date +%s > $O/t0
# (1) paging throughput of the toolkit libraries the round-1 .so links (cold file cache)
( time timeout 150 cat /usr/local/cuda/lib64/libcusolver.so.11 > /dev/null ) 2> $O/cold_cat_cusolver.txt
ls -l /usr/local/cuda/lib64/libcusolver.so.11* /usr/local/cuda/lib64/libcublasLt.so.12* >> $O/cold_cat_cusolver.txt 2>&1
( time timeout 400 dpgo_b200/host/bin/host_tests > $O/host_tests.log 2>&1 ) 2> $O/cold_host_tests_time.txt
# (2) the driver's command
( time timeout 1500 python -m pytest tests/ -x -q -m gpu --durations=25 > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest_time.txt
tail -5 $O/pytest_gpu.log
# (3) blind code
DPGO_B200_EXPERIMENTAL=1 timeout 600 python tests/gpu_three_phase_check.py --full > $O/three_phase.jsonl 2> $O/three_phase.err
echo "three_phase rc=$?"
DPGO_B200_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_x_qx_variant.py -x -q -m gpu > $O/qx_variant.log 2>&1
echo "qx_variant rc=$?"
# (4) bench lines
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --precon-mode 3 --team-steps 0 > $O/bench_n1_mode3.json 2> $O/bench_n1_mode3.err
timeout 300 python bench.py --precon-mode 4 --team-steps 0 > $O/bench_n1_mode4.json 2> $O/bench_n1_mode4.err
timeout 400 python tools/qx_scale.py 64 100 > $O/qx_scale.jsonl 2> $O/qx_scale.err
timeout 300 python tools/dd_probe.py --forms > $O/dd_probe.jsonl 2> $O/dd_probe.err
date +%s > $O/t1
tail -3 $O/bench_n1.json
