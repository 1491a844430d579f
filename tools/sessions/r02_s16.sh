#!/bin/bash
# Round-2 GPU session 16 (1 GPU): device chordal initialization against the oracle on all fixtures, then the whole
# suite in the driver's order.
O=gpurun_out/s16
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_a_parity.py -x -q -m gpu -s -k chordal --durations=8 > $O/pytest_chordal.log 2>&1 ) 2> $O/pytest_chordal_time.txt
tail -12 $O/pytest_chordal.log; grep -h "rotation_iterations" $O/pytest_chordal.log | head
( time timeout 900 python -m pytest tests/ -x -q -m gpu --durations=8 > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest_time.txt
tail -14 $O/pytest_gpu.log; cat $O/pytest_time.txt
