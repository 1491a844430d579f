#!/bin/bash
# Round-2 GPU session 6 (2 GPUs): library after the removal of the losing forms; Q*X at scale with the clean flush;
# NCCL exchange inside the C-ABI across two ranks; the N=2 bench line (both arms) and the N=1 line.
O=gpurun_out/s6
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
( time timeout 900 python -m pytest tests/ -x -q -m gpu --durations=8 > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest_time.txt
tail -3 $O/pytest_gpu.log
timeout 300 python tools/qx_scale.py 64 100 > $O/qx_scale.jsonl 2> $O/qx_scale.err
python -c "
import json
for l in open('$O/qx_scale.jsonl'):
    d=json.loads(l); print(d['L'], [(v['variant'], round(v['flushed_us'],1), round(v['frac_of_measured_peak'],3), v['rel_diff_vs_variant0']) for v in d['variants']])
"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
  tests/gpu_native_exchange_2gpu.py smallGrid3D 5 > $O/exch2_small.log 2>&1; echo "exchange small rc=$?"; tail -4 $O/exch2_small.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 \
  tests/gpu_native_exchange_2gpu.py grid3D 8 > $O/exch2_grid.log 2>&1; echo "exchange grid rc=$?"; tail -4 $O/exch2_grid.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 \
  bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err ) 2> $O/bench_n2_time.txt; echo "bench n2 rc=$?"; cat $O/bench_n2_time.txt
tail -c 1500 $O/bench_n2.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29615 \
  bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > $O/bench_n2_reference.json 2> $O/bench_n2_reference.err ) 2> $O/bench_n2_ref_time.txt
cat $O/bench_n2_ref_time.txt
( time timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err ) 2> $O/bench_n1_time.txt; echo "bench n1 rc=$?"; cat $O/bench_n1_time.txt
tail -c 800 $O/bench_n1.err
ls $O
