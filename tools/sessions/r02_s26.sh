#!/bin/bash
# Round-2 GPU session 26 (2 GPUs, final library): the N=2 bench line, both arms, and the 2-rank exchange / mailbox tests.
O=gpurun_out/s26
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 300 python -m pytest tests/test_gpu_d_async_solve.py -x -q -m gpu > $O/pytest_async.log 2>&1; tail -2 $O/pytest_async.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 \
  tests/gpu_native_exchange_2gpu.py grid3D 8 > $O/exch2_grid.log 2>&1; echo "exchange grid rc=$?"; grep -v "^\*\|OMP_NUM\|^$" $O/exch2_grid.log | tail -3
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 \
  bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err ) 2> $O/bench_n2_time.txt; echo "bench n2 rc=$?"; grep real $O/bench_n2_time.txt
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29643 \
  bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > $O/bench_n2_reference.json 2> $O/bench_n2_reference.err ) 2> $O/bench_n2_ref_time.txt
grep real $O/bench_n2_ref_time.txt
python - <<'PY'
import json
b = json.load(open("gpurun_out/s26/bench_n2.json"))
print("N2", b["value"], b["ms_per_step"], "speedup", b["speedup_vs_1gpu_same_workload"], "e2e", b["e2e"]["value"], "other", b["other_schedule"]["value"],
      b["other_schedule"]["speedup_vs_1gpu_same_workload"], b["parity"])
r = json.load(open("gpurun_out/s26/bench_n2_reference.json")); print("ref", r["value"], r.get("config") == b.get("config"), r["steps"], r["warmup"])
PY
