#!/bin/bash
# Round-2 GPU session 8 (4 GPUs): peer-memory mailboxes across processes, the N=4 bench lines of grid3D (configs[2]) and
# of city10000 / 4 agents with GNC weight updates inside the timed rounds (configs[4]).
O=gpurun_out/s8
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 300 python -m pytest tests/test_gpu_d_async_solve.py -x -q -m gpu > $O/pytest_async.log 2>&1; tail -3 $O/pytest_async.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29621 \
  tests/gpu_peer_mailboxes_2gpu.py torus3D 8 12 > $O/mailboxes_n4.log 2>&1; echo "mailboxes rc=$?"; grep -v "^\*\|OMP_NUM\|^$" $O/mailboxes_n4.log | tail -5
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29622 \
  bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_n4.json 2> $O/bench_n4.err ) 2> $O/bench_n4_time.txt; echo "bench n4 rc=$?"; cat $O/bench_n4_time.txt
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29623 \
  bench.py --gpus 4 --steps 20 --warmup 5 --team-dataset city10000 --team-agents 4 --team-r 3 --gnc-interval 5 \
  > $O/bench_n4_city10000_gnc.json 2> $O/bench_n4_city10000_gnc.err ) 2> $O/bench_n4_city_time.txt; echo "bench city rc=$?"; cat $O/bench_n4_city_time.txt
python - <<'PY'
import json
for f in ("bench_n4.json", "bench_n4_city10000_gnc.json"):
    try:
        b = json.load(open("gpurun_out/s8/" + f))
        print(f, b["value"], b["ms_per_step"], "speedup", b["speedup_vs_1gpu_same_workload"], "e2e", b["e2e"]["value"],
              "other", b["other_schedule"]["value"], b["other_schedule"]["speedup_vs_1gpu_same_workload"], b["parity"])
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -c 600 $O/bench_n4.err; tail -c 600 $O/bench_n4_city10000_gnc.err
