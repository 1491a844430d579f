#!/bin/bash
# Round-2 GPU session 25 (8 GPUs, final library): strong scaling of grid3D / 8 agents (BASELINE configs[2]) and torus3D / 8 agents
# (configs[3]: all-agents schedule and the asynchronous peer-mailbox series), one agent per GPU.
O=gpurun_out/s25
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 \
  bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err ) 2> $O/bench_n8_time.txt; echo "bench n8 rc=$?"; cat $O/bench_n8_time.txt
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29632 \
  bench.py --gpus 8 --steps 20 --warmup 5 --team-dataset torus3D > $O/bench_n8_torus3D.json 2> $O/bench_n8_torus3D.err ) 2> $O/bench_n8_torus_time.txt; echo "bench torus rc=$?"; cat $O/bench_n8_torus_time.txt
python - <<'PY'
import json
for f in ("bench_n8.json", "bench_n8_torus3D.json"):
    try:
        b = json.load(open("gpurun_out/s25/" + f))
        print(f, b["value"], b["ms_per_step"], "speedup", b["speedup_vs_1gpu_same_workload"], "warm", b["warm_l2"]["value"],
              "e2e", b["e2e"]["value"], "other", b["other_schedule"]["value"], b["other_schedule"]["speedup_vs_1gpu_same_workload"],
              "async", (b.get("asynchronous_peer_mailboxes") or {}).get("value"), b["parity"])
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -c 500 $O/bench_n8.err; tail -c 500 $O/bench_n8_torus3D.err
