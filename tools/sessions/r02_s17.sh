#!/bin/bash
# Round-2 GPU session 17 (4 GPUs): the N=4 bench lines again with the staged per-pose kernels (Nesterov updates) and
# the faster GNC weight refresh (symbolic part kept, 4-lanes-per-row diagonal block kernel): grid3D (configs[2]) and
# city10000 / 4 agents with GNC weight updates inside the timed rounds (configs[4]).
O=gpurun_out/s17
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29631 \
  bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_n4.json 2> $O/bench_n4.err ) 2> $O/bench_n4_time.txt; echo "bench n4 rc=$?"; cat $O/bench_n4_time.txt
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29632 \
  bench.py --gpus 4 --steps 20 --warmup 5 --team-dataset city10000 --team-agents 4 --team-r 3 --gnc-interval 5 \
  > $O/bench_n4_city10000_gnc.json 2> $O/bench_n4_city10000_gnc.err ) 2> $O/bench_n4_city_time.txt; echo "bench city rc=$?"; cat $O/bench_n4_city_time.txt
python - <<'PY'
import json
for f in ("bench_n4.json", "bench_n4_city10000_gnc.json"):
    try:
        b = json.load(open("gpurun_out/s17/" + f))
        print(f, b["value"], b["ms_per_step"], "speedup", b["speedup_vs_1gpu_same_workload"], "e2e", b["e2e"]["value"],
              "other", b["other_schedule"]["value"], b["other_schedule"]["speedup_vs_1gpu_same_workload"], b["parity"])
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -c 600 $O/bench_n4.err; tail -c 600 $O/bench_n4_city10000_gnc.err
