#!/bin/bash
# Round-2 GPU session 29 (1 GPU): L2 prefetch of the CTA's strips at the start of the solver kernel (cold first
# application: bench with flushed L2, several agents sharing one GPU): parity + the N=1 bench line.
O=gpurun_out/s29
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_a_parity.py tests/test_gpu_b_team.py -x -q -m gpu > $O/pytest.log 2>&1; tail -2 $O/pytest.log
( time timeout 600 python bench.py --example 0 > $O/bench_n1.json 2> $O/bench_n1.err ) 2> $O/bench_n1_time.txt; echo "bench n1 rc=$?"; grep real $O/bench_n1_time.txt
python - <<'PY'
import json
b = json.load(open("gpurun_out/s29/bench_n1.json"))
print("N1", b["value"], b["ms_per_step"], b["e2e"]["value"], b["solver_kernel_ms_per_step"]["l2_flushed"], b["solver_kernel_ms_per_step"]["l2_warm"], b["roofline"]["frac"])
print(b["fused_phase_ms"])
print(b["grid3D_8agents_all"]["value"], b["grid3D_8agents_colored"]["value"], b["roofline"]["dense_inverse_variant"]["frac"])
PY
