#!/bin/bash
# Round-2 GPU session 30 (1 GPU): the round's final library: whole suite in the driver's order, smoke(), the N=1 bench
# line (both arms), ncu launch list + full capture of the fused solver.
O=gpurun_out/s30
mkdir -p $O
( time timeout 900 python -m pytest tests/ -x -q -m gpu --durations=6 > $O/pytest_gpu.log 2>&1 ) 2> $O/pytest_time.txt
tail -3 $O/pytest_gpu.log; grep real $O/pytest_time.txt
( time python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1 ) 2> $O/smoke_time.txt; tail -1 $O/smoke.log; grep real $O/smoke_time.txt
( time timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err ) 2> $O/bench_n1_time.txt; echo "bench n1 rc=$?"; grep real $O/bench_n1_time.txt; tail -c 300 $O/bench_n1.err
( time timeout 600 python bench.py --impl reference > $O/bench_n1_reference.json 2> $O/bench_n1_reference.err ) 2> $O/bench_n1_ref_time.txt; echo "reference rc=$?"; grep real $O/bench_n1_ref_time.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 3 --team-steps 0 --qx-scale "" --cpu-steps 1 --example 0 > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
python tools/ncu_digest.py launches $O/launches.csv $O/launches_summary.json | head -6
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_rtr_fused -s 3 -c 1 -o $O/fused_full -f \
  python bench.py --steps 2 --warmup 3 --team-steps 0 --qx-scale "" --cpu-steps 1 --example 0 > $O/ncu_fused.log 2>&1
ncu -i $O/fused_full.ncu-rep --page raw --csv > $O/fused_full_raw.csv 2>/dev/null
python tools/ncu_digest.py full $O/fused_full_raw.csv $O/fused_traffic.json k_rtr_fused | tail -1
rm -f $O/fused_full.ncu-rep
python - <<'PY'
import json
b = json.load(open("gpurun_out/s30/bench_n1.json"))
print("N1", b["value"], b["ms_per_step"], b["e2e"]["value"], b["solver_kernel_ms_per_step"]["l2_flushed"], b["roofline"]["frac"],
      [(q["n"], round(q["frac"], 3), {k: round(v["frac"], 3) for k, v in q["pose_kernels"].items()}) for q in b["roofline"]["qx_scale"]])
print(b["fused_phase_ms"]); print(b.get("chordal_initialization")); print(b["cpu_baseline"].get("chordal_initialization"))
print([ (m["dataset"], round(m["iterations_per_s"],1), round(m["speedup_vs_cpu_driver"],2)) for m in b["multi_robot_example"]])
print(b["grid3D_8agents_all"]["value"], b["grid3D_8agents_colored"]["value"], b["roofline"]["dense_inverse_variant"]["frac"])
r = json.load(open("gpurun_out/s30/bench_n1_reference.json")); print("ref", r["value"], r.get("config") == b.get("config"))
PY
