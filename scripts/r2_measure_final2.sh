python -m pytest tests -m gpu -x -q -k "pageable or direct_solver or direct_outer or golden or full_width_matches" 2>&1 | grep -E "^E  |Error|passed|failed" | head -20
python scripts/gpu_direct_solver_bench.py 25 2 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_final.json"))
print(d["ms_per_step"], d["e2e"]["fit_seconds"], d["e2e"]["pageable"], d["gpu_launches"])
print(d["phases_ms"]); print(d["result_check"]["matches_expected"]); print(d["clocks"])
PY
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -s 2200 -c 2000 --csv --log-file gpurun_out/r2_launches_bench.csv python scripts/gpu_host_floor.py 95000 > /dev/null 2>&1
grep -c "lit::" gpurun_out/r2_launches_bench.csv
