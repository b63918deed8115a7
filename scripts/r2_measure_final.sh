python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_final.json"))
print(d["ms_per_step"], d["e2e"]["fit_seconds"], d["e2e"]["pageable"], d["gpu_launches"])
print(d["phases_ms"]); print(d["result_check"]); print(d["clocks"]); print(d["cpu_baseline"]["value"], d["cpu_baseline"]["sample_s"])
PY
python scripts/gpu_fullscale_parity.py --stripe 8192 --out gpurun_out/r2_fullscale_parity_final.json > gpurun_out/r2_fullscale_parity_final.log 2>&1; tail -c 1500 gpurun_out/r2_fullscale_parity_final.log
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -s 600 -c 2400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
grep -c "lit::" gpurun_out/r2_launches_bench.csv
ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3_kernel -s 260 -c 40 -o gpurun_out/r2_gemms_full python scripts/gpu_host_floor.py 95000 > gpurun_out/r2_gemms_full.log 2>&1; tail -2 gpurun_out/r2_gemms_full.log; ls -la gpurun_out/r2_gemms_full.ncu-rep
