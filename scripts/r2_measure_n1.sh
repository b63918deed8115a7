python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_23.json 2> gpurun_out/r2_bench_23.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_23.json"))
print(d["ms_per_step"], d["e2e"]["fit_seconds"], d["e2e"]["pageable"], d["gpu_launches"])
print(d["phases_ms"]); print(d["result_check"]); print(d["cpu_baseline"]["value"], d["cpu_baseline"]["sample_s"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["kind"])
PY
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:lit:: --csv --log-file gpurun_out/r2_streaming_ncu.csv python scripts/gpu_stream_bench.py > gpurun_out/r2_streaming_events.jsonl 2> /dev/null
tail -2 gpurun_out/r2_streaming_events.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lit:: --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
wc -l gpurun_out/r2_launches_bench.csv
(time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_reference_arm.json 2> gpurun_out/r2_reference_arm.err) 2>&1 | tail -3
cut -c1-600 gpurun_out/r2_reference_arm.json
