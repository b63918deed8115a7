python -m pytest tests -m gpu -q 2>&1 | tail -3
for w in config2_gpt2_9400x3072x95000 config4_narratives_2226x3072x81924; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_bench_$w.json") if l.startswith("{")][-1])
    print("$w", d["ms_per_step"], d["e2e"]["fit_seconds"], d["e2e"]["pageable"]["fit_seconds"], d["gpu_launches"], d["result_check"], d["phases_ms"])
except Exception as e:
    print("$w FAILED", e)
PY
  tail -2 gpurun_out/r2_bench_$w.err | cut -c1-300
done
