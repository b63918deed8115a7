for w in config1_wordrate_9400x4x95000 config3_whisper_gpt2_9400x5120x95000 config4_narratives_2226x3072x81924 config5_llama_9400x16384x95000; do
  python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_bench_$w.json") if l.startswith("{")][-1])
    print("$w", d["ms_per_step"], d["e2e"]["fit_seconds"], d["gpu_launches"], d["result_check"]["n_significant"], {k: v for k, v in d["phases_ms"].items() if k.startswith("phase") or k in ("eig","gemm_corr","spd_solve")})
except Exception as e:
    print("$w FAILED", e)
PY
  tail -2 gpurun_out/r2_bench_$w.err | cut -c1-300
done
