# last check of the final commit: full GPU suite, smoke, default bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu --tb=short 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/verify_bench_n1.json 2> gpurun_out/verify_bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/verify_bench_n1.json').read().splitlines() if l.startswith('{')][-1])
print('fit', d['fit_seconds'], 'e2e', d['e2e']['fit_seconds'], d['e2e'].get('pageable'), 'steps', d['steps'], d['warmup'])
print(d['roofline']['frac'], d['roofline']['traffic'], d['gpu_launches'], d['clocks'])
print(d['result_check']['matches_expected'], d['cpu_baseline']['value'])
PY
