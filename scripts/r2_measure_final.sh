# full GPU suite + smoke + bench at N=1 with the final code
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --tb=short 2>&1 | tail -40 > gpurun_out/final_tests.log
cat gpurun_out/final_tests.log | tail -25
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
tail -c 1500 gpurun_out/final_bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/final_bench_n1.json').read().splitlines() if l.startswith('{')][-1])
print('fit', d['fit_seconds'], 'e2e', d['e2e']['fit_seconds'], d['e2e'].get('pageable'))
print(d['phases_ms'])
print(d['e2e_phases_ms'])
print(d['result_check'])
PY
# full-scale parity of the final code (config 2 at V = 95,000; 8,192-voxel stripe against the CPU oracle)
timeout 600 python scripts/gpu_fullscale_parity.py --out gpurun_out/final_fullscale_parity.json 2>&1 | tail -2 | cut -c1-1500
