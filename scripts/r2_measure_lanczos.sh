# adaptive Lanczos: kernel test, fit-level parity, bench at N=1
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --tb=short -k "inner_solver_kernels or golden or midsize or full_width or dual_form or direct" 2>&1 | tail -30 > gpurun_out/lanczos_tests.log
tail -8 gpurun_out/lanczos_tests.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/lanczos_bench_n1.json 2> gpurun_out/lanczos_bench_n1.err
tail -c 600 gpurun_out/lanczos_bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/lanczos_bench_n1.json').read().splitlines() if l.startswith('{')][-1])
print('fit', d['fit_seconds'], 'e2e', d['e2e']['fit_seconds'], d['e2e'].get('pageable'))
print(d['phases_ms'])
print(d['roofline']['launch_ms'], d['roofline']['frac'], d['gpu_launches'])
print(d['result_check'])
PY
