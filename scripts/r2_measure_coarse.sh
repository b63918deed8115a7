# coarse series terms: kernel tests, fit-level parity, bench at N=1
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --tb=short -k "series or golden or midsize or full_width or gemm_corr or smoke or dual_form" 2>&1 | tail -40 > gpurun_out/coarse_tests.log
tail -12 gpurun_out/coarse_tests.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/coarse_bench_n1.json 2> gpurun_out/coarse_bench_n1.err
tail -c 1000 gpurun_out/coarse_bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/coarse_bench_n1.json').read().splitlines() if l.startswith('{')][-1])
print('fit', d['fit_seconds'], 'e2e', d['e2e']['fit_seconds'], d['e2e'].get('pageable'))
print(d['phases_ms'])
print(d['roofline']['launch_ms'], d['roofline']['frac'], d['roofline'].get('tensor_pipe'))
print(d['result_check'])
PY
