# producer-written fp16 pairs: kernel tests, fit-level agreement, bench at N=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "producer or golden or midsize or corr_precisions or dual_form or full_width or empty_voxel" 2>&1 | tail -15 > gpurun_out/pairs_tests.log
cat gpurun_out/pairs_tests.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/pairs_bench_n1.json 2> gpurun_out/pairs_bench_n1.err
tail -c 3000 gpurun_out/pairs_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/pairs_bench_n1.json').read().strip().splitlines()[-1])
print('fit', d['fit_seconds'], 'e2e', d['e2e']['fit_seconds'], d['e2e'].get('pageable'))
print(d['phases_ms'])
print(d['result_check'])
PY
