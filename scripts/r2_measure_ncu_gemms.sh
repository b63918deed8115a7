python -m pytest tests -m gpu -x -q -k "pageable or empty_voxel" 2>&1 | tail -2
python scripts/gpu_f16_store_gemm_only.py 4
python scripts/gpu_corr_gemm_only.py 3
ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3_kernel -s 2 -c 1 -o gpurun_out/r2_f16_store_gemm python scripts/gpu_f16_store_gemm_only.py 4 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3_kernel -s 1 -c 1 -o gpurun_out/r2_corr_gemm python scripts/gpu_corr_gemm_only.py 3 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
