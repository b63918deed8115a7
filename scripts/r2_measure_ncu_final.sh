# ncu evidence for the final code: the pair-output downdate GEMM next to the fp32-output one, and the launch list of one fit
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.sum,lts__t_sector_hit_rate.pct"
python scripts/gpu_f16_store_gemm_only.py 4
python scripts/gpu_f16_store_gemm_only.py 4 pairout
ncu --metrics $M --clock-control none -k regex:gemm_tf32x3_kernel -s 2 -c 1 python scripts/gpu_f16_store_gemm_only.py 4 2>/dev/null | grep -E "gemm_tf32x3|dram__|gpu__time|tensor|inst_exec|hit_rate" | sed 's/^ */  fp32out /' | tee gpurun_out/r2_pairout_gemm_ncu.txt
ncu --metrics $M --clock-control none -k regex:gemm_tf32x3_kernel -s 2 -c 1 python scripts/gpu_f16_store_gemm_only.py 4 pairout 2>/dev/null | grep -E "gemm_tf32x3|dram__|gpu__time|tensor|inst_exec|hit_rate" | sed 's/^ */  pairout /' | tee -a gpurun_out/r2_pairout_gemm_ncu.txt
timeout 400 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -s 2000 -c 1900 --csv --log-file gpurun_out/r2_launches_final.csv python scripts/gpu_host_floor.py 95000 > /dev/null 2>&1
grep -c "lit::" gpurun_out/r2_launches_final.csv
