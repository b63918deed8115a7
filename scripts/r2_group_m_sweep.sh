for gm in 4 6 8 12 16 24; do
  echo "group_m=$gm"
  LIT_GEMM_GROUP_M=$gm python scripts/gpu_corr_gemm_only.py 4 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  events ms', [round(x,2) for x in d['ms'][1:]])"
  LIT_GEMM_GROUP_M=$gm ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:gemm_tf32x3_kernel -s 1 -c 1 python scripts/gpu_corr_gemm_only.py 2 2>/dev/null | grep -E "dram__bytes_read|gpu__time|hit_rate" | sed 's/^ */  /'
done
