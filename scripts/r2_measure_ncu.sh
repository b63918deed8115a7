ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/r2_streaming_ncu.csv python scripts/gpu_stream_bench.py > /dev/null 2>&1
grep -c "lit::" gpurun_out/r2_streaming_ncu.csv
python scripts/gpu_stream_bench.py > gpurun_out/r2_streaming_events.jsonl 2>/dev/null
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
grep -c "lit::" gpurun_out/r2_launches_bench.csv
