CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:pass_tc -s 2 -c 2 --csv --log-file gpurun_out/traffic_full.csv $CMD > gpurun_out/ncu_traffic.log 2>&1; echo "ncu rc=$?"; cat gpurun_out/traffic_full.csv | tail -8
