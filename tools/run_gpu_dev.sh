CMD="python bench.py --rows 65536 --cols 32768 --steps 2 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tc.csv $CMD > gpurun_out/ncu_l.log 2>&1; echo "ncu launches rc=$?"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pass_tc -s 2 -c 2 -o gpurun_out/prof_tc5 $CMD > gpurun_out/ncu_tc.log 2>&1; echo "ncu rc=$?"
