mkdir -p gpurun_out
python tools/b1_breakdown.py 4096 30 > gpurun_out/b1bd.log 2>&1
python tools/b1_breakdown.py 400 6 >> gpurun_out/b1bd.log 2>&1
python tools/b1_breakdown.py 1024 30 >> gpurun_out/b1bd.log 2>&1
CMD="python tools/b1_eval.py 20"
$CMD > gpurun_out/b1warm2_plain.log 2>&1 && ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum -s 400 -c 120 --csv --log-file gpurun_out/b1warm2.csv $CMD > gpurun_out/b1warm2_ncu.log 2>&1
cat gpurun_out/b1bd.log
