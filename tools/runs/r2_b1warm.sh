# warm-cache per-kernel durations of one B=1 evaluation (ncu serialises the launches, caches are NOT flushed)
mkdir -p gpurun_out
CMD="python tools/b1_eval.py 20"
$CMD > gpurun_out/b1warm_plain.log 2>&1 && ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum -s 400 -c 120 --csv --log-file gpurun_out/b1warm.csv $CMD > gpurun_out/b1warm_ncu.log 2>&1
python tools/b1_eval.py 20 400 6 > gpurun_out/b1warm_n400.log 2>&1
cat gpurun_out/b1warm_plain.log gpurun_out/b1warm_n400.log
