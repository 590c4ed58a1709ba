mkdir -p gpurun_out
: > gpurun_out/big3.log
for big in 620 640 660 680 640 660; do
  BIG=$big python tools/b1_breakdown.py 4096 30 60 2>&1 | grep "single_big\|host in/out, cost+grad" >> gpurun_out/big3.log
done
for big in 640 660 680; do
  BIG=$big python tools/b1_breakdown.py 16384 20 10 2>&1 | grep "single_big\|host in/out, cost+grad" >> gpurun_out/big3.log
done
cat gpurun_out/big3.log
