mkdir -p gpurun_out
: > gpurun_out/big2.log
for big in 660 700 740 780 820; do
  BIG=$big python tools/b1_breakdown.py 4096 30 40 2>&1 | grep -v "cost only\|device in\|by events" >> gpurun_out/big2.log
done
for big in 660 700 740 780; do
  BIG=$big python tools/b1_breakdown.py 16384 20 10 2>&1 | grep -v "cost only\|device in\|by events" >> gpurun_out/big2.log
done
cat gpurun_out/big2.log
