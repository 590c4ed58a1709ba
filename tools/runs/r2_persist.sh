mkdir -p gpurun_out
for cfg in "400 6" "400 30" "1024 30" "2048 30"; do
  python tools/b1_breakdown.py $cfg 30 2>&1 | grep -v "cost only\|device in" 
  PERSIST=1 python tools/b1_breakdown.py $cfg 30 2>&1 | grep -v "cost only\|device in"
done > gpurun_out/persist.log 2>&1
cat gpurun_out/persist.log
