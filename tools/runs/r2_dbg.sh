mkdir -p gpurun_out
for n in 400 1024 4096; do
GPMPC_STEP_DEBUG=1 python tools/b1_eval.py 2 $n 6 2>&1 | grep "^step" | head -12 | tail -6 > gpurun_out/dbg$n.log
done
cat gpurun_out/dbg400.log gpurun_out/dbg1024.log gpurun_out/dbg4096.log
