mkdir -p gpurun_out
GPMPC_STEP_DEBUG=1 timeout 120 python tools/b1_eval.py 2 1024 4 > gpurun_out/r2j_dbg1024.txt 2>&1
GPMPC_STEP_DEBUG=1 timeout 120 python tools/b1_eval.py 2 4096 4 > gpurun_out/r2j_dbg4096.txt 2>&1
grep "^step" gpurun_out/r2j_dbg1024.txt | tail -4; grep "^step" gpurun_out/r2j_dbg4096.txt | tail -4
