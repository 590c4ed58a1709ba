mkdir -p gpurun_out
GPMPC_STEP_DEBUG=1 GPMPC_STEP_DEBUG_FILE=gpurun_out/cta_timeline_slices.csv python tools/b1_eval.py 1 4096 3 2>&1 | grep "^step" | head -3 > gpurun_out/dbg_slices.log
cat gpurun_out/dbg_slices.log
