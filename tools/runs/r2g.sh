mkdir -p gpurun_out
timeout 300 python tools/b1_eval.py 20 4096 30 > gpurun_out/r2g_b1.txt 2>&1
timeout 300 python tools/b1_eval.py 8 16384 20 >> gpurun_out/r2g_b1.txt 2>&1
timeout 300 python tools/b1_eval.py 20 2048 30 >> gpurun_out/r2g_b1.txt 2>&1
cat gpurun_out/r2g_b1.txt
timeout 600 python -m pytest tests -m gpu -q -x -k "dimension_sweep or kernel_switch or randomized or persistent or batched_rollouts_vs" > gpurun_out/r2g_pytest.txt 2>&1; tail -3 gpurun_out/r2g_pytest.txt
