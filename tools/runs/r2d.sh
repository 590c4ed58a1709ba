mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -q -x -k "persistent or config3_full_size_rollout or shipped or kernel_switch or determinism" > gpurun_out/r2d_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.txt
tail -5 gpurun_out/r2d_pytest.txt
timeout 300 python tools/b1_eval.py 20 > gpurun_out/r2d_b1.txt 2>&1
GPMPC_NO_PERSISTENT=1 timeout 300 python tools/b1_eval.py 20 >> gpurun_out/r2d_b1.txt 2>&1
timeout 300 python tools/config_bench.py 1 >> gpurun_out/r2d_b1.txt 2>&1
GPMPC_NO_PERSISTENT=1 timeout 300 python tools/config_bench.py 1 >> gpurun_out/r2d_b1.txt 2>&1
timeout 300 python bench.py --config 7 --n 16384 --H 20 --steps 5 --warmup 2 >> gpurun_out/r2d_b1.txt 2>&1
cat gpurun_out/r2d_b1.txt
