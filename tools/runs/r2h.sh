mkdir -p gpurun_out
timeout 300 python tools/b1_eval.py 20 4096 30 > gpurun_out/r2h_b1.txt 2>&1
timeout 300 python tools/b1_eval.py 8 16384 20 >> gpurun_out/r2h_b1.txt 2>&1
timeout 300 python tools/b1_eval.py 20 1024 30 >> gpurun_out/r2h_b1.txt 2>&1
cat gpurun_out/r2h_b1.txt
