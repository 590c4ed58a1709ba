mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -x > gpurun_out/fin_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/fin_pytest.txt
tail -3 gpurun_out/fin_pytest.txt
timeout 60 python bench.py --B 128 --steps 2 --warmup 3 --no-cpu-baseline --no-latency --no-extras > gpurun_out/fin_b128.txt 2>&1
tail -1 gpurun_out/fin_b128.txt | cut -c1-200
timeout 60 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-latency --no-extras > gpurun_out/fin_b1024.txt 2>&1
tail -1 gpurun_out/fin_b1024.txt | cut -c1-200
