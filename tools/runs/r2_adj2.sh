mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/adj2_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/adj2_pytest.txt
tail -5 gpurun_out/adj2_pytest.txt
python tools/b1_eval.py 20 > gpurun_out/adj2_b1.log 2>&1
python tools/b1_eval.py 20 400 6 >> gpurun_out/adj2_b1.log 2>&1
python tools/b1_eval.py 20 1024 30 >> gpurun_out/adj2_b1.log 2>&1
cat gpurun_out/adj2_b1.log
