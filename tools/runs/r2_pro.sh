mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pro_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pro_pytest.txt
tail -5 gpurun_out/pro_pytest.txt
python tools/b1_breakdown.py 4096 30 > gpurun_out/pro_bd.log 2>&1
python tools/b1_breakdown.py 400 6 >> gpurun_out/pro_bd.log 2>&1
python tools/b1_breakdown.py 1024 30 >> gpurun_out/pro_bd.log 2>&1
cat gpurun_out/pro_bd.log
