mkdir -p gpurun_out; rm -f gpurun_out/r2i_fit.txt
for n in 4096 16384 1500; do
  timeout 300 python tools/fit_bench.py $n 5 >> gpurun_out/r2i_fit.txt 2>&1
  GPMPC_NO_LOOKAHEAD=1 timeout 300 python tools/fit_bench.py $n 5 >> gpurun_out/r2i_fit.txt 2>&1
done
cat gpurun_out/r2i_fit.txt
timeout 600 python -m pytest tests -m gpu -q -x -k "fit or gpr or incremental or marginal or config4_size_fit or randomized" > gpurun_out/r2i_pytest.txt 2>&1; tail -3 gpurun_out/r2i_pytest.txt
