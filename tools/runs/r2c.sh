mkdir -p gpurun_out
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench.txt 2> gpurun_out/r2c_bench.err; echo "bench rc=$?" >> gpurun_out/r2c_bench.err
timeout 900 python bench.py --config 5 --instances 2048 > gpurun_out/r2c_cfg5_n1.txt 2>&1
timeout 600 python tools/fullcov_eval.py 16384 128 20 1 > gpurun_out/r2c_fullcov.txt 2>&1
timeout 300 python tools/fullcov_eval.py 16384 32 20 1 >> gpurun_out/r2c_fullcov.txt 2>&1
timeout 300 python tools/fullcov_eval.py 4096 1024 30 1 >> gpurun_out/r2c_fullcov.txt 2>&1
# ncu: launch list of the headline step, then the top kernels (each plain run first)
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-latency --no-extras"
$CMD > gpurun_out/r2c_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/r2c_launches_batch.csv $CMD > gpurun_out/r2c_ncu1.log 2>&1
$CMD > gpurun_out/r2c_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mm_pairs_batch -s 65 -c 1 -o gpurun_out/r2c_pairs $CMD > gpurun_out/r2c_ncu2.log 2>&1
CMD2="python tools/fullcov_eval.py 16384 128 1 0"
$CMD2 > gpurun_out/r2c_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mm_full_pairs -c 2 -o gpurun_out/r2c_full $CMD2 > gpurun_out/r2c_ncu3.log 2>&1
ls -la gpurun_out | tail -20
