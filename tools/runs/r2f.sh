mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r2f_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.txt
tail -18 gpurun_out/r2f_pytest.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench.txt 2> gpurun_out/r2f_bench.err; echo "bench rc=$?" >> gpurun_out/r2f_bench.err
timeout 300 python bench.py --ard --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-latency > gpurun_out/r2f_bench_ard.txt 2>&1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-latency --no-extras"
$CMD > gpurun_out/r2f_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mm_pairs_batch -s 65 -c 1 -o gpurun_out/r2f_pairs $CMD > gpurun_out/r2f_ncu.log 2>&1
cut -c1-400 gpurun_out/r2f_bench.txt; cut -c1-300 gpurun_out/r2f_bench_ard.txt | tail -2
