mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2z_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest.txt
tail -14 gpurun_out/r2z_pytest.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2z_bench.txt 2> gpurun_out/r2z_bench.err; echo "bench rc=$?" >> gpurun_out/r2z_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_ref.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.txt 2>&1
CMD="python tools/b1_eval.py 3"
$CMD > gpurun_out/r2z_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mm_step_single -s 40 -c 1 -o gpurun_out/r2z_single $CMD > gpurun_out/r2z_ncu.log 2>&1
cut -c1-300 gpurun_out/r2z_bench.txt; cat gpurun_out/r2z_smoke.txt | tail -2
