mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc >> gpurun_out/r2a_smi.txt; free -g >> gpurun_out/r2a_smi.txt
for b in pb_ns5 pb_ns4 pb_first; do echo "== $b" >> gpurun_out/r2a_pb.txt; timeout 120 tools/bin/$b 4096 1024 >> gpurun_out/r2a_pb.txt 2>&1; done
timeout 1500 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r2a_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.txt 2> gpurun_out/r2a_bench.err; echo "bench rc=$?" >> gpurun_out/r2a_bench.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_ref.txt 2>&1
tail -5 gpurun_out/r2a_pytest.txt; tail -c 600 gpurun_out/r2a_pb.txt
