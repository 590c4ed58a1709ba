mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "two_devices or split_over_two or persistent or unequal" > gpurun_out/r2v_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest.txt
tail -4 gpurun_out/r2v_pytest.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR bench.py --config 7 --gpus 2 --ntrain 16384 --H 20 --steps 5 --warmup 2 > gpurun_out/r2v_split16k_n2.txt 2>&1
timeout 300 $TR bench.py --config 7 --gpus 2 --ntrain 4096 --H 30 --steps 20 --warmup 3 > gpurun_out/r2v_split4k_n2.txt 2>&1
tail -n 1 gpurun_out/r2v_split16k_n2.txt gpurun_out/r2v_split4k_n2.txt | cut -c1-700
PERSIST=1 python tools/b1_breakdown.py 4096 30 30 2>&1 | head -3
